#!/bin/bash
# Dev helper: build kernel variants HERE (nvcc cross-compiles) into noneq_opt_b200/variants/<name>.so so that one GPU
# call can time several of them (NQ_B200_LIB=<path> selects the library).  Restores the product build at the end.
# usage: scripts/build_variants.sh name1 "<nvcc flags 1>" name2 "<nvcc flags 2>" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p noneq_opt_b200/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  NQ_EXTRA_NVCC_FLAGS="$flags" python -m noneq_opt_b200.build --force > /dev/null
  cp noneq_opt_b200/libnoneq_b200.so noneq_opt_b200/variants/$name.so
  echo "built $name: $flags"
  python scripts/hot_loop_static.py noneq_opt_b200/variants/$name.so | head -3
done
python -m noneq_opt_b200.build --force > /dev/null
