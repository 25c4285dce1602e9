#!/bin/bash
# Dev helper: build kernel variants ON the GPU box and bench config 2 (STRICT, then FAST) for each.
# usage: scripts/variant_strict.sh "<nvcc flags variant 1>" "<flags variant 2>" ...   (env NQ_* are passed through)
for flags in "$@"; do
  NQ_EXTRA_NVCC_FLAGS="$flags" python -m noneq_opt_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  for m in strict fast; do
    NQ_EXTRA_NVCC_FLAGS="$flags" python bench.py --steps 8 --warmup 3 --math $m --no-ising --no-cpu --no-extras 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$flags |', d['math_mode'], '| %.4g traj-steps/s | %.3f ms' % (d['value'], d['ms_per_step']))"
  done
done
