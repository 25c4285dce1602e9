#!/bin/bash
# Dev helper: build kernel variants ON the GPU box, report FAST-mode deviation from the C oracle and speed.
for flags in "$@"; do
  NQ_EXTRA_NVCC_FLAGS="$flags" python -m noneq_opt_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "=== $flags"
  timeout 300 python scripts/precision_study.py fast 8192 2>&1 | grep '^\['
  python bench.py --steps 10 --warmup 3 --math fast --no-ising --no-cpu 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   %.4g traj-steps/s | %.3f ms' % (d['value'], d['ms_per_step']))"
done
