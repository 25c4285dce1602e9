#!/bin/bash
# Dev helper: build Ising kernel variants ON the GPU box and time configs 3-like shapes.
for flags in "$@"; do
  NQ_EXTRA_NVCC_FLAGS="$flags" python -m noneq_opt_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "=== $flags"
  timeout 200 python scripts/ising_bench.py 64,4096,1001 128,2048,201 1024,256,101 2>&1 | tail -3
done
