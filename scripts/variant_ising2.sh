#!/bin/bash
# Dev helper: build Ising kernel variants ON the GPU box and time configs 3 / 4 for each (scripts/ising_profile.py).
for flags in "$@"; do
  NQ_EXTRA_NVCC_FLAGS="$flags" python -m noneq_opt_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "== $flags"
  NQ_EXTRA_NVCC_FLAGS="$flags" python scripts/ising_profile.py c3 | tail -1
  NQ_EXTRA_NVCC_FLAGS="$flags" python scripts/ising_profile.py c4 | tail -1
done
