#!/bin/bash
# The pool's compute-sanitizer is closed ("runs under it have left GPUs needing a reset"), so the race / bounds
# evidence for the hand-rolled synchronisation comes from device asserts compiled in with -DNQ_DEBUG_CHECKS:
#   * pair kernel: every ring slot carries the global step index of the draw it holds; the consumer asserts it
#   * persistent scheduler: every carry ends with a checksum of (position, key chain, next step); the restoring
#     chunk asserts it (a chunk that ran ahead of its predecessor's hand-over would read a stale or torn carry)
#   * Ising segment scheduler: the spin sum handed over must equal a recount of the lattice handed over
#   * staged-tile bounds
# Build the checked library ON the GPU box, run the kernel smoke script and the scheduler-heavy tests, restore.
#   usage (gpurun): scripts/debug_checks.sh > gpurun_out/r02_debug_checks.log 2>&1
set -u
NQ_EXTRA_NVCC_FLAGS="-DNQ_DEBUG_CHECKS" python -m noneq_opt_b200.build --force > /dev/null 2>&1 || { echo "checked build failed"; exit 1; }
echo "== checked build (-DNQ_DEBUG_CHECKS) ok"
export NQ_EXTRA_NVCC_FLAGS="-DNQ_DEBUG_CHECKS"   # keep build.py's digest stable for the runs below
NQ_WAIT_POLLS=0 python scripts/sanitize_smoke.py; echo "== sanitize_smoke rc=$?"
python -m pytest -q -m gpu tests/test_gpu_langevin.py -k "persistent or pair or config5 or full_length or every_coefficient" 2>&1 | tail -3
python -m pytest -q -m gpu tests/test_gpu_ising.py tests/test_gpu_ising_config_length.py -k "segment or config4" 2>&1 | tail -3
python -m pytest -q -m gpu tests/test_gpu_training.py 2>&1 | tail -3
unset NQ_EXTRA_NVCC_FLAGS
python -m noneq_opt_b200.build --force > /dev/null 2>&1 && echo "== product build restored"
