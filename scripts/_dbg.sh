cd $GRAFT_REPO_ROOT
NQ_BENCH_DEBUG=1 python bench.py --steps 24 --warmup 3 --no-ising --no-cpu --no-config5 --no-extras 2>&1 >/dev/null | grep step_resident | awk '{printf "%s ", $5} END {print ""}'
NQ_REPS=16 python scripts/config5_single_gpu.py 100000 2>&1 | grep -E "rep|strict|fast" | awk '{printf "%s ", $3} END {print ""}'
