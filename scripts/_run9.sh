cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_ising.py tests/test_gpu_ising_config_length.py tests/test_gpu_reference.py tests/test_gpu_training.py -q 2>&1 | tail -3
echo "== key batch 32 (default)"; python scripts/ising_bench.py 64,4096,1001 128,2048,201 1024,256,101
echo "== key batch 0"; NQ_B200_LIB=$PWD/noneq_opt_b200/variants/kb0.so python scripts/ising_bench.py 64,4096,1001 128,2048,201
NQ_REPS=4 python scripts/config5_single_gpu.py 10000000 2>&1 | tail -12
