cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -m pytest tests/test_gpu_langevin.py -q -s -k "full_length or closed_form" 2>&1 | grep -E "full-length|work vs exact|passed|failed"
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; tail -c 400 gpurun_out/r02z_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_reference.json 2>/dev/null
NQ_B200_LIB=$PWD/noneq_opt_b200/variants/t_u2.so python bench.py --steps 8 --warmup 3 --math strict --no-ising --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unroll2 | %.4g traj-steps/s | %.3f ms' % (d['value'], d['ms_per_step']))"
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02z_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02z_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:langevin_persistent_kernel -s 2 -c 1 -f -o gpurun_out/r02z_langevin_strict python bench.py --steps 2 --warmup 1 --math strict --no-ising --no-cpu --no-extras > gpurun_out/r02z_ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:langevin_persistent_kernel -s 2 -c 1 -f -o gpurun_out/r02z_langevin_fast python bench.py --steps 2 --warmup 1 --math fast --no-ising --no-cpu --no-extras > gpurun_out/r02z_ncu_f.log 2>&1
ls -la gpurun_out/ | tail -12
