cd $GRAFT_REPO_ROOT
python scripts/full_configs.py > gpurun_out/r02z_full_configs.log 2>&1; cp gpurun_out/full_configs.json gpurun_out/r02z_full_configs.json; tail -5 gpurun_out/r02z_full_configs.log
python scripts/training_soak.py > gpurun_out/r02z_soak.log 2>&1; tail -4 gpurun_out/r02z_soak.log
bash scripts/debug_checks.sh > gpurun_out/r02z_debug_checks.log 2>&1; cat gpurun_out/r02z_debug_checks.log
