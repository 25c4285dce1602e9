cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3
python scripts/ising_profile.py c3_noseg > gpurun_out/r02zc_ising_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ising_fast_kernel -s 1 -c 1 -f -o gpurun_out/r02zc_ising_c3 python scripts/ising_profile.py c3_noseg > gpurun_out/r02zc_ising_ncu.log 2>&1
cat gpurun_out/r02zc_ising_plain.log
