cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_math.py -x -q 2>&1 | tail -2
b() { python bench.py --steps 8 --warmup 3 --math $2 --no-ising --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2 | %.4g traj-steps/s | %.3f ms' % (d['value'], d['ms_per_step']))"; }
for v in algw_p algw_pl2 base_pl2; do
  export NQ_B200_LIB=$PWD/noneq_opt_b200/variants/$v.so
  b $v strict
  b $v fast
done
export NQ_B200_LIB=$PWD/noneq_opt_b200/variants/algw_pl2.so
python scripts/precision_study.py strict 65536 2>&1 | grep '^\['
NQ_PERSISTENT_PER_SM=6 b "algw_pl2 per_sm=6" strict
NQ_PERSISTENT_CHUNK=256 b "algw_pl2 chunk256" strict
NQ_PERSISTENT=0 b "algw_pl2 static" strict
unset NQ_B200_LIB
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
