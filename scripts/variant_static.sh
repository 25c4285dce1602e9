#!/bin/bash
# Dev helper: build kernel variants ON the GPU box and time the STATIC Langevin launch at the config-5 shard size
# (1.25e6 trajectories x 10^4 steps, STRICT then FAST).
for flags in "$@"; do
  NQ_EXTRA_NVCC_FLAGS="$flags" python -m noneq_opt_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  NQ_EXTRA_NVCC_FLAGS="$flags" python - "$flags" <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
_, shift = space.free()
pot = bc.V_biomolecule_shifted(bench.KAPPA, bench.KAPPA, bench.X_M, 0., bench.K_TRAP, bench.BETA)
n, S = 1_250_000, bench.C2_S
seeds = nqr.split(nqr.PRNGKey(0), 10_000_000, first=0, count=n)
x0 = torch.zeros(n, device='cuda')
r, phi = bc._schedule_and_jacobian(np.arange(S + 1), bench.C2_COEFFS, 0., 40.)
r, phi = torch.from_numpy(r).cuda(), torch.from_numpy(np.array(phi)).cuda()
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = bc._simulate(pot, shift, r, phi, x0, seeds, S, 0, bench.DT, bench.KT, 1.0, bench.GAMMA); e1.record()
    torch.cuda.synchronize()
  print(sys.argv[1], '|', mode, '| %.4g traj-steps/s | %.1f ms' % (n * S / (e0.elapsed_time(e1) * 1e-3), e0.elapsed_time(e1)))
PY
done
