"""Dev: time the fused Langevin kernel at trajectory counts that put exactly w warps on every SM
sub-partition (w = 1..6): separates per-warp latency from issue throughput."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
dev = torch.device('cuda')
_, shift = space.free()
pot = bc.V_biomolecule_shifted(bench.KAPPA, bench.KAPPA, bench.X_M, 0., bench.K_TRAP, bench.BETA)
S = 10000
r_np, phi_np = bc._schedule_and_jacobian(np.arange(S + 1), bench.C2_COEFFS, 0., 40.)
r_dev, phi_dev = torch.from_numpy(r_np).to(dev), torch.from_numpy(np.array(phi_np)).to(dev)
for mode in ('fast', 'strict'):
  bc.set_math_mode(mode)
  for w in (1, 2, 3, 4, 5, 6):
    B = 592 * 32 * w
    seeds = nqr.split(nqr.PRNGKey(0), B)
    x0 = torch.zeros(B, device=dev)
    for rep in range(3):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      bc._simulate(pot, shift, r_dev, phi_dev, x0, seeds, S, 0, bench.DT, bench.KT, 1.0, bench.GAMMA)
      e1.record()
      torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'{mode} warps/SMSP={w} B={B} {ms:.3f} ms  {B * S / ms / 1e3:.4g} traj-steps/s  cycles/warp-step={ms * 1e-3 * 1.965e9 / S / w:.0f}')
