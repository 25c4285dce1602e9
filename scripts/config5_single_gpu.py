"""Dev: configs[4] on ONE GPU (10^7 trajectories x 10^4 steps, forward + gradient): throughput with many waves."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
_, shift = space.free()
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
coeffs = np.array([20., 19.95996] + [0.] * 11, np.float32)
B, S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 10_000
seeds = nqr.split(nqr.PRNGKey(0), B)
x0 = torch.zeros(B, device='cuda')
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  for rep in range(int(os.environ.get("NQ_REPS", "2"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g = bc.gradient_terms(pot, coeffs, x0, seeds, 0., 40., 0, shift, S, 1e-8, KT, 1.0, KT / 1e7)
    e1.record(); torch.cuda.synchronize()
    if os.environ.get('NQ_REPS'):
      print(f'  rep {rep}: {e0.elapsed_time(e1):.2f} ms')
  ms = e0.elapsed_time(e1)
  print(f'{mode}: B={B} S={S}: {ms:.1f} ms  {B * S / ms / 1e-3:.4g} traj-steps/s  mean work {float(g["moments"][1] / g["moments"][0]):.4f}')
