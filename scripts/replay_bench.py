"""Dev: the checkpointed forward + replay ("adjoint") gradient at config 2 against the fused single pass."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
_, shift = space.free()
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
coeffs = np.array([20., 19.95996] + [0.] * 11, np.float32)
B, S = 100_000, 10_000
x0 = np.zeros((B, 1, 1), np.float32)
key = nqr.PRNGKey(0)
fused = bc.estimate_gradient_boltzinit(B, pot, 0., 40., 0, shift, S, 1e-8, KT, 1.0, KT / 1e7, record_positions=False)
for K in (50, 125, 500):
  ck = bc.estimate_gradient_boltzinit_checkpointed(B, pot, 0., 40., 0, shift, S, 1e-8, KT, 1.0, KT / 1e7, checkpoint_every=K)
  for name, fn in (('fused', fused), (f'checkpoint_every={K}', ck)):
    for rep in range(3):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); g = fn(coeffs, x0, key)[0]; e1.record(); torch.cuda.synchronize()
    print(f'{name:24s}: {e0.elapsed_time(e1):7.2f} ms  {B * S / e0.elapsed_time(e1) * 1e3:.3g} traj-steps/s  grad[0:3] {g[:3].cpu().numpy()}')
