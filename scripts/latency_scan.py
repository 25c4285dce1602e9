"""Dev: Langevin gradient kernel time vs ensemble size (latency-bound regime for small ensembles)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
_, shift = space.free()
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
pots = {'shifted': (bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA), np.array([20., 19.95996] + [0.] * 11, np.float32), 0., 40., 1e-8, KT, KT / 1e7),
        'spring': (bc.V_spring(1.0), bc.fit_linear_to_cheby(1.0, 1e-3, 5., 10., 12), 5., 10., 1e-3, 1.0, 1.0)}
S = 2000
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  for name, (pot, coeffs, r0, r1, dt, kT, gamma) in pots.items():
    for B in (1000, 4736, 9472, 12288, 16384, 24000):
      seeds = nqr.split(nqr.PRNGKey(0), B)
      x0 = torch.full((B,), r0, device='cuda')
      res = {}
      for form, pair_max in (('pipeline', '65536'), ('one-warp', '0')):   # three-warp pipeline vs langevin_kernel
        os.environ['NQ_PAIR_MAX'] = pair_max
        for rep in range(3):
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record()
          bc.gradient_terms(pot, coeffs, x0, seeds, r0, r1, 0, shift, S, dt, kT, 1.0, gamma)
          e1.record(); torch.cuda.synchronize()
        res[form] = e0.elapsed_time(e1)
      del os.environ['NQ_PAIR_MAX']
      print(f'{mode:6s} {name:8s} B={B:6d}: ' + '  '.join(
          f'{form} {ms:7.3f} ms {ms * 1e-3 * 1.965e9 / S:5.0f} cycles/step' for form, ms in res.items()))
