import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_langevin.npz'))
c = {k.split('/')[1]: g[k] for k in g.files if k.startswith('shifted_long/')}
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
_, shift = space.free()
B, S = c['positions'].shape[0], int(c['S'])
rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  x0 = np.full((B, 1, 1), c['x0'], np.float32)
  pos, lps, works = bc.run_brownian_opt(pot, c['coeffs'], x0, 0., 40., 0, shift, nqr.split(c['seed'], B), S, float(c['dt']), float(c['kT']), 1.0, float(c['gamma']))
  est = bc.estimate_gradient_boltzinit(B, pot, 0., 40., 0, shift, S, float(c['dt']), float(c['kT']), 1.0, float(c['gamma']), record_positions=False)
  grad, (e, (_, lp, w)) = est(c['coeffs'], x0, c['seed'])
  print(mode, 'pos', rel(pos.cpu().numpy(), c['positions']), 'W', rel(w.cpu().numpy(), c['total_work']), 'LP', rel(lp.cpu().numpy(), c['tot_log_prob']), 'grad', rel(grad.cpu().numpy(), c['grad']))
