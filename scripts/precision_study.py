"""Dev script: deviation of the built library from the C oracle on a full-length slice of config 2
(n trajectories x 10^4 steps), for one math mode.  usage: precision_study.py <strict|fast> [n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space  # noqa: E402
from oracle import c_oracle, jaxlike as J, langevin_ref as L  # noqa: E402

mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
S, B_total = 10_000, 100_000
_, shift = space.free()
coeffs = np.array([2.0e+01, 1.99599600e+01] + [0.] * 11, np.float32)
pot_o = L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
pot_g = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
r, phi = L.make_schedule_chebyshev(S, coeffs, 0., 40.)
seeds_np = J.split(J.PRNGKey(0), B_total)[:n]
o = c_oracle.langevin(pot_o, r, phi, np.zeros(n, np.float32), seeds_np, S, 0, 1e-8, KT, KT / 1e7, 1.0, record=True)
bc.set_math_mode(mode)
seeds = nqr.split(J.PRNGKey(0), B_total, 0, n)
g = bc.gradient_terms(pot_g, coeffs, np.zeros(n, np.float32), seeds, 0., 40., 0, shift, S, 1e-8, KT, 1.0, KT / 1e7)
w, lp, xf = g['work'].cpu().numpy(), g['logp'].cpu().numpy(), g['x_final'].cpu().numpy()
werr = np.abs(w - o['work']) / np.abs(o['work']).max()
xerr = np.abs(xf - o['positions'][:, -1]) / np.abs(o['positions']).max()
lerr = np.abs(lp - o['logp']) / np.abs(o['logp']).max()
gs = g['grad_sums'].cpu().numpy()
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
q = lambda e: 'max %.2e q99.9 %.2e q99 %.2e med %.2e' % (e.max(), np.quantile(e, .999), np.quantile(e, .99), np.median(e))
print(f'[{mode} n={n}] x_final: {q(xerr)}')
print(f'[{mode} n={n}] work   : {q(werr)}   mean-ratio-1 {abs(w.astype(np.float64).mean() / o["work"].astype(np.float64).mean() - 1):.2e}')
print(f'[{mode} n={n}] logp   : {q(lerr)}')
print(f'[{mode} n={n}] grads  : logP {rel(gs[0], o["grad_sums"][0]):.2e} reparam {rel(gs[1], o["grad_sums"][1]):.2e}')

# work against the EXACT work along the oracle's own binary32 trajectory (binary64 over the recorded positions): how far
# the reference-order evaluation (difference of two O(100) energies per step) and the kernel's closed form sit from it
same = xf == o['positions'][:, -1]
pos = o['positions'].astype(np.float64)
rr = np.asarray(r, np.float64)
exact = (0.75 * ((pos[:, :-1] - rr[None, 1:]) ** 2 - (pos[:, :-1] - rr[None, :-1]) ** 2)).sum(1)
ek = np.abs(w - exact)[same] / np.abs(exact).max()
eo = np.abs(o['work'] - exact)[same] / np.abs(exact).max()
print(f'[{mode} n={n}] {same.mean() * 100:.2f} % of the trajectories end on the oracle\'s bits; on those, work vs exact: '
      f'kernel med {np.median(ek):.2e} max {ek.max():.2e} | reference-order oracle med {np.median(eo):.2e} max {eo.max():.2e}')
