"""Generates the committed fixtures under tests/golden/.

The reference itself cannot run in the build image (jax & co. are not installable), so these
vectors come from the CPU oracle (oracle/), which is pinned to the RNG known answers and to the
reference's own exact tests (see oracle/__init__.py).  They freeze the oracle: a later change to
the oracle or to the kernels that moves any of these bits shows up as a test failure.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ising_ref as I, jaxlike as J, langevin_ref as L, parameterization_ref as P  # noqa: E402


def rng():
  key = J.PRNGKey(20240607)
  g = dict(key=[int(v) for v in key],
           split7=J.split(key, 7).tolist(),
           bits9=J.random_bits(key, (9,)).tolist(),
           uniform9_bits=J.uniform(key, (9,)).view(np.uint32).tolist(),
           normal9_bits=J.normal(key, (9,)).view(np.uint32).tolist())
  json.dump(g, open(os.path.join(HERE, 'rng.json'), 'w'), indent=1)


def langevin():
  out = {}
  kT = 4.183
  beta = 1 / kT
  xm = 14.
  kap = 21.3863 / (beta * xm ** 2)
  cases = dict(
      spring=dict(pot=L.V_spring(1.0), coeffs=L.fit_linear_to_cheby(1.0, 1 / 96, 5, 10, 6), x0=5.0, r0=(5., 10.),
                  Neq=7, S=96, B=24, dt=1 / 96, kT=1.0, mass=1.0, gamma=1.0, seed=0),
      shifted=dict(pot=L.V_biomolecule_shifted(kap, kap, xm, 0., 1.5, beta),
                   coeffs=np.array([20., 19.96, 0.4, -0.3, 0.2], np.float32), x0=0.0, r0=(0., 40.),
                   Neq=0, S=160, B=24, dt=1e-8, kT=kT, mass=1.0, gamma=kT / 1e7, seed=7),
      symmetric=dict(pot=L.V_biomolecule(kap, kap * 0.8, xm / 2, 1.0, 1.5, beta),
                     coeffs=np.array([0., 7., 0.5], np.float32), x0=-7.0, r0=(-7., 7.),
                     Neq=3, S=130, B=24, dt=1e-8, kT=kT, mass=1.0, gamma=kT / 1e7, seed=3),
  )
  for name, c in cases.items():
    x0 = np.full(c['B'], c['x0'], np.float32)
    o = L.estimate_gradient(c['pot'], c['coeffs'], x0, J.PRNGKey(c['seed']), c['r0'][0], c['r0'][1], c['Neq'],
                            c['S'], c['dt'], c['kT'], c['mass'], c['gamma'], record=True)
    out[name + '_coeffs'] = np.asarray(c['coeffs'], np.float32)
    out[name + '_positions'] = o['traj']['positions']
    out[name + '_log_probs'] = o['traj']['log_probs']
    out[name + '_works'] = o['traj']['works']
    out[name + '_total_work'] = o['total_work']
    out[name + '_tot_log_prob'] = o['tot_log_prob']
    out[name + '_grad_logP'] = o['grad_logP']
    out[name + '_grad_reparam'] = o['grad_reparam']
    out[name + '_r'] = o['r']
  np.savez_compressed(os.path.join(HERE, 'langevin.npz'), **out)


def ising_schedule(T, D, seed, scale=0.1):
  w = (scale * np.random.RandomState(seed).normal(size=(2, D))).astype(np.float32)
  lt = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[0]), 0., 0.), P.log_temp_baseline())
  h = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[1]), 0., 0.), P.field_baseline())
  times = np.linspace(0, 1, T, dtype=np.float32)
  return w, times, lt(times), h(times), lt.jacobian(times), h.jacobian(times)


def ising():
  out = {}
  for name, Lside, T, R, init in (('L64', 64, 6, 3, 'down'), ('L10', 10, 5, 3, 'down'), ('L128', 128, 4, 2, 'random')):
    w, times, lt, h, Jl, Jh = ising_schedule(T, 8, 0)
    if init == 'down':
      s0 = -np.ones((Lside, Lside), np.int32)
    else:
      s0 = I.random_spins((Lside, Lside), 0.5, J.PRNGKey(1))
    seeds = J.split(J.PRNGKey(0), R)
    finals, summaries, states = [], [], []
    grads = {k: [] for k in ('work', 'entropy')}
    for r in range(R):
      fin, summ, st = I.simulate_ising(lt, h, s0, seeds[r], return_states=True)
      finals.append(fin)
      summaries.append(np.stack([getattr(summ, f) for f in I.IsingSummary._fields]))
      states.append(st)
      for lname, key in (('total_work', 'work'), ('total_entropy_production', 'entropy')):
        cf = I.closed_form_gradient(lname, lt, h, Jl, Jh, s0, seeds[r])
        grads[key].append(np.concatenate([cf['grad'][0], cf['grad'][1], cf['grad_logP'][0], cf['grad_logP'][1],
                                          cf['grad_reparam'][0], cf['grad_reparam'][1], [cf['loss']]]))
    out[name + '_weights'] = w
    out[name + '_log_temp'] = lt
    out[name + '_field'] = h
    out[name + '_spins0'] = s0
    out[name + '_final'] = np.stack(finals).astype(np.int8)
    out[name + '_states'] = np.stack(states).astype(np.int8)
    out[name + '_summary'] = np.stack(summaries)          # [R, 7, T-1]
    out[name + '_grad_work'] = np.stack(grads['work'])
    out[name + '_grad_entropy'] = np.stack(grads['entropy'])
  np.savez_compressed(os.path.join(HERE, 'ising.npz'), **out)


if __name__ == '__main__':
  rng()
  langevin()
  ising()
  print('golden fixtures written to', HERE)
