"""The oracle against the REFERENCE'S OWN SOURCE.

tests/golden/reference_*.npz hold outputs of /root/reference/noneq_opt/{barrier_crossing,ising}.py,
imported verbatim and executed over oracle/jax_shim (torch-CPU float32 stand-ins for jax, jax_md,
tensorflow_probability and distrax; threefry streams from oracle/jaxlike.py) by
tests/golden/make_reference_golden.py.  This pins the oracle's reading of the reference -- step order,
key plumbing, work / log-prob bookkeeping, the estimator objectives and what autodiff makes of them --
to the reference's code rather than to a restatement of it.  Tolerances cover only the last bits of
exp / log / erf_inv / sigmoid and reduction order (torch kernels vs the oracle's float64-then-round).
"""
import os

import numpy as np
import pytest

from oracle import ising_ref as I, jaxlike as J, langevin_ref as L, parameterization_ref as P

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)


def case(file, name):
  g = np.load(os.path.join(GOLD, file))
  return {k.split('/')[1]: g[k] for k in g.files if k.startswith(name + '/')}


def rel(a, b):
  return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


@pytest.mark.parametrize('name', ['spring', 'shifted', 'biomolecule_periodic', 'shifted_long'])
def test_langevin_oracle_reproduces_the_reference_source(name):
  c = case('reference_langevin.npz', name)
  pot = {'spring': L.V_spring(1.0), 'shifted': L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA),
         'shifted_long': L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA),
         'biomolecule_periodic': L.V_biomolecule(KAPPA, KAPPA, XM / 2, 0.5, 1.5, BETA)}[name]
  shift = ('periodic', float(c['box'])) if float(c['box']) > 0 else 'free'
  # 2500 steps across the barrier amplify last-bit differences of exp / log between any two float32
  # implementations (here torch's kernels under the reference's code vs the oracle): 5.5e-6 measured
  tol = 1e-5 if name == 'shifted_long' else 1e-6
  B = c['positions'].shape[0]
  o = L.estimate_gradient(pot, c['coeffs'], np.full(B, c['x0'], np.float32), c['seed'], float(c['r0'][0]), float(c['r0'][1]),
                          int(c['Neq']), int(c['S']), float(c['dt']), float(c['kT']), float(c['mass']), float(c['gamma']),
                          shift=shift, record=True)
  # run_brownian_opt (barrier_crossing.py:97-143), one call per row of split(seed, B)
  assert rel(o['traj']['positions'], c['positions'][:, :, 0, 0]) < tol
  assert rel(o['traj']['log_probs'], c['log_probs']) < tol
  assert np.abs(o['traj']['works'] - c['works']).max() < 2 * tol * np.abs(c['total_work']).max()
  assert np.all(c['works'][:, 0] == 0) and c['positions'].shape[1] == int(c['S']) + 1
  if name == 'spring':   # no transcendental on the path except erf_inv: the two agree bit for bit
    np.testing.assert_array_equal(o['traj']['positions'], c['positions'][:, :, 0, 0])
    np.testing.assert_array_equal(o['traj']['works'], c['works'])
  # estimate_gradient_boltzinit (:204-224): jax.value_and_grad of sum(logp) sg(W) + W through the scan
  np.testing.assert_allclose(c['est_positions'], c['positions'])
  assert rel(o['total_work'], c['total_work']) < tol and rel(o['tot_log_prob'], c['tot_log_prob']) < tol
  assert rel(o['grad'], c['grad']) < 1e-5
  assert rel(o['tot_log_prob'] * o['total_work'] + o['total_work'], c['estimator']) < tol


@pytest.mark.parametrize('name', ['small', 'fastpath', 'chain', 'cube'])
def test_ising_oracle_reproduces_the_reference_source(name):
  c = case('reference_ising.npz', name)
  T, shape = int(c['T']), c['spins0'].shape
  times = np.linspace(0, 1, T, dtype=np.float32)
  np.testing.assert_array_equal(times, c['times'])
  # IsingSchedule(Chebyshev, Chebyshev)(times) (ising.py:19-25, parameterization.py:144-179)
  np.testing.assert_allclose(P.Chebyshev(c['w_lt'])(times), c['log_temp'], atol=1e-7)
  np.testing.assert_allclose(P.Chebyshev(c['w_h'])(times), c['field'], atol=1e-7)
  # random_spins (:71-72)
  np.testing.assert_array_equal(I.random_spins(shape, 0.5, J.PRNGKey(11)), c['spins0'])
  fields = [str(f) for f in c['summary_fields']]
  n = float(np.prod(shape))
  for r in range(c['seeds'].shape[0]):
    fin, summ, states = I.simulate_ising(c['log_temp'], c['field'], c['spins0'], c['seeds'][r], return_states=True)
    np.testing.assert_array_equal(fin, c['final_spins'][r])          # spin history bit-exact, sweep by sweep
    np.testing.assert_array_equal(states, c['states'][r])
    for i, f in enumerate(fields):
      got, want = getattr(summ, f), c['summary'][r, i]
      if f == 'magnetization':
        np.testing.assert_array_equal(got, want)
      else:   # float32 sums over L^2 sites: a few ulp of the largest partial sum
        assert np.abs(got - want).max() <= 8 * np.spacing(np.float32(max(n, np.abs(want).max()))), f
    # estimate_gradient_rev (:179-198): jax.grad of sum(fwd_lp) sg(loss) + loss wrt the schedule weights
    so = P.Chebyshev(c['w_lt']), P.Chebyshev(c['w_h'])
    for loss, key in (('total_entropy_production', 'grad_entropy'), ('total_work', 'grad_work')):
      cf = I.closed_form_gradient(loss, so[0](times), so[1](times), so[0].jacobian(times), so[1].jacobian(times),
                                  c['spins0'], c['seeds'][r])
      assert rel(cf['grad'][0], c[key][r, 0]) < 2e-5 and rel(cf['grad'][1], c[key][r, 1]) < 2e-5, (loss, r)


@pytest.mark.parametrize('name', ['spring', 'shifted', 'biomolecule_periodic'])
def test_langevin_estimator_variants_and_eqm_prepass(name):
  """estimate_gradient_boltzinit_logP / _reparam (barrier_crossing.py:262-311) and mapped_brownian_eqm
  (:249-256, which hard-codes dt = 1e-6, kT = 1e-5, mass = gamma = 1 whatever it is given)."""
  c = case('reference_langevin.npz', name)
  pot = {'spring': L.V_spring(1.0), 'shifted': L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA),
         'biomolecule_periodic': L.V_biomolecule(KAPPA, KAPPA, XM / 2, 0.5, 1.5, BETA)}[name]
  shift = ('periodic', float(c['box'])) if float(c['box']) > 0 else 'free'
  B = c['positions'].shape[0]
  o = L.estimate_gradient(pot, c['coeffs'], np.full(B, c['x0'], np.float32), c['seed'], float(c['r0'][0]), float(c['r0'][1]),
                          int(c['Neq']), int(c['S']), float(c['dt']), float(c['kT']), float(c['mass']), float(c['gamma']), shift=shift,
                          record=True)
  assert rel(o['grad_logP'], c['grad_logP']) < 1e-5 and rel(o['grad_reparam'], c['grad_reparam']) < 1e-5
  assert rel(o['grad_logP'] + o['grad_reparam'], c['grad']) < 1e-5
  _, rng0 = L.trajectory_keys(J.PRNGKey(5), B)
  xs = L.run_brownian_stationary_trap(pot, np.full(B, c['x0'], np.float32), float(c['r0'][0]), rng0, 40, 1e-6, 1e-5, 1.0, 1.0,
                                      shift)
  assert np.abs(xs - c['eqm'][:, 0, 0]).max() < 1e-6 * max(1.0, np.abs(c['eqm']).max())


def test_protocol_algebra_matches_the_reference_source():
  """parameterization.py:67-179, 268-344 and the baselines of ising.py:45-55, for the oracle and for the
  product's host-side classes."""
  from noneq_opt_b200 import ising as nq_ising, parameterization as p10n
  c = case('reference_schedules.npz', 'all')
  x, x25 = c['x'], c['x25']
  for mod, lt_base, h_base in ((P, P.log_temp_baseline(0.69, 10., 1), P.field_baseline(-1., 1.)),
                               (p10n, nq_ising.log_temp_baseline(0.69, 10., 1), nq_ising.field_baseline(-1., 1.))):
    cheb = mod.Chebyshev(c['w'])
    np.testing.assert_allclose(cheb(x), c['chebyshev'], atol=2e-7)
    np.testing.assert_allclose(mod.PiecewiseLinear(c['knots'], 0., 1., -1., 2.)(x), c['piecewise'], atol=3e-7)
    np.testing.assert_allclose(lt_base(x), c['log_temp_baseline'], atol=3e-7)
    np.testing.assert_allclose(h_base(x), c['field_baseline'], atol=1e-7)
    np.testing.assert_allclose(mod.AddBaseline(mod.ConstrainEndpoints(cheb, 0., 0.), lt_base)(x), c['composed'], atol=3e-7)
    np.testing.assert_allclose(mod.ChangeDomain(cheb, 2., 5.)(x25), c['change_domain'], atol=2e-7)
    np.testing.assert_allclose(mod.Constant(np.float32(0.7))(x), c['constant'], atol=0)


# ---- reference_extra.npz: rank-generic Brownian states and a 200-sweep Ising run (make_reference_golden.py)
def _extra_pot(name):
  return {'spring_7x3': L.V_spring(2.0), 'shifted_4x1': L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA),
          'spring_periodic_5x2': L.V_spring(1.0)}[name]


@pytest.mark.parametrize('name', ['spring_7x3', 'shifted_4x1', 'spring_periodic_5x2'])
def test_multi_particle_apply_reproduces_the_reference_source(name):
  """simulate.brownian (simulate.py:33-99) on [N, dim] states: one key per state, jax.random.normal(split, [1, N, dim]),
  log_prob summed over every element; the energy closures read position[0][0] only."""
  c = case('reference_extra.npz', name)
  sc = L.step_constants(float(c['dt']), float(c['kT']), 1.0, float(c['gamma']))
  shift = ('periodic', float(c['box'])) if float(c['box']) > 0 else 'free'
  x, key = c['R0'], J.PRNGKey(9)
  for k in range(4):
    x, key, lp, _ = L.brownian_apply_nd(_extra_pot(name), sc, x, key, float(c['r0']), shift)
    np.testing.assert_array_equal(key, c['keys'][k])
    assert rel(x, c['positions'][k]) < 1e-6
    assert abs(float(lp) - float(c['log_probs'][k])) <= 2e-6 * abs(float(c['log_probs'][k]))


def test_ising_oracle_reproduces_the_reference_source_over_200_sweeps():
  """T = 201 under the drivers' schedule composition: spin history bit-exact (5 intermediate lattices + the last),
  summaries to a few ulp, both loss gradients -- numpy oracle and its C form."""
  from oracle import c_oracle
  c = case('reference_extra.npz', 'long')
  L_, T = int(c['L']), int(c['T'])
  times = np.linspace(0, 1, T, dtype=np.float32)
  lt = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(c['w'][0]), 0., 0.), P.log_temp_baseline())(times)
  h = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(c['w'][1]), 0., 0.), P.field_baseline())(times)
  np.testing.assert_allclose(lt, c['log_temp'], atol=2e-7)
  np.testing.assert_allclose(h, c['field'], atol=2e-7)
  s0 = -np.ones((L_, L_), np.int32)
  sweeps = [int(v) for v in c['state_sweeps']]
  fin_c, summ_c, snaps_c = c_oracle.ising(c['log_temp'], c['field'], s0.astype(np.int8), c['seeds'], snapshots=sweeps)
  np.testing.assert_array_equal(fin_c, c['final_spins'])
  np.testing.assert_array_equal(snaps_c, c['states'])
  jac = [P.ConstrainEndpoints(P.Chebyshev(np.asarray(c['w'][i], np.float64), np.float64), 0., 0.).jacobian(times.astype(np.float64))
         for i in (0, 1)]
  for r in range(2):
    fin, summ, states = I.simulate_ising(c['log_temp'], c['field'], s0, c['seeds'][r], return_states=True)
    np.testing.assert_array_equal(fin, c['final_spins'][r])
    np.testing.assert_array_equal(states[[s - 1 for s in sweeps]], c['states'][r])
    for i, f in enumerate(I.IsingSummary._fields):
      want = c['summary'][r, i]
      tol = 0 if f == 'magnetization' else 8 * np.spacing(np.float32(max(L_ * L_, np.abs(want).max())))
      assert np.abs(getattr(summ, f) - want).max() <= tol, f
      assert np.abs(summ_c[r, i] - want).max() <= tol, f
    for loss, key in (('total_entropy_production', 'grad_entropy'), ('total_work', 'grad_work')):
      o = I.closed_form_gradient(loss, c['log_temp'], c['field'], jac[0], jac[1], s0, c['seeds'][r])
      assert rel(o['grad'][0], c[key][r, 0]) < 1e-4 and rel(o['grad'][1], c[key][r, 1]) < 1e-4
