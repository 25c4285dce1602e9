"""Pins oracle/ising_ref.py to the reference's own exact and property tests
(/root/reference/tests/ising_test.py) and cross-checks the closed-form gradients."""
import os

import numpy as np
import pytest

from oracle import ising_ref as I, jaxlike as J, parameterization_ref as P

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ising.npz'))


# ---- exact known answers: tests/ising_test.py:15-66
@pytest.mark.parametrize('inp, expected', [
    (np.array([0, 1, 2]), np.array([3, 2, 1])),
    (np.arange(9).reshape(3, 3), np.array([[12, 13, 14], [15, 16, 17], [18, 19, 20]])),
])
def test_sum_neighbors(inp, expected):
  np.testing.assert_array_equal(I.sum_neighbors(inp), expected)


@pytest.mark.parametrize('spins, log_temp, field, expected', [
    (np.array([-1, 1, 1, 1, -1]), 0., 1., np.array([2., -2., -6., -2., 2.])),
    (np.array([-1, 1, 1, 1, -1]), 0., -1., np.array([-2., 2., -2., 2., -2.])),
    (np.array([-1, 1, 1, 1, -1]), np.log(2.), 1., np.array([1., -1., -3., -1., 1.])),
    (np.array([-1, 1, 1]), 100., -1., np.array([0., 0., 0.])),
])
def test_flip_logits(spins, log_temp, field, expected):
  np.testing.assert_allclose(expected, I.flip_logits(spins, log_temp, field), rtol=1e-7, atol=1e-30)


@pytest.mark.parametrize('shape, expected_even, expected_odd', [
    ((2, 2), np.array([[0, 1], [1, 0]]), np.array([[1, 0], [0, 1]])),
    ((4, 2), np.array([[0, 1], [1, 0], [0, 1], [1, 0]]), np.array([[1, 0], [0, 1], [1, 0], [0, 1]])),
])
def test_even_odd_mask(shape, expected_even, expected_odd):
  even, odd = I.even_odd_masks(shape)
  np.testing.assert_array_equal(expected_even, even)
  np.testing.assert_array_equal(expected_odd, odd)


def test_odd_shape_raises():
  with pytest.raises(ValueError):
    I.even_odd_masks((3, 4))


# ---- property tests: tests/ising_test.py:71-127
@pytest.mark.parametrize('shape', [(10000,), (64, 64), (16, 16, 16)])
@pytest.mark.parametrize('field', [-1.0, 0.0, 0.5])
@pytest.mark.parametrize('temperature', [float(np.exp(-1.)), 1.0, float(np.exp(1.))])
def test_detailed_balance_and_entropy_production(shape, field, temperature):
  init_seed, sim_seed = J.split(J.PRNGKey(0))
  lt = np.log(np.float32(temperature))
  s0 = I.random_spins(shape, .5, init_seed)
  lp0 = -I.energy(s0, field) / np.float32(temperature)
  s1, summ = I.update(s0, (lt, field), (lt, field), sim_seed)
  lp1 = -I.energy(s1, field) / np.float32(temperature)
  np.testing.assert_allclose(lp0 + summ.forward_log_prob, lp1 + summ.reverse_log_prob, rtol=5e-4)
  np.testing.assert_allclose(summ.dissipated_heat, summ.entropy_production * np.float32(temperature), rtol=1e-4,
                             atol=1e-3)


@pytest.mark.parametrize('shape, field, temperature, expected', [
    ((1000,), 0., 1., .5), ((32, 32), 3., .2, 1.), ((8, 8, 8), -4., .5, 0.)])
def test_statistics(shape, field, temperature, expected):
  init_seed, sim_seed = J.split(J.PRNGKey(0), 2)
  s0 = I.random_spins(shape, .5, init_seed)
  T = 200
  lt = np.full(T, np.log(temperature), np.float32)
  h = np.full(T, field, np.float32)
  s, _ = I.simulate_ising(lt, h, s0, sim_seed)
  np.testing.assert_allclose(((s + 1) / 2).mean(), expected, atol=.06)


# ---- gradients: closed forms vs brute-force differentiation of the frozen-flip surrogate
def _schedule(wl, wh, times):
  d = wl.dtype.type
  lt = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(wl, d), 0., 0.), P.log_temp_baseline(dtype=d))
  h = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(wh, d), 0., 0.), P.field_baseline(dtype=d))
  return lt(times.astype(d)), h(times.astype(d))


@pytest.mark.parametrize('loss', ['total_work', 'total_entropy_production'])
def test_closed_form_gradient_matches_finite_differences(loss):
  w = 0.1 * np.random.RandomState(0).normal(size=(2, 8))
  times = np.linspace(0, 1, 6)
  s0 = -np.ones((10, 10), np.int32)
  t32 = times.astype(np.float32)
  lt = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[0].astype(np.float32)), 0., 0.), P.log_temp_baseline())
  h = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[1].astype(np.float32)), 0., 0.), P.field_baseline())
  cf = I.closed_form_gradient(loss, lt(t32), h(t32), lt.jacobian(t32), h.jacobian(t32), s0, J.PRNGKey(5))
  fd = I.fd_gradient(loss, _schedule, w[0], w[1], times, s0, J.PRNGKey(5))
  for name, key in (('grad', 'reinforce'), ('grad_logP', 'logP'), ('grad_reparam', 'reparam')):
    for k in range(2):
      scale = max(np.abs(fd[key][k]).max(), 1.0)
      np.testing.assert_allclose(cf[name][k], fd[key][k], atol=1e-5 * scale)


def test_golden_fixture_is_current():
  """The committed vectors are what the oracle produces today."""
  lt, h, s0 = GOLD['L64_log_temp'], GOLD['L64_field'], GOLD['L64_spins0']
  seeds = J.split(J.PRNGKey(0), 3)
  fin, summ = I.simulate_ising(lt, h, s0, seeds[1])
  np.testing.assert_array_equal(fin, GOLD['L64_final'][1])
  np.testing.assert_array_equal(np.stack([getattr(summ, f) for f in I.IsingSummary._fields]), GOLD['L64_summary'][1])
