"""Checks oracle/langevin_ref.py: physics anchors of BASELINE.md, closed-form gradients against
finite differences, float32 vs binary64 agreement, golden fixture currency."""
import os

import numpy as np
import pytest

from oracle import jaxlike as J, langevin_ref as L

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'langevin.npz'))
KT = 4.183
BETA = 1 / KT
XM = 14.
KAPPA = 21.3863 / (BETA * XM ** 2)


def test_force_is_minus_energy_gradient():
  pot = L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0.3, 1.5, BETA)
  x = np.linspace(-5, 35, 41)
  e = 1e-6
  fd = -(L.energy(pot, x + e, 3.0, np.float64) - L.energy(pot, x - e, 3.0, np.float64)) / (2 * e)
  np.testing.assert_allclose(L.force(pot, x, 3.0, np.float64), fd, rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(L.force(pot, x, 3.0), fd, rtol=2e-5, atol=2e-4)


def test_barrier_height_anchor():
  # kappa * beta * x_m^2 = 21.3863  =>  10 kT barrier (make_animation.py:183-184)
  pot = L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 0.0, BETA)
  e_top = L.energy(pot, np.array([XM]), 0., np.float64)[0]
  e_well = L.energy(pot, np.array([0.]), 0., np.float64)[0]
  np.testing.assert_allclose((e_top - e_well) / KT, 10.0, atol=0.01)


def test_linear_protocol_mean_work_anchor():
  # harmonic trap dragged 5 -> 10 in tau = 1 (k = gamma = kT = 1): <W> = 25/e for the linear protocol
  S, B = 400, 4000
  coeffs = L.fit_linear_to_cheby(1.0, 1.0 / S, 5, 10, 12)
  out = L.estimate_gradient(L.V_spring(1.0), coeffs, np.full(B, 5.0, np.float32), J.PRNGKey(0), 5., 10., 50, S,
                            1.0 / S, 1.0, 1.0, 1.0)
  w = out['total_work'].astype(np.float64)
  assert abs(w.mean() - 25 / np.e) < 4 * w.std() / np.sqrt(B) + 0.08   # + O(dt) discretisation


@pytest.mark.parametrize('kind', ['spring', 'shifted', 'symmetric'])
def test_closed_form_gradient_matches_finite_differences(kind):
  pot = dict(spring=L.V_spring(1.3), shifted=L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA),
             symmetric=L.V_biomolecule(KAPPA, 0.7 * KAPPA, XM / 2, 0.5, 1.5, BETA))[kind]
  B, S = 48, 40
  coeffs = np.array([20., 19., 0.5, -0.3, 0.2, 0.1])
  _, rng0 = L.trajectory_keys(J.PRNGKey(3), B)
  x0 = np.zeros(B)
  args = (0., 40., 3, S, 1e-8, KT, 1.0, KT / 1e7)
  cf = L.estimate_gradient(pot, coeffs, x0, None, *args, dtype=np.float64, rng0=rng0)
  fd = L.fd_gradient(pot, coeffs, x0, rng0, *args)
  for a, b in (('grad', 'reinforce'), ('grad_logP', 'logP'), ('grad_reparam', 'reparam')):
    np.testing.assert_allclose(cf[a], fd[b], rtol=1e-6, atol=1e-6 * np.abs(fd[b]).max())


def test_float32_oracle_tracks_binary64_truth():
  pot = L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
  B, S = 64, 500
  coeffs = GOLD['shifted_coeffs']
  x0 = np.zeros(B, np.float32)
  args = (0., 40., 0, S, 1e-8, KT, 1.0, KT / 1e7)
  o32 = L.estimate_gradient(pot, coeffs, x0, J.PRNGKey(1), *args, record=True)
  o64 = L.estimate_gradient(pot, coeffs, x0, J.PRNGKey(1), *args, dtype=np.float64, record=True)
  scale = np.abs(o64['traj']['positions']).max()
  assert np.abs(o32['traj']['positions'] - o64['traj']['positions']).max() / scale < 2e-5
  assert np.abs(o32['total_work'] - o64['total_work']).max() / np.abs(o64['total_work']).max() < 2e-5
  np.testing.assert_allclose(o32['grad'], o64['grad'], rtol=1e-4, atol=1e-5 * np.abs(o64['grad']).max())


def test_key_chain_layout():
  # seeds = split(seed, B); per trajectory `key, split = split(key)`; state rng = split
  seeds, rng0 = L.trajectory_keys(J.PRNGKey(0), 5)
  np.testing.assert_array_equal(seeds, J.split(J.PRNGKey(0), 5))
  for n in range(5):
    np.testing.assert_array_equal(rng0[n], J.split(seeds[n])[1])


def test_golden_fixture_is_current():
  pot = L.V_spring(1.0)
  coeffs = GOLD['spring_coeffs']
  o = L.estimate_gradient(pot, coeffs, np.full(24, 5.0, np.float32), J.PRNGKey(0), 5., 10., 7, 96, 1 / 96, 1.0,
                          1.0, 1.0, record=True)
  np.testing.assert_array_equal(o['traj']['positions'], GOLD['spring_positions'])
  np.testing.assert_array_equal(o['total_work'], GOLD['spring_total_work'])
  np.testing.assert_allclose(o['grad_logP'], GOLD['spring_grad_logP'], rtol=1e-12)
