"""The C form of the oracle (timed CPU arm of bench.py) agrees with the numpy oracle."""
import os

import numpy as np

from oracle import c_oracle, ising_ref as I, jaxlike as J, langevin_ref as L

GOLD_L = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'langevin.npz'))
GOLD_I = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ising.npz'))
KT = 4.183
BETA = 1 / KT
XM = 14.
KAPPA = 21.3863 / (BETA * XM ** 2)


def test_builds_and_reports_threads():
  assert os.path.exists(c_oracle.build())
  assert c_oracle.max_threads() >= 1


def test_langevin_matches_numpy_oracle():
  pot = L.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
  coeffs = GOLD_L['shifted_coeffs']
  B, S = 24, 160
  r, phi = L.make_schedule_chebyshev(S, coeffs, 0., 40.)
  seeds = J.split(J.PRNGKey(7), B)
  out = c_oracle.langevin(pot, r, phi, np.zeros(B, np.float32), seeds, S, 0, 1e-8, KT, KT / 1e7, 1.0, record=True)
  np.testing.assert_allclose(out['positions'], GOLD_L['shifted_positions'], rtol=0, atol=2e-6 * 40)
  np.testing.assert_allclose(out['work'], GOLD_L['shifted_total_work'], rtol=2e-6)
  np.testing.assert_allclose(out['grad_sums'][0] / B, GOLD_L['shifted_grad_logP'], rtol=1e-5, atol=1e-7)
  np.testing.assert_allclose(out['grad_sums'][1] / B, GOLD_L['shifted_grad_reparam'], rtol=1e-5, atol=1e-7)
  # bit-identical in the overwhelming majority of entries (same arithmetic, libm vs numpy exp/log)
  assert (out['positions'] == GOLD_L['shifted_positions']).mean() > 0.9


def test_langevin_spring_with_equilibration():
  pot = L.V_spring(1.0)
  r, phi = L.make_schedule_chebyshev(96, GOLD_L['spring_coeffs'], 5., 10.)
  out = c_oracle.langevin(pot, r, phi, np.full(24, 5.0, np.float32), J.split(J.PRNGKey(0), 24), 96, 7, 1 / 96, 1.0, 1.0,
                          1.0, record=True)
  np.testing.assert_allclose(out['positions'], GOLD_L['spring_positions'], rtol=0, atol=2e-6 * 10)
  np.testing.assert_allclose(out['logp'], GOLD_L['spring_tot_log_prob'], rtol=2e-6)


def test_ising_bit_exact_vs_numpy_oracle():
  for name, R in (('L64', 3), ('L10', 3)):
    lt, h, s0 = GOLD_I[name + '_log_temp'], GOLD_I[name + '_field'], GOLD_I[name + '_spins0']
    final, summary = c_oracle.ising(lt, h, s0, J.split(J.PRNGKey(0), R))
    np.testing.assert_array_equal(final, GOLD_I[name + '_final'])
    ref = GOLD_I[name + '_summary']
    scale = np.abs(ref[:, 6]).max()
    np.testing.assert_allclose(summary, ref, rtol=0, atol=4 * scale * 2.0 ** -23)
