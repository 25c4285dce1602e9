"""Ising parity AT CONFIG LENGTH (BASELINE.json configs[2] and configs[3]) against the C form of the oracle
(oracle/c `oracle_ising`, which restates noneq_opt/ising.py:109-173) -- the short-T tests of test_gpu_ising.py do
not exercise the per-sweep key derivation split(seed, T-1) at T = 1001 (its counter pairing depends on T), the
float32 summary bookkeeping over 1000 sweeps or the REINFORCE gradient over 1000 sweeps.

Spins: bit-exact (final lattice and intermediate states).  Summaries: the ulp bounds of test_gpu_ising.check_summary.
Gradients: 1e-4 relative to the largest component (north star)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ising_ref as I, jaxlike as J, parameterization_ref as P
from test_gpu_ising import FIELDS, _schedule, check_summary, rel

pytestmark = pytest.mark.gpu


def _config3():
  """L = 64, R = 4096, T = 1001, all -1 start, w = 0.1 RandomState(0).normal((2, 16))."""
  from noneq_opt_b200 import ising
  L, R, T = 64, 4096, 1001
  w = (0.1 * np.random.RandomState(0).normal(size=(2, 16))).astype(np.float32)
  sched = _schedule(w)
  times = np.linspace(0, 1, T, dtype=np.float32)
  params = sched(times)
  s0 = -np.ones((L, L), np.int32)
  seeds = J.split(J.PRNGKey(0), R)
  return L, R, T, w, sched, times, params, s0, seeds


def test_config3_replicas_bit_exact_and_summaries_at_T1001(cuda):
  from noneq_opt_b200 import ising
  L, R, T, w, sched, times, params, s0, seeds = _config3()
  # the schedule tables the kernel sees are the oracle's (monic Chebyshev, ConstrainEndpoints, baselines)
  lt_o = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[0]), 0., 0.), P.log_temp_baseline())(times)
  h_o = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[1]), 0., 0.), P.field_baseline())(times)
  np.testing.assert_array_equal(np.asarray(params.log_temp), lt_o)
  np.testing.assert_array_equal(np.asarray(params.field), h_o)
  pick = [0, 1, R - 1]
  snaps_at = [1, 250, 500, 999]            # lattices after these sweeps (1-based), plus the final one
  fin_o, summ_o, snaps_o = c_oracle.ising(params.log_temp, params.field, s0.astype(np.int8), seeds[pick], snapshots=snaps_at)
  # (i) the whole config-3 ensemble in one launch: final lattices + all seven summary rows of the three replicas
  final, summary = ising.simulate_ising(params, s0, seeds)
  assert final.spins.shape == (R, L, L) and summary.work.shape == (R, T - 1)
  got_final = final.spins[pick].cpu().numpy()
  np.testing.assert_array_equal(got_final, fin_o.astype(got_final.dtype))
  for i, r in enumerate(pick):
    got = np.stack([getattr(summary, f)[r].cpu().numpy() for f in FIELDS])     # [7, T-1]
    check_summary(got, summ_o[i], L * L)
  # (ii) the same three replicas alone, with every intermediate state
  f3, (s3, states) = ising.simulate_ising(params, s0, seeds[pick], return_states=True)
  np.testing.assert_array_equal(f3.spins.cpu().numpy(), got_final)
  st = states.spins.cpu().numpy()                                              # [3, T-1, L, L]
  for q, t in enumerate(snaps_at):
    np.testing.assert_array_equal(st[:, t - 1], snaps_o[:, q].astype(st.dtype))
  np.testing.assert_array_equal(st[:, -1], got_final)
  for f in FIELDS:   # a replica's record does not depend on which launch produced it
    assert torch.equal(getattr(s3, f), getattr(summary, f)[pick])
  # the trajectory is not trivial: the protocol drives the magnetisation from -1 through zero
  m = summary.magnetization[pick].cpu().numpy()
  assert m[:, 0].max() < -0.5 and m[:, -1].min() > 0.5


@pytest.mark.parametrize('loss', ['total_work', 'total_entropy_production'])
def test_config3_reinforce_gradient_at_T1001(cuda, loss):
  """estimate_gradient_rev (ising.py:179-198) over 1000 sweeps against the oracle's closed-form gradient
  (binary64 on the float32 trajectory, itself checked against finite differences in tests/test_oracle_ising.py)."""
  from noneq_opt_b200 import ising
  L, R, T, w, sched, times, params, s0, seeds = _config3()
  loss_fn = getattr(ising, loss)
  pick = [0, 1, R - 1]
  grad, summary = ising.estimate_gradient_rev(loss_fn)(sched, times, s0, seeds[pick])
  got = np.concatenate([grad.log_temp.wrapped.wrapped.weights.cpu().numpy(),
                        grad.field.wrapped.wrapped.weights.cpu().numpy()], -1)              # [3, 32]
  for i, r in enumerate(pick):
    o = I.closed_form_gradient(loss, np.asarray(params.log_temp), np.asarray(params.field), _jac(w[0], times),
                               _jac(w[1], times), s0, seeds[r])
    ref = np.concatenate(o['grad'])
    assert rel(got[i], ref) < 1e-4, (r, rel(got[i], ref))


def _jac(weights, times):
  """d value_t / d w_j of AddBaseline(ConstrainEndpoints(Chebyshev(w), 0, 0), baseline) on (0, 1):
  x (1 - x) T~_j(2x - 1)  (SURVEY Appendix B.2), from the oracle's own parameterisation in binary64."""
  return P.ConstrainEndpoints(P.Chebyshev(np.asarray(weights, np.float64), np.float64), 0., 0.).jacobian(
      times.astype(np.float64))


def test_config4_lattice_through_the_segment_scheduler_bit_exact(cuda, monkeypatch):
  """configs[3] lattice (1024^2, random_spins(0.5, PRNGKey(1)) start, baseline schedule), 240 sweeps, 150 replicas
  on 148 SMs, i.e. THROUGH the segment scheduler (every replica's sweeps cut into ticketed segments; lattice, spin
  sum and bond sum carried through the workspace).  Replicas 0 and 149 against the C oracle: final lattices
  bit-exact, summaries within the ulp bounds; and the unsegmented launch must give the same bits."""
  from noneq_opt_b200 import ising
  L, R, T = 1024, 150, 241
  times = np.linspace(0, 1, T, dtype=np.float32)
  lt = np.asarray(ising.log_temp_baseline(0.69, 10, 1)(times), np.float32)
  h = np.asarray(ising.field_baseline(-1, 1)(times), np.float32)
  np.testing.assert_array_equal(lt, P.log_temp_baseline(0.69, 10, 1)(times))
  np.testing.assert_array_equal(h, P.field_baseline(-1, 1)(times))
  s0 = I.random_spins((L, L), .5, J.PRNGKey(1))
  np.testing.assert_array_equal(ising.random_spins([L, L], .5, J.PRNGKey(1)).cpu().numpy(), s0)
  seeds = J.split(J.PRNGKey(0), 256)[:R]
  pick = [0, R - 1]
  fin_o, summ_o = c_oracle.ising(lt, h, s0.astype(np.int8), seeds[pick])
  final, summary = ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds)
  got = final.spins[pick].cpu().numpy()
  np.testing.assert_array_equal(got, fin_o.astype(got.dtype))
  for i, r in enumerate(pick):
    check_summary(np.stack([getattr(summary, f)[r].cpu().numpy() for f in FIELDS]), summ_o[i], L * L)
  monkeypatch.setenv('NQ_ISING_NO_SEGMENTS', '1')
  final1, summary1 = ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds)
  assert torch.equal(final1.spins, final.spins)
  for f in FIELDS:
    assert torch.equal(getattr(summary1, f), getattr(summary, f))
