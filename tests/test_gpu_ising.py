"""GPU parity of the Ising path (csrc/ising.cu through the C ABI and the reference-shaped Python
API) against the CPU oracle and the committed golden vectors.

Spins are compared BIT-EXACTLY (same threefry uniforms, integer acceptance thresholds).  Float
summaries: forward/reverse log-probs are exact sums of the oracle's float32 terms (tolerance 2
float32 ulp of the sum); work / heat / energy are differences of O(L^2) float32 energy sums in the
reference, whose own rounding noise is a few ulp of |E| -- the tolerance says so.  Gradients:
1e-4 relative to the largest component (north star)."""
import os

import numpy as np
import pytest
import torch

from oracle import ising_ref as I, jaxlike as J, parameterization_ref as P

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ising.npz'))
FIELDS = I.IsingSummary._fields


def rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check_summary(got, ref, n_sites):
  """got/ref: [7, ...] arrays in IsingSummary field order."""
  got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
  e_scale = max(np.abs(ref[6]).max(), 2.0 * n_sites)
  ulp_e = e_scale * 2.0 ** -23
  for f, name in enumerate(FIELDS):
    if name in ('work', 'dissipated_heat', 'energy'):
      assert np.abs(got[f] - ref[f]).max() <= 4 * ulp_e, name
    elif name == 'magnetization':
      np.testing.assert_array_equal(got[f].astype(np.float32), ref[f].astype(np.float32))
    else:
      scale = max(np.abs(ref[2]).max(), np.abs(ref[3]).max())
      assert np.abs(got[f] - ref[f]).max() <= 4 * scale * 2.0 ** -23, name


def _schedule(weights):
  from noneq_opt_b200 import ising, parameterization as p10n
  return ising.IsingSchedule(
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(weights[0]), 0., 0.), ising.log_temp_baseline()),
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(weights[1]), 0., 0.), ising.field_baseline()))


@pytest.mark.parametrize('name, L, T, R', [('L64', 64, 6, 3), ('L10', 10, 5, 3), ('L128', 128, 4, 2)])
def test_golden_spin_histories_bit_exact(cuda, name, L, T, R):
  from noneq_opt_b200 import ising
  sched = _schedule(GOLD[name + '_weights'])
  times = np.linspace(0, 1, T, dtype=np.float32)
  params = sched(times)
  np.testing.assert_array_equal(params.log_temp, GOLD[name + '_log_temp'])
  np.testing.assert_array_equal(params.field, GOLD[name + '_field'])
  seeds = J.split(J.PRNGKey(0), R)
  final, (summary, states) = ising.simulate_ising(params, GOLD[name + '_spins0'], seeds, return_states=True)
  assert final.spins.shape == (R, L, L) and states.spins.shape == (R, T - 1, L, L)
  np.testing.assert_array_equal(final.spins.cpu().numpy(), GOLD[name + '_final'])
  np.testing.assert_array_equal(states.spins.cpu().numpy(), GOLD[name + '_states'])
  got = np.stack([getattr(summary, f).cpu().numpy() for f in FIELDS])        # [7, R, T-1]
  check_summary(got, np.transpose(GOLD[name + '_summary'], (1, 0, 2)), L * L)
  # single-replica call: same nesting as lax.scan over update (ising.py:153-173)
  f1, s1 = ising.simulate_ising(params, GOLD[name + '_spins0'], seeds[1])
  assert f1.spins.shape == (L, L) and s1.work.shape == (T - 1,)
  np.testing.assert_array_equal(f1.spins.cpu().numpy(), GOLD[name + '_final'][1])


@pytest.mark.parametrize('name, L, T, R', [('L64', 64, 6, 3), ('L10', 10, 5, 3), ('L128', 128, 4, 2)])
@pytest.mark.parametrize('loss', ['work', 'entropy'])
def test_golden_gradients(cuda, name, L, T, R, loss):
  from noneq_opt_b200 import ising
  sched = _schedule(GOLD[name + '_weights'])
  times = np.linspace(0, 1, T, dtype=np.float32)
  seeds = J.split(J.PRNGKey(0), R)
  loss_fn = ising.total_work if loss == 'work' else ising.total_entropy_production
  gold = GOLD[f'{name}_grad_{loss}']              # [R, 6*8 + 1]
  grad, lp, rp, summary = ising.estimate_gradient_fwd_REINFORCE_track_terms(loss_fn)(
      sched, times, GOLD[name + '_spins0'], seeds)
  def flat(g):
    return np.concatenate([g.log_temp.wrapped.wrapped.weights.cpu().numpy(),
                           g.field.wrapped.wrapped.weights.cpu().numpy()], -1)
  for got, sl in ((flat(grad), slice(0, 16)), (flat(lp), slice(16, 32)), (flat(rp), slice(32, 48))):
    for r in range(R):
      assert rel(got[r], gold[r, sl]) < 1e-4
  g_rev, _ = ising.estimate_gradient_rev(loss_fn)(sched, times, GOLD[name + '_spins0'], seeds)
  g_fwd, _ = ising.estimate_gradient_fwd(loss_fn)(sched, times, GOLD[name + '_spins0'], seeds)
  np.testing.assert_allclose(flat(g_rev), flat(g_fwd), rtol=1e-4)        # tests/ising_test.py:129-145
  np.testing.assert_allclose(flat(g_rev), flat(grad), rtol=1e-6)
  g1, s1 = ising.estimate_gradient_rev(loss_fn)(sched, times, GOLD[name + '_spins0'], seeds[0])
  assert g1.field.wrapped.wrapped.weights.shape == (8,) and s1.work.shape == (T - 1,)


def test_random_spins_and_measures_bit_exact(cuda):
  from noneq_opt_b200 import ising
  for L in (10, 64, 256):
    got = ising.random_spins([L, L], .5, J.PRNGKey(1)).cpu().numpy()
    np.testing.assert_array_equal(got, I.random_spins((L, L), .5, J.PRNGKey(1)))
  s = I.random_spins((64, 64), .3, J.PRNGKey(2))
  np.testing.assert_array_equal(ising.sum_neighbors(s).cpu().numpy(), I.sum_neighbors(s))
  st = ising.IsingState(s, ising.IsingParameters(np.float32(0.3), np.float32(-0.7)))
  np.testing.assert_allclose(float(ising.energy(st)), I.energy(s, -0.7), rtol=1e-6)
  np.testing.assert_allclose(ising.flip_logits(s, st.params).cpu().numpy(), I.flip_logits(s, 0.3, -0.7), rtol=1e-6)


@pytest.mark.parametrize('L', [256, 64])
@pytest.mark.parametrize('field', [-1.0, 0.5])
@pytest.mark.parametrize('temperature', [float(np.exp(-1.)), float(np.exp(1.))])
def test_detailed_balance_and_entropy_production(cuda, L, field, temperature):
  """tests/ising_test.py:71-109 re-expressed against the new API; also bit-exact vs the oracle."""
  from noneq_opt_b200 import ising
  init_seed, sim_seed = J.split(J.PRNGKey(0))
  params = ising.IsingParameters(np.log(np.float32(temperature)), np.float32(field))
  s0 = ising.random_spins([L, L], .5, init_seed)
  state0 = ising.IsingState(s0, params)
  lp0 = -float(ising.energy(state0)) / temperature
  state1, summ = ising.update(state0, params, sim_seed)
  lp1 = -float(ising.energy(state1)) / temperature
  np.testing.assert_allclose(lp0 + float(summ.forward_log_prob), lp1 + float(summ.reverse_log_prob), rtol=5e-4)
  np.testing.assert_allclose(float(summ.dissipated_heat), float(summ.entropy_production) * temperature, rtol=1e-4,
                             atol=1e-3)
  s1_o, summ_o = I.update(s0.cpu().numpy(), (params.log_temp, field), (params.log_temp, field), sim_seed)
  np.testing.assert_array_equal(state1.spins.cpu().numpy(), s1_o)
  check_summary(np.array([float(getattr(summ, f)) for f in FIELDS]), np.array([getattr(summ_o, f) for f in FIELDS]), L * L)


@pytest.mark.parametrize('L, field, temperature, expected', [(64, 0., 1e6, .5), (256, 3., .2, 1.), (64, -4., .5, 0.)])
def test_statistics(cuda, L, field, temperature, expected):
  from noneq_opt_b200 import ising
  init_seed, sim_seed = J.split(J.PRNGKey(0), 2)
  s0 = ising.random_spins([L, L], .5, init_seed)
  params = ising.IsingParameters(np.full(1000, np.log(temperature), np.float32), np.full(1000, field, np.float32))
  state, _ = ising.simulate_ising(params, s0, sim_seed)
  np.testing.assert_allclose(float(((state.spins + 1) / 2).float().mean()), expected, atol=.05)


def test_counts_and_partials_match_oracle_trace(cuda):
  from noneq_opt_b200 import ising
  L, T = 64, 5
  lt, h = GOLD['L64_log_temp'][:T], GOLD['L64_field'][:T]
  s0 = I.random_spins((L, L), .5, J.PRNGKey(3))
  seed = J.PRNGKey(9)
  trace = []
  I.simulate_ising(lt, h, s0, seed, trace=trace)
  _, _, spins, _, _, seeds, broadcast, bits = ising._prepare(ising.IsingParameters(lt, h), s0, seed)
  run = ising._run_kernel(lt, h, bits, broadcast, seeds, L, want_partials=True, want_counts=True)
  counts = run.counts.cpu().numpy()[0]            # [T-1, half, active/flipped, 16]
  for t in range(1, T):
    tr = trace[t - 1]
    cur = tr['spins_before']
    for half in range(2):
      mask = tr['masks'][half].astype(bool)
      slot = 8 * (cur > 0) + (I.sum_neighbors(cur) + 4) // 2
      np.testing.assert_array_equal(counts[t - 1, half, 0], np.bincount(slot[mask], minlength=16))
      np.testing.assert_array_equal(counts[t - 1, half, 1], np.bincount(slot[mask & (tr['flips'][half] > 0)], minlength=16))
      cur = tr['spins_mid'] if half == 0 else tr['spins_after']
  cf = I.closed_form_gradient('total_work', lt, h, np.eye(T), np.eye(T), s0, seed)
  part = run.partials.cpu().numpy()[:, 0]
  assert rel(part[0], cf['per_step']['dfwd_dh'][1:]) < 1e-5
  assert rel(part[1], cf['per_step']['dfwd_dl'][1:]) < 1e-5


def test_training_step_runs(cuda):
  """tests/ising_test.py:148-175: 20 Adam steps stay finite; batch 32, 11 time steps."""
  from noneq_opt_b200 import ising, optimizers, parameterization as p10n, random as nqr, tree
  for loss_fn in (ising.total_work, ising.total_entropy_production):
    for mode in ('fwd', 'rev'):
      schedule = ising.IsingSchedule(log_temp=p10n.Constant(1.), field=p10n.Chebyshev(np.ones(8, np.float32)))
      opt = optimizers.adam(1e-3)
      state = opt.init_fn(schedule)
      step_fn = ising.get_train_step(opt, -np.ones([10, 10], np.float32), 32, 11, loss_fn, mode)
      seed = J.PRNGKey(0)
      for step in range(20):
        seed, split = nqr.split_host(seed)
        state, summary = step_fn(state, step, split)
      assert summary.work.shape == (32, 10)
      for leaf in tree.tree_leaves(state):
        assert np.isfinite(np.asarray(leaf)).all()
  out = ising.get_train_step_REINFORCE_track_terms(opt, -np.ones([10, 10], np.float32), 8, 5, ising.total_work, 'fwd')(
      opt.init_fn(schedule), 0, J.PRNGKey(1))
  assert len(out) == 5


def test_config4_lattice_properties_and_one_sweep_parity(cuda):
  """configs[3] lattice size (1024^2, shared-memory resident, CTA per replica): a short run is
  bit-exact vs the oracle, and the running energy agrees with a from-scratch measurement."""
  from noneq_opt_b200 import ising
  L, T, R = 1024, 3, 2
  s0 = I.random_spins((L, L), .5, J.PRNGKey(1))
  lt = np.log(np.array([2.5, 2.3, 2.1], np.float32))
  h = np.array([-0.2, 0.0, 0.2], np.float32)
  seeds = J.split(J.PRNGKey(0), R)
  final, summary = ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds)
  fin_o, summ_o = I.simulate_ising(lt, h, s0, seeds[1])
  np.testing.assert_array_equal(final.spins[1].cpu().numpy(), fin_o)
  got = np.stack([getattr(summary, f)[1].cpu().numpy() for f in FIELDS])
  check_summary(got, np.stack([getattr(summ_o, f) for f in FIELDS]), L * L)
  e_meas = float(ising.energy(ising.IsingState(final.spins[0], ising.IsingParameters(lt[-1], h[-1]))))
  np.testing.assert_allclose(float(summary.energy[0, -1]), e_meas, rtol=1e-6)


def test_shape_errors(cuda):
  from noneq_opt_b200 import ising
  p = ising.IsingParameters(np.zeros(3, np.float32), np.zeros(3, np.float32))
  with pytest.raises(ValueError):
    ising.simulate_ising(p, -np.ones((9, 9), np.int32), J.PRNGKey(0))
  with pytest.raises(ValueError):
    ising.simulate_ising(p, -np.ones((8, 8), np.int32), J.PRNGKey(0), checkpoint_every=5)
  with pytest.raises(NotImplementedError):
    ising.simulate_ising(p, -np.ones((2, 2, 2, 2), np.int32), J.PRNGKey(0))
  with pytest.raises(ValueError):
    ising.simulate_ising(p, -np.ones((4, 6, 5), np.int32), J.PRNGKey(0))


def test_masked_update_building_block(cuda):
  """ising.masked_update (ising.py:109-126) with the reference's own masks reproduces update()."""
  from noneq_opt_b200 import ising, random as nqr
  L = 64
  s0 = ising.random_spins([L, L], .5, J.PRNGKey(4))
  params = ising.IsingParameters(np.float32(0.2), np.float32(-0.3))
  seed = J.PRNGKey(12)
  seed_even, seed_odd = nqr.split_host(seed, 2)
  mask_even, mask_odd = ising.even_odd_masks((L, L))
  st, ef, er = ising.masked_update(ising.IsingState(s0, params), params, mask_even, seed_even)
  st, of, orv = ising.masked_update(st, params, mask_odd, seed_odd)
  ref, summ = ising.update(ising.IsingState(s0, params), params, seed)
  np.testing.assert_array_equal(st.spins.cpu().numpy(), ref.spins.cpu().numpy())
  np.testing.assert_allclose(float(ef + of), float(summ.forward_log_prob), rtol=2e-5)
  np.testing.assert_allclose(float(er + orv), float(summ.reverse_log_prob), rtol=2e-5)
  s_o, f_o, r_o, _, _ = I.masked_update(s0.cpu().numpy(), 0.2, -0.3, mask_even, seed_even)
  np.testing.assert_allclose(float(ef), f_o, rtol=2e-5)


@pytest.mark.parametrize('shape', [(1000,), (6, 10), (8, 6, 4), (16, 16, 16)])
def test_other_lattice_ranks_bit_exact(cuda, shape):
  """sum_neighbors / even_odd_masks are rank-generic in the reference (ising.py:75-106; its tests use
  1-D chains and 64^3 cubes): the generic kernel handles 1-3 dimensions and non-square lattices."""
  from noneq_opt_b200 import ising
  T, R = 5, 3
  rs = np.random.RandomState(1)
  lt = rs.uniform(-0.5, 1.0, T).astype(np.float32)
  h = rs.uniform(-1, 1, T).astype(np.float32)
  s0 = I.random_spins(shape, .5, J.PRNGKey(7))
  np.testing.assert_array_equal(ising.random_spins(shape, .5, J.PRNGKey(7)).cpu().numpy(), s0)
  seeds = J.split(J.PRNGKey(2), R)
  final, (summary, states) = ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds, return_states=True)
  assert final.spins.shape == (R,) + shape and states.spins.shape == (R, T - 1) + shape
  for r in range(R):
    fin_o, summ_o, st_o = I.simulate_ising(lt, h, s0, seeds[r], return_states=True)
    np.testing.assert_array_equal(final.spins[r].cpu().numpy(), fin_o)
    np.testing.assert_array_equal(states.spins[r].cpu().numpy(), st_o)
    got = np.stack([getattr(summary, f)[r].cpu().numpy() for f in FIELDS])
    check_summary(got, np.stack([getattr(summ_o, f) for f in FIELDS]), int(np.prod(shape)))
  # gradients through the same closed forms
  from noneq_opt_b200 import parameterization as p10n
  sched = ising.IsingSchedule(p10n.Chebyshev(np.array([0.2, -0.1, 0.05], np.float32)),
                              p10n.Chebyshev(np.array([-0.5, 0.7, 0.1], np.float32)))
  times = np.linspace(0, 1, T, dtype=np.float32)
  g, _ = ising.estimate_gradient_rev(ising.total_entropy_production)(sched, times, s0, seeds[0])
  so = P.Chebyshev(np.array([0.2, -0.1, 0.05], np.float32)), P.Chebyshev(np.array([-0.5, 0.7, 0.1], np.float32))
  cf = I.closed_form_gradient('total_entropy_production', so[0](times), so[1](times), so[0].jacobian(times),
                              so[1].jacobian(times), s0, seeds[0])
  assert rel(g.log_temp.weights.cpu().numpy(), cf['grad'][0]) < 1e-4
  assert rel(g.field.weights.cpu().numpy(), cf['grad'][1]) < 1e-4


@pytest.mark.parametrize('shape, field, temperature, expected', [((10000,), 0., 1., .5), ((64, 64, 64), -4., .5, 0.)])
def test_statistics_reference_shapes(cuda, shape, field, temperature, expected):
  """tests/ising_test.py:111-127 at the reference's own 1-D and 3-D sizes (64^3 spins live in the
  workspace, not in shared memory); 200 sweeps are enough to reach the asserted proportions."""
  from noneq_opt_b200 import ising
  init_seed, sim_seed = J.split(J.PRNGKey(0), 2)
  s0 = ising.random_spins(shape, .5, init_seed)
  params = ising.IsingParameters(np.full(200, np.log(temperature), np.float32), np.full(200, field, np.float32))
  state, summ = ising.simulate_ising(params, s0, sim_seed)
  np.testing.assert_allclose(float(((state.spins + 1) / 2).float().mean()), expected, atol=.05)
  # heat == T * entropy production on every sweep (tests/ising_test.py:91-109)
  np.testing.assert_allclose(summ.dissipated_heat.cpu().numpy(), summ.entropy_production.cpu().numpy() * temperature,
                             rtol=1e-4, atol=2e-2)


def test_cta_team_path_states_counts_and_gradients(cuda):
  """L = 256 runs one 1024-thread CTA per replica (histograms through shared-memory atomics, summary
  emitted by warp 0 while the other warps start the next sweep): states, counts and gradients."""
  from noneq_opt_b200 import ising, parameterization as p10n
  L, T, R = 256, 4, 3
  wl, wh = np.array([0.3, -0.2, 0.1], np.float32), np.array([-0.4, 0.6, 0.2], np.float32)
  sched = ising.IsingSchedule(p10n.Chebyshev(wl), p10n.Chebyshev(wh))
  times = np.linspace(0, 1, T, dtype=np.float32)
  params = sched(times)
  s0 = I.random_spins((L, L), .5, J.PRNGKey(11))
  seeds = J.split(J.PRNGKey(3), R)
  final, (summary, states) = ising.simulate_ising(params, s0, seeds, return_states=True)
  for r in (0, 2):
    fin_o, summ_o, st_o = I.simulate_ising(params.log_temp, params.field, s0, seeds[r], return_states=True)
    np.testing.assert_array_equal(states.spins[r].cpu().numpy(), st_o)
    np.testing.assert_array_equal(final.spins[r].cpu().numpy(), fin_o)
    check_summary(np.stack([getattr(summary, f)[r].cpu().numpy() for f in FIELDS]),
                  np.stack([getattr(summ_o, f) for f in FIELDS]), L * L)
  for loss_fn, name in ((ising.total_work, 'total_work'), (ising.total_entropy_production, 'total_entropy_production')):
    g, _ = ising.estimate_gradient_rev(loss_fn)(sched, times, s0, seeds)
    so = P.Chebyshev(wl), P.Chebyshev(wh)
    cf = I.closed_form_gradient(name, so[0](times), so[1](times), so[0].jacobian(times), so[1].jacobian(times), s0, seeds[1])
    assert rel(g.log_temp.weights[1].cpu().numpy(), cf['grad'][0]) < 1e-4
    assert rel(g.field.weights[1].cpu().numpy(), cf['grad'][1]) < 1e-4


def test_update_with_a_batch_of_seeds(cuda):
  from noneq_opt_b200 import ising
  L = 64
  s0 = I.random_spins((L, L), .5, J.PRNGKey(5))
  params = ising.IsingParameters(np.float32(0.4), np.float32(0.25))
  seeds = J.split(J.PRNGKey(8), 5)
  state, summ = ising.update(ising.IsingState(s0, params), params, seeds)
  assert state.spins.shape == (5, L, L) and summ.work.shape == (5,)
  for r in range(5):
    s_o, summ_o = I.update(s0, (0.4, 0.25), (0.4, 0.25), seeds[r])
    np.testing.assert_array_equal(state.spins[r].cpu().numpy(), s_o)


def test_lattice_larger_than_shared_memory(cuda):
  """L = 2048 (1 MB of bits per replica) does not fit one SM's shared memory: the same kernel runs
  with its colour planes in global memory (L2).  Bit-exact against the oracle."""
  from noneq_opt_b200 import ising
  L, T, R = 2048, 3, 2
  s0 = I.random_spins((L, L), .5, J.PRNGKey(1))
  lt = np.log(np.array([2.5, 2.3, 2.1], np.float32))
  h = np.array([-0.2, 0.0, 0.2], np.float32)
  seeds = J.split(J.PRNGKey(0), R)
  final, summary = ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds)
  fin_o, summ_o = I.simulate_ising(lt, h, s0, seeds[1])
  np.testing.assert_array_equal(final.spins[1].cpu().numpy(), fin_o)
  check_summary(np.stack([getattr(summary, f)[1].cpu().numpy() for f in FIELDS]),
                np.stack([getattr(summ_o, f) for f in FIELDS]), L * L)


def test_arbitrary_loss_function_through_torch_autograd(cuda):
  """LossFn is an arbitrary callable in the reference (ising.py:177).  A loss written with torch ops is
  differentiated wrt the summary / end-point states by torch autograd and chained analytically to the
  schedule; checked against the closed-form fast path and against brute-force finite differences."""
  from noneq_opt_b200 import ising, parameterization as p10n
  T, R = 6, 4
  w = (0.1 * np.random.RandomState(0).normal(size=(2, 8))).astype(np.float32)
  sched = _schedule(w)
  times = np.linspace(0, 1, T, dtype=np.float32)
  s0 = -np.ones((10, 10), np.float32)
  seeds = J.split(J.PRNGKey(5), R)

  def flat(g):
    return np.concatenate([g.log_temp.wrapped.wrapped.weights.cpu().numpy(), g.field.wrapped.wrapped.weights.cpu().numpy()], -1)

  def my_work(initial_state, final_state, summary):
    return summary.work.sum()

  def my_entropy(initial_state, final_state, summary):
    return summary.entropy_production.sum() + ising.log_prob(initial_state) - ising.log_prob(final_state)

  for custom, builtin in ((my_work, ising.total_work), (my_entropy, ising.total_entropy_production)):
    g_c, _ = ising.estimate_gradient_rev(custom)(sched, times, s0, seeds)
    g_b, _ = ising.estimate_gradient_rev(builtin)(sched, times, s0, seeds)
    np.testing.assert_allclose(flat(g_c), flat(g_b), rtol=2e-4, atol=2e-4 * np.abs(flat(g_b)).max())

  def exotic(initial_state, final_state, summary):
    return (1e-3 * (summary.dissipated_heat ** 2).sum() + 0.1 * summary.energy[-1] + 0.01 * summary.forward_log_prob.sum()
            - summary.reverse_log_prob.mean() + 0.05 * ising.energy(final_state))

  def exotic_oracle(initial_spins, final_spins, log_temp, field, summary, dtype=np.float32):
    return (1e-3 * (np.asarray(summary.dissipated_heat, np.float64) ** 2).sum() + 0.1 * summary.energy[-1]
            + 0.01 * np.sum(summary.forward_log_prob) - np.mean(summary.reverse_log_prob)
            + 0.05 * I.energy(final_spins, field[-1], np.float64))

  I.LOSSES['exotic'] = exotic_oracle
  try:
    def sched_fn(wl, wh, ts):
      d = wl.dtype.type
      lt = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(wl, d), 0., 0.), P.log_temp_baseline(dtype=d))
      hh = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(wh, d), 0., 0.), P.field_baseline(dtype=d))
      return lt(ts.astype(d)), hh(ts.astype(d))
    g, _ = ising.estimate_gradient_rev(exotic)(sched, times, s0, seeds[2])
    fd = I.fd_gradient('exotic', sched_fn, w[0].astype(np.float64), w[1].astype(np.float64), times.astype(np.float64),
                       s0.astype(np.int32), seeds[2])
    ref = np.concatenate(fd['reinforce'])
    assert rel(flat(g), ref) < 5e-4
  finally:
    del I.LOSSES['exotic']


@pytest.mark.gpu
@pytest.mark.parametrize('L,R,T', [(512, 149, 9), (1024, 150, 6), (512, 149, 1100), (64, 2400, 9), (128, 2404, 41),
                                   (64, 4096, 1001)])
def test_segment_scheduled_run_is_bit_identical(cuda, L, R, T, monkeypatch):
  """When R is not a multiple of the resident CTAs, nq_ising_run cuts every replica's sweeps into
  segments handed out through a ticket counter (lattice, M and B carried through the workspace) -- per
  replica for the CTA-team kernel (L >= 256), per CTA of four lock-stepped replicas for the warp-team kernel
  (L <= 128; the last case is configs[2]).  Outputs must be bit-identical to the unsegmented launch, and one
  replica to the oracle."""
  import torch
  from noneq_opt_b200 import ising
  monkeypatch.setenv('NQ_ISING_WARP_SEGMENTS', '1')    # warp-team segments are opt-in (no gain measured)
  lt = np.log(np.linspace(2.6, 2.0, T).astype(np.float32))
  h = np.linspace(-0.3, 0.3, T).astype(np.float32)
  params = ising.IsingParameters(lt, h)
  s0 = I.random_spins((L, L), .5, J.PRNGKey(21))
  seeds = J.split(J.PRNGKey(4), R)

  def run():
    final, summary = ising.simulate_ising(params, s0, seeds, return_states=(T == 9))
    states = None
    if T == 9:
      summary, states = summary
    out = [final.spins] + [getattr(summary, f) for f in FIELDS]
    if states is not None:
      out.append(states.spins[-3:])
    torch.cuda.synchronize()
    return [o.clone() for o in out]

  seg = run()
  monkeypatch.setenv('NQ_ISING_NO_SEGMENTS', '1')
  flat = run()
  monkeypatch.delenv('NQ_ISING_NO_SEGMENTS')
  for a, b in zip(seg, flat):
    assert torch.equal(a, b)
  if T > 100:     # long run: segments are capped at ~512 sweeps; the oracle check is done by the short cases
    return
  r = R - 2
  fin_o, summ_o = I.simulate_ising(lt, h, s0, seeds[r])
  np.testing.assert_array_equal(seg[0][r].cpu().numpy(), fin_o)
  check_summary(np.stack([seg[1 + i][r].cpu().numpy() for i in range(len(FIELDS))]),
                np.stack([getattr(summ_o, f) for f in FIELDS]), L * L)


def test_config4_ensemble_carries_energy_and_magnetization_through_segments(cuda):
  """configs[3] shape (256 replicas of 1024^2, segment-scheduled over 148 SMs), 40 sweeps: the running
  magnetisation and energy each replica reports after its LAST sweep must equal a from-scratch
  measurement of its final lattice (the integer spin and bond sums are handed from segment to
  segment), and the work/heat bookkeeping must telescope: E_final - E_initial = sum(work) - sum(heat)."""
  from noneq_opt_b200 import ising
  L, R, T = 1024, 256, 41
  times = np.linspace(0, 1, T, dtype=np.float32)
  lt, h = ising.log_temp_baseline(0.69, 10, 1)(times), ising.field_baseline(-1, 1)(times)
  s0 = ising.random_spins([L, L], .5, J.PRNGKey(1))
  seeds = J.split(J.PRNGKey(0), R)
  final, summary = ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds)
  spins = final.spins.to(torch.float64)                                  # [R, L, L]
  nbr = sum(torch.roll(spins, sh, ax) for ax in (1, 2) for sh in (-1, 1))
  e_meas = -(spins * (nbr / 2 + float(h[-1]))).sum((1, 2))
  m_meas = spins.mean((1, 2))
  np.testing.assert_allclose(summary.energy[:, -1].double().cpu().numpy(), e_meas.cpu().numpy(), rtol=2e-7)
  np.testing.assert_allclose(summary.magnetization[:, -1].double().cpu().numpy(), m_meas.cpu().numpy(), atol=1e-7)
  s0f = torch.as_tensor(np.asarray(s0.cpu() if hasattr(s0, 'cpu') else s0), dtype=torch.float64)
  nbr0 = sum(torch.roll(s0f, sh, ax) for ax in (0, 1) for sh in (-1, 1))
  e0 = float(-(s0f * (nbr0 / 2 + float(h[0]))).sum())
  tele = summary.work.double().sum(1) - summary.dissipated_heat.double().sum(1)
  np.testing.assert_allclose((e_meas - e0).cpu().numpy(), tele.cpu().numpy(), rtol=0, atol=2e-5 * L * L)
  assert len(set(np.round(m_meas.cpu().numpy(), 9))) > R // 2          # replicas are distinct trajectories


@pytest.mark.parametrize('shape', [(64, 64), (10, 10), (6, 96), (1024, 1024), (64,), (10000,), (4, 6, 32), (8, 6, 4), (2, 2), (2, 32)])
def test_measure_entry_point_matches_numpy(cuda, shape):
  """nq_ising_measure: {sum of spins, sum over bonds of s_i s_j} per lattice, word-parallel when the
  contiguous axis is a multiple of 32 and per site otherwise; periodic in every axis (an axis of length
  2 counts its bond twice, as sum_neighbors does)."""
  import ctypes as C
  from noneq_opt_b200 import _lib, ising
  R = 3
  rs = np.random.RandomState(sum(shape))
  spins = np.where(rs.uniform(size=(R,) + shape) < 0.4, 1, -1).astype(np.int32)
  bits = ising._pack(torch.from_numpy(spins).cuda(), lattice_ndim=len(shape))
  out = torch.empty((R, 2), dtype=torch.int64, device='cuda')
  dims = (C.c_int32 * 3)(*(list(shape) + [1] * (3 - len(shape))))
  _lib.check(_lib.lib().nq_ising_measure(_lib.ptr(bits), len(shape), dims, R, _lib.ptr(out), _lib.stream_arg()))
  s64 = spins.astype(np.int64)
  bonds = sum((s64 * np.roll(s64, -1, axis=1 + ax)).sum(axis=tuple(range(1, 1 + len(shape)))) for ax in range(len(shape)))
  np.testing.assert_array_equal(out.cpu().numpy()[:, 0], s64.reshape(R, -1).sum(1))
  np.testing.assert_array_equal(out.cpu().numpy()[:, 1], bonds)


@pytest.mark.parametrize('loss_name', ['total_work', 'total_entropy_production'])
@pytest.mark.parametrize('L,R,T', [(64, 7, 9), (10, 3, 700)])
def test_gradient_assembly_kernel_matches_torch_contraction(cuda, loss_name, L, R, T):
  """nq_ising_grad_assemble against the same contraction written with torch float64 ops
  (ising._loss_and_sensitivities + matmul with the schedule Jacobians): losses bit-equal, gradients to
  float64 rounding.  T = 700 crosses the kernel's 512-sweep tile."""
  from noneq_opt_b200 import _lib, ising, parameterization as p10n
  rs = np.random.RandomState(T)
  sched = ising.IsingSchedule(
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev((0.1 * rs.normal(size=5)).astype(np.float32)), 0., 0.),
                       ising.log_temp_baseline(0.69, 10, 1)),
      p10n.Chebyshev(np.array([-0.4, 0.6, 0.2], np.float32)))
  times = np.linspace(0, 1, T, dtype=np.float32)
  s0 = I.random_spins((L, L), .5, J.PRNGKey(3))
  loss_fn = getattr(ising, loss_name)
  t = ising._gradient_terms(loss_fn, sched, times, s0, J.split(J.PRNGKey(1), R))
  run = t['run']
  _, lt, h, J_lt, J_h = ising._schedule_tables(sched, times)
  loss, dL_dh, dL_dl = ising._loss_and_sensitivities(ising._loss_kind(loss_fn), run, lt, h, L)
  P = run.partials.double()
  dF_dh, dF_dl = torch.zeros_like(dL_dh), torch.zeros_like(dL_dl)
  dF_dh[:, 1:], dF_dl[:, 1:] = P[_lib.NQ_IS_DFWD_DH], P[_lib.NQ_IS_DFWD_DL]
  Jl, Jh = torch.from_numpy(J_lt).cuda(), torch.from_numpy(J_h).cuda()
  assert torch.equal(t['loss'], loss)
  lw = t['loss'].double()[:, None]
  for got, want in ((t['logP'][0], (lw * dF_dl) @ Jl), (t['logP'][1], (lw * dF_dh) @ Jh),
                    (t['reparam'][0], dL_dl @ Jl), (t['reparam'][1], dL_dh @ Jh)):
    scale = float(want.abs().max()) + 1e-300
    assert float((got - want).abs().max()) <= 1e-11 * scale + 1e-9


@pytest.mark.parametrize('shape,R,T', [((64, 64, 64), 2, 4), ((4, 2, 64), 3, 6), ((8, 6, 128), 2, 5), ((12, 4, 64), 2, 4)])
def test_bit_packed_3d_kernel_is_bit_exact(cuda, shape, R, T, monkeypatch):
  """ising_fast3d_kernel: cuboid lattices with d2 % 64 == 0, d1 even, d0 % 4 == 0 (64^3 is the reference's own 3-D test
  shape, tests/ising_test.py:71) run bit-packed in shared memory -- six neighbour words per word, 2 x 7 site classes,
  one threefry block per TWO spin-updates (counter partner (i + d0/2, j, k) has the same colour).  Spin histories
  bit-exact against the oracle, summaries within the ulp bounds, and everything equal to the byte-per-spin generic
  kernel (NQ_ISING_NO_FAST3D=1)."""
  from noneq_opt_b200 import ising
  lt = np.log(np.linspace(4.2, 3.4, T).astype(np.float32))
  h = np.linspace(-0.4, 0.5, T).astype(np.float32)
  params = ising.IsingParameters(lt, h)
  s0 = I.random_spins(shape, .45, J.PRNGKey(13))
  seeds = J.split(J.PRNGKey(6), R)
  n_sites = int(np.prod(shape))

  def run():
    final, (summary, states) = ising.simulate_ising(params, s0, seeds, return_states=True)
    torch.cuda.synchronize()
    return final.spins.clone(), [getattr(summary, f).clone() for f in FIELDS], states.spins.clone()
  fin, summ, states = run()
  assert tuple(fin.shape) == (R,) + shape and tuple(states.shape) == (R, T - 1) + shape
  for r in range(R):
    fin_o, summ_o, states_o = I.simulate_ising(lt, h, s0, seeds[r], return_states=True)
    np.testing.assert_array_equal(fin[r].cpu().numpy(), fin_o)
    np.testing.assert_array_equal(states[r].cpu().numpy(), states_o)
    check_summary(np.stack([s[r].cpu().numpy() for s in summ]), np.stack([getattr(summ_o, f) for f in FIELDS]), n_sites)
  monkeypatch.setenv('NQ_ISING_NO_FAST3D', '1')
  fin_g, summ_g, states_g = run()
  assert torch.equal(fin_g, fin) and torch.equal(states_g, states)
  for a, b in zip(summ_g, summ):
    assert torch.equal(a, b)


def test_bit_packed_3d_kernel_gradients(cuda):
  """REINFORCE gradients of both losses on a 3-D lattice that takes the bit-packed path, against the oracle's
  closed-form gradient (binary64 on the float32 trajectory)."""
  from noneq_opt_b200 import ising, parameterization as p10n
  shape, T = (4, 6, 64), 7
  w = (0.2 * np.random.RandomState(5).normal(size=(2, 5))).astype(np.float32)
  sched = ising.IsingSchedule(p10n.Chebyshev(w[0] + np.float32([1.2, 0, 0, 0, 0])), p10n.Chebyshev(w[1]))
  times = np.linspace(0, 1, T, dtype=np.float32)
  params = sched(times)
  s0 = I.random_spins(shape, .5, J.PRNGKey(3))
  seeds = J.split(J.PRNGKey(8), 2)
  Jlt = P.Chebyshev(np.zeros(5), np.float64).jacobian(times.astype(np.float64))
  for loss in ('total_work', 'total_entropy_production'):
    g, _ = ising.estimate_gradient_rev(getattr(ising, loss))(sched, times, s0, seeds)
    got = np.concatenate([g.log_temp.weights.cpu().numpy(), g.field.weights.cpu().numpy()], -1)
    for r in range(2):
      o = I.closed_form_gradient(loss, np.asarray(params.log_temp), np.asarray(params.field), Jlt, Jlt, s0, seeds[r])
      assert rel(got[r], np.concatenate(o['grad'])) < 1e-4


@pytest.mark.parametrize('n,R,T', [(10000, 2, 5), (8, 3, 6), (12, 2, 6), (128, 2, 5), (256, 2, 4), (2052, 2, 4), (4224, 2, 3)])
def test_bit_packed_1d_kernel_is_bit_exact(cuda, n, R, T, monkeypatch):
  """ising_fast1d_kernel: periodic chains with N % 4 == 0 (10 000 is the reference's own 1-D test shape,
  tests/ising_test.py:71) run bit-packed in shared memory -- each colour plane stored as two halves of N/4 sites so
  that jax's counter partner p + N/2 is the same bit one half further, 2 x 3 site classes, one threefry block per TWO
  spin-updates.  The sizes cover a partly filled last word (10 000: 2500 = 78 words + 4 bits; 2052: one bit), one-word
  halves (8, 12, 128) and whole words (256, 4224).  Spin histories bit-exact against the oracle, summaries within the
  ulp bounds, and everything equal to the byte-per-spin generic kernel (NQ_ISING_NO_FAST1D=1)."""
  from noneq_opt_b200 import ising
  lt = np.log(np.linspace(1.6, 0.9, T).astype(np.float32))
  h = np.linspace(-0.3, 0.6, T).astype(np.float32)
  params = ising.IsingParameters(lt, h)
  s0 = I.random_spins((n,), .55, J.PRNGKey(21))
  seeds = J.split(J.PRNGKey(4), R)

  def run():
    final, (summary, states) = ising.simulate_ising(params, s0, seeds, return_states=True)
    torch.cuda.synchronize()
    return final.spins.clone(), [getattr(summary, f).clone() for f in FIELDS], states.spins.clone()
  fin, summ, states = run()
  assert tuple(fin.shape) == (R, n) and tuple(states.shape) == (R, T - 1, n)
  for r in range(R):
    fin_o, summ_o, states_o = I.simulate_ising(lt, h, s0, seeds[r], return_states=True)
    np.testing.assert_array_equal(fin[r].cpu().numpy(), fin_o)
    np.testing.assert_array_equal(states[r].cpu().numpy(), states_o)
    check_summary(np.stack([s[r].cpu().numpy() for s in summ]), np.stack([getattr(summ_o, f) for f in FIELDS]), n)
  monkeypatch.setenv('NQ_ISING_NO_FAST1D', '1')
  fin_g, summ_g, states_g = run()
  assert torch.equal(fin_g, fin) and torch.equal(states_g, states)
  for a, b in zip(summ_g, summ):
    assert torch.equal(a, b)


def test_bit_packed_1d_kernel_gradients_and_ensemble(cuda):
  """REINFORCE gradients of both losses on the reference's 1-D shape through the bit-packed path, against the
  oracle's closed-form gradient; and an ensemble of chains starting from one shared lattice."""
  from noneq_opt_b200 import ising, parameterization as p10n
  n, T = 10000, 9
  w = (0.2 * np.random.RandomState(9).normal(size=(2, 4))).astype(np.float32)
  sched = ising.IsingSchedule(p10n.Chebyshev(w[0] + np.float32([0.4, 0, 0, 0])), p10n.Chebyshev(w[1]))
  times = np.linspace(0, 1, T, dtype=np.float32)
  params = sched(times)
  s0 = I.random_spins((n,), .5, J.PRNGKey(2))
  seeds = J.split(J.PRNGKey(12), 3)
  Jlt = P.Chebyshev(np.zeros(4), np.float64).jacobian(times.astype(np.float64))
  for loss in ('total_work', 'total_entropy_production'):
    g, _ = ising.estimate_gradient_rev(getattr(ising, loss))(sched, times, s0, seeds)
    got = np.concatenate([g.log_temp.weights.cpu().numpy(), g.field.weights.cpu().numpy()], -1)
    for r in range(3):
      o = I.closed_form_gradient(loss, np.asarray(params.log_temp), np.asarray(params.field), Jlt, Jlt, s0, seeds[r])
      assert rel(got[r], np.concatenate(o['grad'])) < 1e-4
