"""Device-resident training step (noneq_opt_b200.training.LangevinTrainer): the CUDA
graph replay must follow the same trajectory through coefficient space as the step-by-step path
(estimate_gradient_boltzinit + optimizers.adam on the host) the reference's driver loop maps to."""
import numpy as np
import pytest
import torch

from oracle import jaxlike as J

pytestmark = pytest.mark.gpu


def _host_loop(bc, optimizers, nqr, space, energy, coeffs0, key, B, S, Neq, dt, kT, gamma, eqm, steps, r0i, r0f,
               term='reinforce', max_grad=None, lr=(0.1, 8, 0.5)):
  _, shift = space.free()
  opt = optimizers.adam(optimizers.exponential_decay(*lr))
  state = opt.init_fn(coeffs0)
  factory = {'reinforce': bc.estimate_gradient_boltzinit, 'logP': bc.estimate_gradient_boltzinit_logP,
             'reparam': bc.estimate_gradient_boltzinit_reparam}[term]
  grad_fn = factory(B, energy, r0i, r0f, Neq, shift, S, dt, kT, 1.0, gamma)
  eqm_fn = None
  if eqm:
    eqm_fn = bc.mapped_brownian_eqm(B, energy, np.full((1, 1), r0i, np.float32), r0i, shift, eqm['sim_length'],
                                    eqm['dt'], eqm['kT'], 1.0, eqm['gamma'], reference_quirk=False)
  hist, works = [np.asarray(opt.params_fn(state)).copy()], []
  for j in range(steps):
    key, k1, k2 = nqr.split_host(key, 3)
    x0 = eqm_fn(k1) if eqm_fn else np.full((B, 1, 1), r0i, np.float32)
    grad, (_, (_, _, W)) = grad_fn(opt.params_fn(state), x0, k2)
    if max_grad:
      grad = optimizers.clip_grads(grad, max_grad)
    state = opt.update_fn(j, grad, state)
    hist.append(np.asarray(opt.params_fn(state)).copy())
    works.append(float(W.double().mean()))
  return np.stack(hist), np.array(works), key


@pytest.mark.parametrize('case', ['spring_eqm', 'biomolecule_clip', 'spring_eager'])
def test_graph_replay_follows_the_host_loop(cuda, case):
  from noneq_opt_b200 import barrier_crossing as bc, optimizers, random as nqr, space, training
  _, shift = space.free()
  if case.startswith('spring'):
    energy, B, S, Neq, dt, kT, gamma = bc.V_spring(1.0), 512, 200, 10, 5e-3, 1.0, 1.0
    r0i, r0f = 5., 10.
    coeffs0 = bc.fit_linear_to_cheby(1.0, 5e-3, r0i, r0f, 12)
    eqm = dict(sim_length=50, dt=5e-3, kT=1.0, gamma=1.0)
    term, max_grad = 'reinforce', None
  else:
    KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
    energy, B, S, Neq, dt, kT, gamma = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA), 384, 300, 0, 1e-8, KT, KT / 1e7
    r0i, r0f = 0., 40.
    coeffs0 = np.array([20., 19.95996] + [0.] * 11, np.float32)
    eqm, term, max_grad = None, 'logP', 0.05
  steps = 5
  key = J.PRNGKey(7)
  hist_h, works_h, key_h = _host_loop(bc, optimizers, nqr, space, energy, coeffs0, key, B, S, Neq, dt, kT, gamma, eqm,
                                      steps, r0i, r0f, term, max_grad)
  tr = training.LangevinTrainer(B, energy, coeffs0, r0i, r0f, Neq, shift, S, dt, kT, 1.0, gamma, key, eqm=eqm,
                                term=term, step_size=0.1, decay_steps=8, decay_rate=0.5, max_grad=max_grad,
                                use_graph=(case != 'spring_eager'))
  out = tr.run(steps)
  torch.cuda.synchronize()
  assert tr.launches_per_step == (8 if eqm else 6)
  hist_d = out['coeffs'].cpu().numpy()
  np.testing.assert_array_equal(hist_d[0], hist_h[0])
  scale = np.abs(hist_h).max()
  assert np.abs(hist_d - hist_h).max() / scale < 2e-6, np.abs(hist_d - hist_h).max() / scale
  np.testing.assert_allclose(out['mean_work'].cpu().numpy(), works_h, rtol=2e-6)
  np.testing.assert_array_equal(nqr.keys_to_numpy(tr.key).reshape(2), np.asarray(key_h, np.uint32).reshape(2))
  assert int(tr.step_count) == steps


def test_parameterization_protocol_and_works_history(cuda):
  """A ConstrainEndpoints(PiecewiseLinear) protocol goes through its affine form (Jacobian + offset);
  the first step's gradient must equal the step-by-step estimator's."""
  from noneq_opt_b200 import barrier_crossing as bc, parameterization as p10n, space, training
  _, shift = space.free()
  B, S = 256, 160
  knots = (np.linspace(5, 10, 8)[1:-1] + 0.1 * np.sin(np.arange(6))).astype(np.float32)
  proto = p10n.PiecewiseLinear(knots, 0., 1., 5., 10.)
  key = J.PRNGKey(3)
  tr = training.LangevinTrainer(B, bc.V_spring(1.0), proto, 5., 10., 5, shift, S, 5e-3, 1.0, 1.0, 1.0, key,
                                step_size=1e-2)
  out = tr.run(2, record_works=True)
  from noneq_opt_b200 import random as nqr
  _, _, k2 = nqr.split_host(key, 3)
  est = bc.estimate_gradient_boltzinit(B, bc.V_spring(1.0), 5., 10., 5, shift, S, 5e-3, 1.0, 1.0, 1.0)
  grad, (_, (_, _, W)) = est(proto, np.full((B, 1, 1), 5., np.float32), k2)
  # Adam's first step moves every variable by -lr * sign(g)
  step0 = (out['coeffs'][1] - out['coeffs'][0]).cpu().numpy()
  np.testing.assert_allclose(step0, -1e-2 * np.sign(grad.cpu().numpy()), rtol=1e-4)
  np.testing.assert_allclose(out['works'][0].cpu().numpy(), W.cpu().numpy(), rtol=1e-5, atol=1e-5)
  assert out['works'].shape == (2, B)


def test_graph_capture_of_the_persistent_scheduler(cuda):
  """45 000 trajectories take the persistent-scheduler launch (ring + carry in the workspace, an init
  kernel before the main one): the captured graph must replay it exactly like the eager sequence."""
  from noneq_opt_b200 import barrier_crossing as bc, space, training
  _, shift = space.free()
  B, S = 45_000, 300
  coeffs0 = bc.fit_linear_to_cheby(1.0, 1.0 / S, 5., 10., 12)
  key = J.PRNGKey(11)
  runs = []
  for graph in (False, True):
    tr = training.LangevinTrainer(B, bc.V_spring(1.0), coeffs0, 5., 10., 0, shift, S, 1.0 / S, 1.0, 1.0, 1.0, key,
                                  step_size=0.05, use_graph=graph)
    out = tr.run(3)
    torch.cuda.synchronize()
    runs.append((out['coeffs'].cpu().numpy(), out['mean_work'].cpu().numpy(), tr.launches_per_step))
  np.testing.assert_array_equal(runs[0][0], runs[1][0])
  np.testing.assert_array_equal(runs[0][1], runs[1][1])
  assert runs[0][2] == runs[1][2] == 7      # split3, split(B), schedule, ring init, persistent kernel, finalize, adam
  assert np.all(np.diff(runs[1][1]) < 0)     # three Adam steps on the linear protocol lower the mean work


# ================================================================================ Ising (training.IsingTrainer)
def _ising_rel(a, b):
  return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('ndim', [1, 2, 3])
def test_device_class_tables_are_bit_equal_to_the_host_tables(cuda, ndim):
  """nq_ising_class_tables (thresholds and log-sigmoid tables built on the device from device schedule tables)
  against ising.class_tables (numpy) over a dense (log-temperature, field) grid: same float32 operation order,
  binary64 transcendentals rounded once -- every integer threshold and every float32-valued table entry (logit,
  log_sigmoid(-+logit)) must be bit-equal; the binary64 sigmoid row (used for the derivative partials only) may
  differ by the last bits of the two libraries' binary64 exp (CUDA's is not correctly rounded): <= 4 ulp."""
  import ctypes as C
  from noneq_opt_b200 import _lib, ising
  rs = np.random.RandomState(ndim)
  lts = np.linspace(-1.6, 3.2, 97, dtype=np.float32)
  hs = np.linspace(-3.5, 3.5, 83, dtype=np.float32)
  lt = np.concatenate([[0.3], np.repeat(lts, len(hs)), rs.uniform(-2, 3, 3000)]).astype(np.float32)
  h = np.concatenate([[0.0], np.tile(hs, len(lts)), rs.uniform(-4, 4, 3000)]).astype(np.float32)
  T = len(lt)
  thr_h, lut_h = ising.class_tables(lt, h, ndim)
  lt_d, h_d = torch.from_numpy(lt).cuda(), torch.from_numpy(h).cuda()
  thr_d = torch.empty((T - 1, 16), dtype=torch.int32, device='cuda')
  lut_d = torch.empty((T - 1, 4, 16), dtype=torch.float64, device='cuda')
  _lib.check(_lib.lib().nq_ising_class_tables(_lib.ptr(lt_d), _lib.ptr(h_d), T, ndim, _lib.ptr(thr_d), _lib.ptr(lut_d),
                                              _lib.stream_arg()))
  np.testing.assert_array_equal(thr_d.cpu().numpy().view(np.uint32), thr_h)
  lut_g = lut_d.cpu().numpy()
  np.testing.assert_array_equal(lut_g[:, :3].view(np.uint64), lut_h[:, :3].view(np.uint64))
  ulp = np.abs(lut_g[:, 3].view(np.int64) - lut_h[:, 3].view(np.int64))
  assert ulp.max() <= 4, ulp.max()


def test_ising_trainer_follows_the_reference_source_through_graph_replay(cuda):
  """tests/golden/reference_train_step.npz -- three steps of the reference's own get_train_step (ising.py:263-290),
  generated by make_reference_golden.py -- followed by training.IsingTrainer: one CUDA-graph replay per step,
  no host arithmetic and no synchronisation inside a step."""
  import os
  from noneq_opt_b200 import ising, parameterization as p10n, training
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_train_step.npz'))
  c = {k.split('/')[1]: g[k] for k in g.files if k.startswith('train/')}
  sched = ising.IsingSchedule(p10n.Chebyshev(c['w_lt']), p10n.Chebyshev(c['w_h']))
  tr = training.IsingTrainer(sched, -np.ones((16, 16), np.float32), int(c['B']), int(c['T']), c['seed'],
                             loss_function=ising.total_entropy_production, step_size=float(c['lr']), max_grad=1e4,
                             seed_source='split')
  for j in range(3):
    tr.step()
    got = tr.variables.cpu().numpy().reshape(2, -1)
    assert _ising_rel(got, c['weights'][j]) < 2e-5, (j, _ising_rel(got, c['weights'][j]))
    ep = tr.summary.entropy_production.cpu().numpy()
    assert np.abs(ep - c['entropy_production'][j]).max() <= 2e-4 * max(1.0, np.abs(c['entropy_production'][j]).max())
  assert tr._graph is not None and tr.launches_per_step >= 8
  assert int(tr.step_count.item()) == 3


@pytest.mark.parametrize('loss,term,use_graph', [('total_entropy_production', 'reinforce', True),
                                                 ('total_work', 'reinforce', True),
                                                 ('total_entropy_production', 'logP', False)])
def test_ising_trainer_equals_the_host_train_step(cuda, loss, term, use_graph):
  """The drivers' schedule composition AddBaseline(ConstrainEndpoints(Chebyshev(w), 0, 0), baseline)
  (run_ising_opt.py:45-66), seed_stream keys, clip_grads(max_grad = 1), adam(exponential_decay(0.1, n, 0.01)):
  the device-resident step against ising.get_train_step* on the host.  Schedule tables: bit-equal (same float32
  order); variables after every step: equal to float32 rounding of the batch mean."""
  from noneq_opt_b200 import ising, optimizers, parameterization as p10n, training
  L, R, T, n, D = 32, 24, 21, 4, 8
  rs = np.random.RandomState(3)
  w0 = (0.05 * rs.normal(size=(2, D))).astype(np.float32)
  sched = ising.IsingSchedule(
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w0[0]), 0., 0.), ising.log_temp_baseline(min_temp=0.65)),
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w0[1]), 0., 0.), ising.field_baseline()))
  spins0 = -np.ones((L, L), np.float32)
  loss_fn = getattr(ising, loss)
  opt = optimizers.adam(optimizers.exponential_decay(0.1, n, 0.01))
  state = opt.init_fn(sched)
  factory = {'reinforce': ising.get_train_step_REINFORCE, 'logP': ising.get_train_step_logP}[term]
  step_fn = factory(opt, spins0, R, T, loss_fn, 'rev', max_grad=1.0)
  stream = ising.seed_stream(0)
  tr = training.IsingTrainer(sched, spins0, R, T, J.PRNGKey(0), loss_function=loss_fn, term=term, max_grad=1.0,
                             step_size=0.1, decay_steps=n, decay_rate=0.01, use_graph=use_graph)
  times = np.linspace(0, 1, T, dtype=np.float32)
  for j in range(n):
    host_tables = opt.params_fn(state)(times)
    state, summary, mean_grad = step_fn(state, j, next(stream))
    tr.step()
    np.testing.assert_array_equal(tr.log_temp.cpu().numpy(), np.asarray(host_tables.log_temp))   # tables of THIS step
    np.testing.assert_array_equal(tr.field.cpu().numpy(), np.asarray(host_tables.field))
    for f in ising.IsingSummary._fields:
      assert torch.equal(getattr(tr.summary, f), getattr(summary, f)), f
    p = opt.params_fn(state)
    want = np.concatenate([np.asarray(p.log_temp.wrapped.wrapped.weights), np.asarray(p.field.wrapped.wrapped.weights)])
    got = tr.variables.cpu().numpy()
    assert _ising_rel(got, want) < 5e-6, (j, _ising_rel(got, want))
    mg = np.concatenate([np.asarray(mean_grad.log_temp.wrapped.wrapped.weights),
                         np.asarray(mean_grad.field.wrapped.wrapped.weights)])
    assert _ising_rel(tr.grad.cpu().numpy(), mg) < 5e-6
  cur = tr.current_schedule()
  np.testing.assert_array_equal(np.asarray(cur.log_temp.wrapped.wrapped.weights), tr.variables[:D].cpu().numpy())


def test_ising_trainer_rejects_what_it_cannot_run(cuda):
  from noneq_opt_b200 import ising, parameterization as p10n, training
  sched = ising.IsingSchedule(p10n.Chebyshev(np.zeros(4, np.float32)), p10n.Chebyshev(np.zeros(4, np.float32)))
  with pytest.raises(NotImplementedError):
    training.IsingTrainer(sched, -np.ones((8, 8), np.float32), 2, 5, J.PRNGKey(0), loss_function=lambda a, b, s: s.work.sum())
  with pytest.raises(ValueError):
    training.IsingTrainer(sched, -np.ones((7, 8), np.float32), 2, 5, J.PRNGKey(0))
