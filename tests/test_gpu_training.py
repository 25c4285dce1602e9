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
