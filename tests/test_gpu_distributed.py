"""GPU: two ranks (gloo rendezvous on 127.0.0.1, both on cuda:0 when the box has one GPU) shard the
ensemble, derive their slices of the GLOBAL key split and all-reduce float64 sums; the result must
equal the single-rank estimator for both the Langevin and the Ising path."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
COEFFS = np.array([20., 19.95996, 0.3, -0.2, 0.1] + [0.] * 8, np.float32)
W_LT, W_H = np.array([0.3, -0.2, 0.1], np.float32), np.array([-0.4, 0.6, 0.2], np.float32)
B, S, R, T, L = 203, 150, 11, 6, 64


def _langevin(nqd, bc, space):
  pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
  return nqd.langevin_gradient_sharded(B, pot, 0., 40., 0, space.free()[1], S, 1e-8, KT, 1.0, KT / 1e7)


def _ising_inputs(ising, p10n):
  sched = ising.IsingSchedule(p10n.Chebyshev(W_LT), p10n.Chebyshev(W_H))
  return sched, np.linspace(0, 1, T, dtype=np.float32), -np.ones((L, L), np.float32)


def _worker(rank, ws, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
  import torch.distributed as dist
  from noneq_opt_b200 import barrier_crossing as bc, distributed as nqd, ising, parameterization as p10n, space
  torch.cuda.set_device(rank % torch.cuda.device_count())
  dist.init_process_group('gloo', rank=rank, world_size=ws)
  try:
    key = np.array([0, 9], np.uint32)
    grad, aux = _langevin(nqd, bc, space)(COEFFS, np.zeros((1, 1), np.float32), key)
    sched, times, s0 = _ising_inputs(ising, p10n)
    g_is, aux_is = nqd.ising_gradient_sharded(ising.total_entropy_production, R)(sched, times, s0, key)
    if rank == 0:
      out.put((grad.cpu().numpy(), aux['moments'].cpu().numpy(), g_is.log_temp.weights, g_is.field.weights,
               aux_is['loss_mean']))
  finally:
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(cuda):
  import torch.multiprocessing as mp
  from noneq_opt_b200 import barrier_crossing as bc, distributed as nqd, ising, parameterization as p10n, space
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  grad2, mom2, glt2, gh2, loss2 = q.get(timeout=300)
  for p in procs:
    p.join(timeout=300)
    assert p.exitcode == 0
  key = np.array([0, 9], np.uint32)
  grad1, aux1 = _langevin(nqd, bc, space)(COEFFS, np.zeros((1, 1), np.float32), key)
  np.testing.assert_allclose(mom2, aux1['moments'].cpu().numpy(), rtol=1e-12)
  np.testing.assert_allclose(grad2, grad1.cpu().numpy(), rtol=1e-6, atol=1e-9)
  sched, times, s0 = _ising_inputs(ising, p10n)
  g1, a1 = nqd.ising_gradient_sharded(ising.total_entropy_production, R)(sched, times, s0, key)
  np.testing.assert_allclose(glt2, g1.log_temp.weights, rtol=1e-6)
  np.testing.assert_allclose(gh2, g1.field.weights, rtol=1e-6)
  assert abs(loss2 - a1['loss_mean']) < 1e-9 * max(1.0, abs(loss2))
  # and the single-rank sharded estimator is the batch mean of the reference-shaped estimator
  from noneq_opt_b200 import random as nqr
  grads, _ = ising.estimate_gradient_rev(ising.total_entropy_production)(sched, times, s0, nqr.split(key, R))
  np.testing.assert_allclose(g1.field.weights, grads.field.weights.double().mean(0).cpu().numpy(), rtol=1e-6)


def test_sharded_train_step_moves_the_schedule(cuda):
  from noneq_opt_b200 import distributed as nqd, ising, optimizers, parameterization as p10n
  sched, _, s0 = _ising_inputs(ising, p10n)
  opt = optimizers.adam(1e-2)
  state = opt.init_fn(sched)
  step = nqd.ising_train_step_sharded(opt, s0, 8, T)
  state2, aux = step(state, 0, np.array([0, 3], np.uint32))
  new = opt.params_fn(state2)
  assert np.all(np.abs(np.asarray(new.field.weights) - W_H) > 1e-4) and np.isfinite(aux['loss_mean'])
