"""Multi-GPU plumbing for the ensemble path: one process per GPU, contiguous index slices, one
all-reduce of un-normalised float64 sums.

Trajectories / replicas are independent given their keys, and every rank derives the keys of its
own slice from the GLOBAL `jax.random.split(seed, B)` counter range (`random.split(..., first=,
count=)`), so the combined result does not depend on the number of ranks.  The only collective
is `all_reduce(SUM)` over a few dozen doubles (NCCL on the GPU box; gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def init_from_env(backend: str = 'nccl'):
  """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
  if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
    if backend == 'nccl':
      torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend=backend)
  return world()


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
  """(first, count) of the contiguous slice rank `rank` owns; slices differ by at most one.  Every rank must own at
  least one trajectory / replica: an empty shard would fail its kernel launch BEFORE the all-reduce and leave the
  other ranks waiting in the collective, so the mismatch is rejected on all ranks up front."""
  if int(n_total) < int(world_size):
    raise ValueError(f'{n_total} trajectories / replicas cannot be sharded over {world_size} ranks (every rank needs one)')
  base, rem = divmod(int(n_total), int(world_size))
  first = rank * base + min(rank, rem)
  return first, base + (1 if rank < rem else 0)


def allreduce_sums(tensors: Sequence[torch.Tensor]) -> Sequence[torch.Tensor]:
  """Sum each tensor over all ranks with ONE collective (flattened into one float64 buffer)."""
  rank, ws = world()
  if ws == 1:
    return list(tensors)
  flat = torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM)
  out, o = [], 0
  for t in tensors:
    out.append(flat[o:o + t.numel()].reshape(t.shape))
    o += t.numel()
  return out


def langevin_gradient_sharded(batch_size, energy_fn, r0_init, r0_final, Neq, shift, simulation_steps, dt,
                              temperature, mass, gamma):
  """estimate_gradient_boltzinit over all ranks: each rank simulates its slice of the global batch,
  then one all-reduce combines [grad sums | moments].  `init_pos` is this rank's slice
  `[count, 1, 1]` (or a scalar start broadcast to every trajectory).  Returns
  (grad [D] float32, dict(moments=float64[8] global, local=<rank-local outputs>))."""
  from . import barrier_crossing as bc
  from . import random as nqrandom

  def _estimate(coeffs, init_pos, seed):
    rank, ws = world()
    first, count = shard_range(batch_size, rank, ws)
    seeds = nqrandom.split(seed, batch_size, first=first, count=count)
    out = bc.gradient_terms(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift,
                            simulation_steps, dt, temperature, mass, gamma)
    if ws > 1:
      dist.all_reduce(out['sums_flat'], op=dist.ReduceOp.SUM)   # in place: grad_sums and moments are views of it
    grad_sums, moments = out['grad_sums'], out['moments']
    grad = ((grad_sums[0] + grad_sums[1]) / batch_size).to(torch.float32)
    return grad, dict(moments=moments, grad_sums=grad_sums, local=out)
  return _estimate


def ising_gradient_sharded(loss_function, batch_size, term: str = 'reinforce'):
  """The batch-mean gradient of `ising.estimate_gradient_rev(loss_function)` (ising.py:179-198, batched
  as in get_train_step :283-286) over all ranks: each rank runs its contiguous slice of the replicas
  keyed by rows of the GLOBAL `split(seed, batch_size)`, then one all-reduce combines
  [grad sums of both schedule components | n, sum(loss), sum(loss^2)].  Returns
  (mean grad pytree like the schedule (float32 numpy leaves), dict(loss_mean, loss_var, local_summary))."""
  from . import ising
  from . import random as nqrandom
  if term not in ('reinforce', 'logP', 'reparam'):
    raise ValueError(f'unknown gradient term {term!r}')

  def _estimate(schedule, times, initial_spins, seed):
    rank, ws = world()
    first, count = shard_range(batch_size, rank, ws)
    seeds = nqrandom.split(seed, batch_size, first=first, count=count)
    t = ising._gradient_terms(loss_function, schedule, times, initial_spins, seeds)
    pick = {'reinforce': tuple(a + b for a, b in zip(t['logP'], t['reparam'])), 'logP': t['logP'],
            'reparam': t['reparam']}[term]
    loss = t['loss'].to(torch.float64)
    mom = torch.stack([torch.tensor(float(count), dtype=torch.float64, device=loss.device), loss.sum(), (loss * loss).sum()])
    g_lt, g_h, mom = allreduce_sums([pick[0].to(torch.float64).sum(0), pick[1].to(torch.float64).sum(0), mom])
    grad = ising._pack_grad_like(schedule, (g_lt / batch_size)[None], (g_h / batch_size)[None], True)
    from . import tree
    grad = tree.tree_map(lambda x: x[0].to(torch.float32).cpu().numpy(), grad)
    mean = float(mom[1] / mom[0])
    return grad, dict(loss_mean=mean, loss_var=float(mom[2] / mom[0]) - mean * mean, local_summary=t['summary'])
  return _estimate


def ising_train_step_sharded(optimizer, initial_spins, batch_size: int, time_steps: int, loss_function=None,
                             max_grad=1e4):
  """`ising.get_train_step` (ising.py:263-290) with the batch sharded over the ranks: every rank applies
  the same clipped Adam update to its copy of the optimiser state (the all-reduced gradient is
  identical on all ranks), so the schedules stay in lock step without a broadcast."""
  import numpy as np
  from . import ising, optimizers
  loss_function = loss_function or ising.total_entropy_production
  estimator = ising_gradient_sharded(loss_function, batch_size)
  times = np.linspace(0, 1, time_steps, dtype=np.float32)
  update_fn, params_fn = optimizer[1], optimizer[2]

  def _train_step(opt_state, step, seed):
    grad, aux = estimator(params_fn(opt_state), times, initial_spins, seed)
    return update_fn(step, optimizers.clip_grads(grad, max_grad), opt_state), aux
  return _train_step
