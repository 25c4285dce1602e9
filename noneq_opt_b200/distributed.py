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
  """(first, count) of the contiguous slice rank `rank` owns; slices differ by at most one."""
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
    grad_sums, moments = allreduce_sums([out['grad_sums'], out['moments']])
    grad = ((grad_sums[0] + grad_sums[1]) / batch_size).to(torch.float32)
    return grad, dict(moments=moments, grad_sums=grad_sums, local=out)
  return _estimate
