"""The N>1 host logic on CPU: world_size-2 gloo ranks shard a batch contiguously and one
all-reduce of un-normalised float64 sums reproduces the single-rank result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from noneq_opt_b200 import distributed as nqd


def test_shard_range_partitions_exactly():
  import pytest
  with pytest.raises(ValueError):      # an empty shard would hang the other ranks in the all-reduce
    nqd.shard_range(7, 0, 8)
  for n, ws in [(10, 3), (100000, 8), (8, 8), (10_000_000, 8), (5, 1)]:
    spans = [nqd.shard_range(n, r, ws) for r in range(ws)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == n
    for (f0, c0), (f1, _) in zip(spans, spans[1:]):
      assert f0 + c0 == f1
    assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, ws, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
  dist.init_process_group('gloo', rank=rank, world_size=ws)
  try:
    rs = np.random.RandomState(0)
    per_traj = rs.normal(size=(1001, 13))                # stand-in for per-trajectory grad terms
    first, count = nqd.shard_range(1001, rank, ws)
    local = per_traj[first:first + count]
    grad_sums = torch.tensor(np.stack([local.sum(0), (local ** 2).sum(0)]))
    moments = torch.tensor([float(count), local[:, 0].sum(), 0, 0, 0, 0, 0, 0], dtype=torch.float64)
    g, m = nqd.allreduce_sums([grad_sums, moments])
    if rank == 0:
      out.put((g.numpy(), m.numpy()))
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_single_rank():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  g, m = q.get(timeout=120)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  per_traj = np.random.RandomState(0).normal(size=(1001, 13))
  np.testing.assert_allclose(g[0], per_traj.sum(0), rtol=1e-12)
  np.testing.assert_allclose(g[1], (per_traj ** 2).sum(0), rtol=1e-12)
  assert m[0] == 1001


def test_single_process_is_identity():
  t = [torch.arange(4, dtype=torch.float64)]
  assert nqd.allreduce_sums(t)[0] is t[0]
  assert nqd.world() == (0, 1)
