"""Control-flow helpers -- drop-in for `noneq_opt/control_flow.py`.

`checkpoint_scan(f, init, xs, checkpoint_every)` (control_flow.py:16-31) is `lax.scan` with
`jax.checkpoint` around every block of `checkpoint_every` steps: the results are those of a plain
scan, the memory of reverse-mode differentiation drops to one carry per block.  Here the scan is
a host loop over pytrees (used for small generic scans); the memory/recompute trade-off the
reference buys with it lives in the kernels -- `nq_langevin_forward_ckpt` / `nq_langevin_replay`
store one (x, rng) carry per block and regenerate the block in the gradient pass.
"""
from __future__ import annotations

import numpy as np

from . import tree


def check_divides(length: int, checkpoint_every: int):
  outer, residual = divmod(int(length), int(checkpoint_every))
  if residual:
    raise ValueError('`checkpoint_every` must evenly divide the length of `xs`. '
                     f'Got {checkpoint_every} and {length}.')
  return outer


def scan(f, init, xs, length=None):
  """Host restatement of jax.lax.scan over pytrees of arrays (leading axis = time)."""
  leaves = tree.tree_leaves(xs)
  n = int(np.shape(leaves[0])[0]) if leaves else int(length)
  carry, ys = init, []
  for i in range(n):
    x = tree.tree_map(lambda a: a[i], xs) if leaves else None
    carry, y = f(carry, x)
    ys.append(y)
  stacked = tree.tree_map(lambda *a: np.stack([np.asarray(v) for v in a]), *ys) if ys else None
  return carry, stacked


def checkpoint_scan(f, init, xs, checkpoint_every):
  """Replicates `scan` but validates (and honours) the block structure of gradient checkpointing."""
  leaves = tree.tree_leaves(xs)
  length = int(np.shape(leaves[0])[0])
  outer = check_divides(length, checkpoint_every)
  carry, blocks = init, []
  for b in range(outer):
    sl = slice(b * checkpoint_every, (b + 1) * checkpoint_every)
    carry, ys = scan(f, carry, tree.tree_map(lambda a: a[sl], xs))
    blocks.append(ys)
  return carry, tree.tree_map(lambda *a: np.concatenate(a, 0), *blocks)
