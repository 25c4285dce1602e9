"""The optimiser pieces the reference's drivers take from `jax.example_libraries.optimizers`
(ising.py:6,288; run_ising_opt.py:30-31; run_barrier_crossing_opt.py:113-114): `adam`,
`exponential_decay`, `clip_grads`, with the same (init_fn, update_fn, params_fn) triple and the
same update rule, over the pytrees of `noneq_opt_b200.tree`.  Host-side: a schedule has a few
dozen variables.
"""
from __future__ import annotations

from typing import Callable, NamedTuple

import numpy as np

from . import tree


class Optimizer(NamedTuple):
  init_fn: Callable
  update_fn: Callable
  params_fn: Callable


class OptimizerState(NamedTuple):
  x: object
  m: object
  v: object


def constant(step_size):
  return lambda i: step_size


def exponential_decay(step_size, decay_steps, decay_rate):
  return lambda i: step_size * decay_rate ** (i / decay_steps)


def _schedule(step_size):
  return step_size if callable(step_size) else constant(step_size)


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):
  """m <- (1-b1) g + b1 m ; v <- (1-b2) g^2 + b2 v ; x <- x - lr(i) mhat / (sqrt(vhat) + eps)."""
  lr = _schedule(step_size)

  def init_fn(x0):
    f32 = tree.tree_map(lambda a: np.asarray(a, np.float32), x0)
    zeros = tree.tree_map(lambda a: np.zeros_like(np.asarray(a, np.float32)), x0)
    return OptimizerState(f32, zeros, tree.tree_map(np.copy, zeros))

  def update_fn(i, g, state):
    x, m, v = state
    g = tree.tree_map(lambda a: np.asarray(a.detach().cpu().numpy() if hasattr(a, 'detach') else a, np.float32), g)
    m = tree.tree_map(lambda gg, mm: ((1 - b1) * gg + b1 * mm).astype(np.float32), g, m)
    v = tree.tree_map(lambda gg, vv: ((1 - b2) * np.square(gg) + b2 * vv).astype(np.float32), g, v)
    step = np.float32(lr(i))

    def upd(xx, mm, vv):
      mhat = mm / np.float32(1 - b1 ** (i + 1))
      vhat = vv / np.float32(1 - b2 ** (i + 1))
      return (xx - step * mhat / (np.sqrt(vhat) + np.float32(eps))).astype(np.float32)
    return OptimizerState(tree.tree_map(upd, x, m, v), m, v)

  def params_fn(state):
    return state.x

  return Optimizer(init_fn, update_fn, params_fn)


def l2_norm(t):
  leaves = [np.asarray(a.detach().cpu().numpy() if hasattr(a, 'detach') else a, np.float64) for a in tree.tree_leaves(t)]
  return float(np.sqrt(sum(float(np.vdot(a, a)) for a in leaves)))


def clip_grads(grad_tree, max_norm):
  """Scale the whole tree so that its global l2 norm is at most `max_norm`."""
  norm = l2_norm(grad_tree)
  def clip(g):
    g = np.asarray(g.detach().cpu().numpy() if hasattr(g, 'detach') else g, np.float32)
    return g if norm < max_norm else (g * np.float32(max_norm / norm)).astype(np.float32)
  return tree.tree_map(clip, grad_tree)
