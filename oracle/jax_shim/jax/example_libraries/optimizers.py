"""jax.example_libraries.optimizers: the pieces the reference's train steps and tests use (adam,
exponential_decay, clip_grads, l2_norm) over the shim's pytrees."""
import collections

import torch

from .._core import asarray, tree_leaves, tree_map

Optimizer = collections.namedtuple('Optimizer', 'init_fn update_fn params_fn')
OptimizerState = collections.namedtuple('OptimizerState', 'x m v')


def constant(step_size):
  return lambda i: step_size


def exponential_decay(step_size, decay_steps, decay_rate):
  return lambda i: step_size * decay_rate ** (i / decay_steps)


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):
  lr = step_size if callable(step_size) else constant(step_size)

  def init(x0):
    x0 = tree_map(asarray, x0)
    return OptimizerState(x0, tree_map(torch.zeros_like, x0), tree_map(torch.zeros_like, x0))

  def update(i, g, state):
    x, m, v = state
    m = tree_map(lambda gg, mm: (1 - b1) * gg + b1 * mm, g, m)
    v = tree_map(lambda gg, vv: (1 - b2) * gg * gg + b2 * vv, g, v)

    def step(xx, mm, vv):
      mhat, vhat = mm / (1 - b1 ** (i + 1)), vv / (1 - b2 ** (i + 1))
      return (xx - lr(i) * mhat / (torch.sqrt(vhat) + eps)).detach()
    return OptimizerState(tree_map(step, x, m, v), tree_map(lambda t: t.detach(), m), tree_map(lambda t: t.detach(), v))

  def get_params(state):
    return state.x
  return Optimizer(init, update, get_params)


def l2_norm(tree):
  return torch.sqrt(sum(torch.vdot(l.reshape(-1), l.reshape(-1)) for l in tree_leaves(tree)))


def clip_grads(grad_tree, max_norm):
  norm = l2_norm(grad_tree)
  return tree_map(lambda g: torch.where(norm < max_norm, g, g * (max_norm / norm)), grad_tree)
