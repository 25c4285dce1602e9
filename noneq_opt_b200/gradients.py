"""Utilities for computing gradients -- drop-in for `noneq_opt/gradients.py`.

`value_and_jacfwd(f)` (gradients.py:26-36) returns `f`'s value together with the Jacobian wrt its
FIRST argument (a pytree), one forward-mode tangent per scalar variable.  The reference builds it
from `jax.vmap(jax.jvp)`; without a tracing autodiff the same contract is met with
`torch.func.jvp` when `f` is written with torch ops, and the hot path does not go through here
at all (its Jacobians are the closed forms of `parameterization.*.jacobian` and the kernels'
per-step sensitivities).
"""
from __future__ import annotations

import numpy as np
import torch

from . import tree


def value_and_jacfwd(f):
  """Returns a function computing (value, jacobian wrt the first argument) of `f`.

  Leaves of the first argument must be torch tensors (or array-likes, converted to float64
  tensors); the Jacobian pytree mirrors the output structure with, for every output leaf, a pytree
  shaped like the first argument whose leaves have shape `out.shape + leaf.shape`."""
  def val_and_jac(*args, **kwargs):
    leaves, treedef = tree.tree_flatten(args[0])
    leaves = [l if isinstance(l, torch.Tensor) else torch.as_tensor(np.asarray(l), dtype=torch.float64)
              for l in leaves]

    def flat_f(*ls):
      out = f(tree.tree_unflatten(treedef, list(ls)), *args[1:], **kwargs)
      out_leaves, out_def = tree.tree_flatten(out)
      flat_f.out_def = out_def
      return tuple(out_leaves)

    value = None
    columns = []  # per input scalar: tuple of output tangents
    for li, leaf in enumerate(leaves):
      for k in range(leaf.numel()):
        tangents = [torch.zeros_like(l) for l in leaves]
        tangents[li].view(-1)[k] = 1
        value, tang = torch.func.jvp(flat_f, tuple(leaves), tuple(tangents))
        columns.append(tang)
    out_def = flat_f.out_def
    jac_leaves = []
    for oi in range(len(value)):
      stacked = torch.stack([col[oi] for col in columns], -1)   # out.shape + [n_inputs]
      pieces, o = [], 0
      for leaf in leaves:
        pieces.append(stacked[..., o:o + leaf.numel()].reshape(tuple(stacked.shape[:-1]) + tuple(leaf.shape)))
        o += leaf.numel()
      jac_leaves.append(tree.tree_unflatten(treedef, pieces))
    return tree.tree_unflatten(out_def, list(value)), tree.tree_unflatten(out_def, jac_leaves)
  return val_and_jac
