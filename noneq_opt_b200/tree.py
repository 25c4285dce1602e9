"""A small pytree protocol (the part of jax.tree_util the reference path relies on).

The reference registers every `Parameterization` as a jax pytree whose children are its
`variables` and whose aux data are its `constants` (parameterization.py:55-64), and maps over
NamedTuples such as `IsingSchedule` (ising.py:19-25, 64, 68, 230-233, 286).  Optimisers and
gradient estimators in this package use the same structure through these helpers.  Leaves are
numpy arrays / python scalars (host side; the schedule variables are a few dozen floats).
"""
from __future__ import annotations

import numpy as np


def _is_namedtuple(x):
  return isinstance(x, tuple) and hasattr(x, '_fields')


def _is_node(x):
  return hasattr(x, 'tree_flatten') and hasattr(type(x), 'tree_unflatten')


def tree_flatten(tree):
  """-> (leaves, treedef); treedef is an opaque rebuild closure description."""
  leaves = []

  def go(x):
    if _is_node(x):
      children, aux = x.tree_flatten()
      return ('node', type(x), aux, [go(c) for c in children])
    if _is_namedtuple(x):
      return ('namedtuple', type(x), None, [go(c) for c in x])
    if isinstance(x, (list, tuple)):
      return ('seq', type(x), None, [go(c) for c in x])
    if isinstance(x, dict):
      keys = sorted(x.keys())
      return ('dict', keys, None, [go(x[k]) for k in keys])
    if x is None:
      return ('none', None, None, [])
    leaves.append(x)
    return ('leaf', None, None, [])

  return leaves, go(tree)


def tree_unflatten(treedef, leaves):
  it = iter(leaves)

  def go(d):
    kind, typ, aux, kids = d
    if kind == 'leaf':
      return next(it)
    if kind == 'none':
      return None
    vals = [go(k) for k in kids]
    if kind == 'node':
      return typ.tree_unflatten(aux, vals)
    if kind == 'namedtuple':
      return typ(*vals)
    if kind == 'seq':
      return typ(vals)
    return dict(zip(typ, vals))

  return go(treedef)


def tree_leaves(tree):
  return tree_flatten(tree)[0]


def tree_map(f, tree, *rest):
  leaves, treedef = tree_flatten(tree)
  others = [tree_flatten(r)[0] for r in rest]
  for o in others:
    if len(o) != len(leaves):
      raise ValueError('tree_map: trees have different structures')
  return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


def ravel(tree):
  """Concatenate all leaves into one float vector; returns (vector, unravel_fn)."""
  leaves, treedef = tree_flatten(tree)
  arrs = [np.asarray(l) for l in leaves]
  shapes = [a.shape for a in arrs]
  sizes = [a.size for a in arrs]
  flat = np.concatenate([a.reshape(-1).astype(np.float64) for a in arrs]) if arrs else np.zeros(0)

  def unravel(v, dtype=np.float32):
    out, o = [], 0
    for sh, sz in zip(shapes, sizes):
      out.append(np.asarray(v[o:o + sz], dtype).reshape(sh))
      o += sz
    return tree_unflatten(treedef, out)

  return flat, unravel
