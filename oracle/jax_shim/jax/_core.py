"""Core of the shim: the Array type, pytrees, transformations."""
import functools

import numpy as np
import torch

torch.set_grad_enabled(True)
torch.set_flush_denormal(True)   # XLA:CPU runs with denormals flushed to zero (ising_test.py:36 expects exp(-100) == 0)


class _Size(int):
  """`x.size` is an int for jax arrays and a method for torch tensors: be both."""
  def __new__(cls, tensor):
    obj = int.__new__(cls, tensor.numel())
    obj._tensor = tensor
    return obj

  def __call__(self, *a, **k):
    return torch.Tensor.size(self._tensor, *a, **k)


class Array(torch.Tensor):
  """torch tensor with jax.Array manners: no in-place operators, int means give floats."""

  def __iadd__(self, o): return self + o
  def __isub__(self, o): return self - o
  def __imul__(self, o): return self * o
  def __itruediv__(self, o): return self / o

  def mean(self, *a, **k):
    t = self if self.is_floating_point() else self.to(torch.float32)
    return torch.Tensor.mean(t, *a, **k)

  def astype(self, dtype):
    return self.to(_torch_dtype(dtype))

  @property
  def size(self):
    return _Size(self)

  @property
  def at(self):
    raise NotImplementedError('x.at[...] is not part of the shim')


_DTYPES = {np.float32: torch.float32, np.float64: torch.float32, np.int32: torch.int32, np.int64: torch.int32,
           float: torch.float32, int: torch.int32, bool: torch.bool, np.bool_: torch.bool}


def _torch_dtype(dtype):
  if isinstance(dtype, torch.dtype):
    return dtype
  if dtype in _DTYPES:
    return _DTYPES[dtype]
  return _DTYPES[np.dtype(dtype).type]


def asarray(x, dtype=None):
  """x64 is disabled in jax by default: float64 -> float32, int64 -> int32.  uint32 arrays (PRNG keys)
  stay numpy."""
  if isinstance(x, np.ndarray) and x.dtype == np.uint32:
    return x
  if isinstance(x, torch.Tensor):
    t = x
  else:
    a = np.asarray(x)
    if a.dtype == np.float64:
      a = a.astype(np.float32)
    elif a.dtype == np.int64:
      a = a.astype(np.int32)
    t = torch.from_numpy(np.array(a, order="C"))   # (ascontiguousarray would turn 0-d into 1-d)
  if dtype is not None:
    t = t.to(_torch_dtype(dtype))
  elif t.dtype == torch.float64:
    t = t.to(torch.float32)
  elif t.dtype == torch.int64:
    t = t.to(torch.int32)
  return t.as_subclass(Array)


# ------------------------------------------------------------------------------ pytrees
_REGISTRY = set()


def register_pytree_node_class(cls):
  _REGISTRY.add(cls)
  return cls


def _is_namedtuple(x):
  return isinstance(x, tuple) and hasattr(x, '_fields')


def tree_flatten(tree):
  leaves = []

  def go(x):
    if type(x) in _REGISTRY:
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
  return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


# ------------------------------------------------------------------------------ transformations
def jit(fun=None, **unused):
  if fun is None:
    return lambda f: f
  return fun


def checkpoint(fun=None, **unused):
  return jit(fun)


def _stack(items, axis=0):
  first = items[0]
  if isinstance(first, np.ndarray):
    return np.stack(items, axis)
  return torch.stack([asarray(i) for i in items], dim=axis).as_subclass(Array)


def vmap(fun, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (list, tuple)) else [in_axes] * len(args)
    n = None
    for a, ax in zip(args, axes):
      if ax is not None:
        assert ax == 0, 'the shim maps over axis 0 only'
        n = len(tree_leaves(a)[0])
        break
    outs = []
    for i in range(n):
      call = [tree_map(lambda l: l[i], a) if ax is not None else a for a, ax in zip(args, axes)]
      outs.append(fun(*call))
    leaves0, treedef = tree_flatten(outs[0])
    cols = [[tree_flatten(o)[0][j] for o in outs] for j in range(len(leaves0))]
    return tree_unflatten(treedef, [_stack(c, out_axes) for c in cols])
  return mapped


def _require_grad(leaf):
  t = asarray(leaf)
  if not t.is_floating_point():
    return t
  if t.requires_grad:
    return t                      # already part of a graph: differentiate through it
  return t.detach().clone().requires_grad_(True).as_subclass(Array)


_DEPTH = [0]   # nesting of value_and_grad calls: inner ones keep their graph for the outer one


def _detach_tree(x):
  return tree_map(lambda l: l.detach().as_subclass(Array) if isinstance(l, torch.Tensor) else l, x)


def value_and_grad(fun, argnums=0, has_aux=False):
  def wrapped(*args, **kwargs):
    args = list(args)
    leaves, treedef = tree_flatten(args[argnums])
    leaves = [_require_grad(l) for l in leaves]
    args[argnums] = tree_unflatten(treedef, leaves)
    outermost = _DEPTH[0] == 0
    _DEPTH[0] += 1
    try:
      out = fun(*args, **kwargs)
    finally:
      _DEPTH[0] -= 1
    value, aux = out if has_aux else (out, None)
    diff = [l for l in leaves if l.is_floating_point()]
    if isinstance(value, torch.Tensor) and value.requires_grad:
      grads = torch.autograd.grad(value, diff, create_graph=not outermost, allow_unused=True)
    else:
      grads = [None] * len(diff)
    it = iter(grads)
    full = []
    for l in leaves:
      g = next(it) if l.is_floating_point() else None
      full.append((torch.zeros_like(l) if g is None else g).as_subclass(Array))
    grad_tree = tree_unflatten(treedef, full)
    if outermost:
      value, aux, grad_tree = _detach_tree(value), _detach_tree(aux), _detach_tree(grad_tree)
    return ((value, aux), grad_tree) if has_aux else (value, grad_tree)
  return wrapped


def grad(fun, argnums=0, has_aux=False):
  vg = value_and_grad(fun, argnums, has_aux)

  def wrapped(*args, **kwargs):
    (value, g) = vg(*args, **kwargs)
    return (g, value[1]) if has_aux else g
  return wrapped


def jvp(fun, primals, tangents):
  """Forward-mode derivative along `tangents` (pytrees like `primals`): torch forward AD on the leaves."""
  import torch.autograd.forward_ad as fwad
  with fwad.dual_level():
    duals = []
    for p, t in zip(primals, tangents):
      pl, treedef = tree_flatten(p)
      tl = tree_leaves(t)
      dl = []
      for a, b in zip(pl, tl):
        a = asarray(a)
        if a.is_floating_point():
          dl.append(fwad.make_dual(a.detach(), asarray(b).to(a.dtype).reshape(a.shape)).as_subclass(Array))
        else:
          dl.append(a)
      duals.append(tree_unflatten(treedef, dl))
    out = fun(*duals)
    leaves, outdef = tree_flatten(out)
    prim, tang = [], []
    for l in leaves:
      if isinstance(l, torch.Tensor):
        u = fwad.unpack_dual(l)
        prim.append(u.primal.detach().as_subclass(Array))
        tang.append((torch.zeros_like(u.primal) if u.tangent is None else u.tangent.detach()).as_subclass(Array))
      else:
        prim.append(l)
        tang.append(l)
    return tree_unflatten(outdef, prim), tree_unflatten(outdef, tang)


def stop_gradient(x):
  return tree_map(lambda l: l.detach().as_subclass(Array) if isinstance(l, torch.Tensor) else l, x)


def scan(f, init, xs=None, length=None, reverse=False):
  if xs is None:
    n = length
  else:
    n = len(tree_leaves(xs)[0])
  carry, ys = init, []
  order = range(n - 1, -1, -1) if reverse else range(n)
  for i in order:
    x = None if xs is None else tree_map(lambda l: l[i], xs)
    carry, y = f(carry, x)
    ys.append(y)
  if reverse:
    ys = ys[::-1]
  if not ys:
    return carry, None
  leaves0, treedef = tree_flatten(ys[0])
  cols = [[tree_flatten(y)[0][j] for y in ys] for j in range(len(leaves0))]
  return carry, tree_unflatten(treedef, [_stack(c) for c in cols])
