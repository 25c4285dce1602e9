"""The part of jax.numpy the reference uses, on float32 / int32 torch-CPU tensors."""
import numpy as _np
import torch as _torch

from ._core import Array, asarray, _torch_dtype

newaxis = None
pi = _np.pi
ndarray = Array


def _w(t):
  return t.as_subclass(Array) if isinstance(t, _torch.Tensor) else t


def array(x, dtype=None):
  return asarray(x, dtype)


def float32(x):
  return asarray(x, _np.float32)


def int32(x):
  return asarray(x, _np.int32)


def _unary(fn):
  return lambda x: _w(fn(asarray(x)))


exp, log, sqrt, abs, tanh, square = (_unary(f) for f in (_torch.exp, _torch.log, _torch.sqrt, _torch.abs, _torch.tanh,
                                                          _torch.square))
log1p, expm1, sign = _unary(_torch.log1p), _unary(_torch.expm1), _unary(_torch.sign)


def _axis_kw(axis):
  return {} if axis is None else {'dim': axis}


def sum(x, axis=None):
  return _w(_torch.sum(asarray(x), **_axis_kw(axis)))


def mean(x, axis=None):
  t = asarray(x)
  t = t if t.is_floating_point() else t.to(_torch.float32)
  return _w(_torch.mean(t, **_axis_kw(axis)))


def std(x, axis=None):
  return _w(_torch.std(asarray(x), unbiased=False, **_axis_kw(axis)))


def reshape(x, shape):
  return _w(asarray(x).reshape(tuple(shape) if hasattr(shape, '__iter__') else (shape,)))


def concatenate(xs, axis=0):
  return _w(_torch.cat([asarray(x) for x in xs], dim=axis))


def stack(xs, axis=0):
  return _w(_torch.stack([asarray(x) for x in xs], dim=axis))


def split(x, indices_or_sections, axis=0):
  t = asarray(x)
  if isinstance(indices_or_sections, int):
    return [_w(p) for p in _torch.chunk(t, indices_or_sections, dim=axis)]
  idx = [0] + [int(i) for i in indices_or_sections] + [t.shape[axis]]
  return [_w(t.narrow(axis, a, b - a)) for a, b in zip(idx[:-1], idx[1:])]


def arange(*a, dtype=None):
  return asarray(_np.arange(*a), dtype)


def linspace(start, stop, num=50, dtype=None):
  """jax.numpy.linspace: step_i = float32(i) / float32(num - 1); out = start (1 - step) + stop step in
  float32, with the end point appended exactly."""
  s, e = (_np.float32(float(asarray(v))) for v in (start, stop))
  num = int(num)
  if num == 1:
    return asarray(_np.array([s], _np.float32), dtype)
  step = (_np.arange(num - 1, dtype=_np.float32) / _np.float32(num - 1)).astype(_np.float32)
  out = (s * (_np.float32(1) - step) + e * step).astype(_np.float32)
  return asarray(_np.concatenate([out, [e]]).astype(_np.float32), dtype)


def ones(shape, dtype=_np.float32):
  return asarray(_np.ones(shape), dtype)


def zeros(shape, dtype=_np.float32):
  return asarray(_np.zeros(shape), dtype)


def ones_like(x):
  return _w(_torch.ones_like(asarray(x)))


def zeros_like(x):
  return _w(_torch.zeros_like(asarray(x)))


def roll(x, shift, axis=None):
  return _w(_torch.roll(asarray(x), shift, axis))


def where(c, a, b):
  return _w(_torch.where(asarray(c).bool(), asarray(a), asarray(b)))


def less(a, b):
  return _w(asarray(a) < asarray(b))


def maximum(a, b):
  return _w(_torch.maximum(asarray(a), asarray(b)))


def minimum(a, b):
  return _w(_torch.minimum(asarray(a), asarray(b)))


def meshgrid(*xs, indexing='xy'):
  return [_w(g) for g in _torch.meshgrid(*[asarray(x) for x in xs], indexing=indexing)]


def einsum(spec, *ops):
  ts = [asarray(o) for o in ops]
  ts = [t if t.is_floating_point() else t.to(_torch.float32) for t in ts]
  return _w(_torch.einsum(spec, *ts))


def interp(x, xp, fp, left=None, right=None):
  """numpy.interp semantics (clamped ends unless left / right are given), differentiable in fp."""
  x, xp, fp = asarray(x), asarray(xp), asarray(fp)
  idx = _torch.clamp(_torch.searchsorted(xp.contiguous(), x.contiguous(), right=True) - 1, 0, xp.shape[0] - 2)
  x0, x1, y0, y1 = xp[idx], xp[idx + 1], fp[idx], fp[idx + 1]
  t = _torch.clamp((x - x0) / (x1 - x0), 0., 1.)
  out = y0 + t * (y1 - y0)
  if left is not None:
    out = _torch.where(x < xp[0], asarray(left).to(out.dtype), out)
  if right is not None:
    out = _torch.where(x > xp[-1], asarray(right).to(out.dtype), out)
  return _w(out)


def float64(x):
  return asarray(x, _np.float32)   # x64 is disabled: jnp.float64(x) yields a float32


sin, cos = _unary(_torch.sin), _unary(_torch.cos)


def add(a, b):
  return _w(asarray(a) + asarray(b))


def multiply(a, b):
  return _w(asarray(a) * asarray(b))


def cumsum(x, axis=None):
  t = asarray(x)
  return _w(_torch.cumsum(t.reshape(-1) if axis is None else t, 0 if axis is None else axis))


def allclose(a, b, rtol=1e-5, atol=1e-8):
  return bool(_torch.allclose(asarray(a).to(_torch.float32), asarray(b).to(_torch.float32), rtol=rtol, atol=atol))
