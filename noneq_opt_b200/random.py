"""jax.random-compatible key handling backed by the sm_100a threefry kernels.

The reference draws every random number through `jax.random` (barrier_crossing.py:79,133,221;
ising.py:72,132,161,283).  Keys here are raw `uint32[2]` arrays exactly as `jax.random.PRNGKey`
produces them, and `split` / `bits` / `uniform` / `normal` reproduce jax's ORIGINAL
(non-partitionable) threefry2x32 counter layout bit for bit, so code ported from the reference
sees the same streams.  All generation happens on the GPU (csrc/core.cu); results are CUDA
tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def PRNGKey(seed: int) -> np.ndarray:
  """jax.random.PRNGKey: [seed >> 32, seed & 0xffffffff] as uint32[2] (host array)."""
  seed = int(seed)
  return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _host_key(key) -> np.ndarray:
  if isinstance(key, torch.Tensor):
    key = key.detach().cpu().numpy()
  key = np.asarray(key)
  if key.shape != (2,):
    raise ValueError(f'a PRNG key must have shape (2,), got {key.shape}')
  return key.astype(np.int64).astype(np.uint32) if key.dtype != np.uint32 else key


def as_key_tensor(keys, device=None) -> torch.Tensor:
  """[..., 2] keys -> int32-viewed CUDA tensor (torch has no native uint32 arithmetic)."""
  device = device or _lib.require_cuda()
  if isinstance(keys, torch.Tensor):
    t = keys.to(device)
    if t.dtype == torch.uint32:
      t = t.view(torch.int32)
    return t.to(torch.int32).contiguous()
  a = np.ascontiguousarray(np.asarray(keys).astype(np.int64).astype(np.uint32)) \
      if np.asarray(keys).dtype != np.uint32 else np.ascontiguousarray(keys)
  return torch.from_numpy(a.view(np.int32)).to(device)


def keys_to_numpy(t: torch.Tensor) -> np.ndarray:
  return t.detach().cpu().numpy().view(np.uint32)


def split(key, num: int = 2, first: int = 0, count: int | None = None) -> torch.Tensor:
  """jax.random.split(key, num)[first:first+count] as an int32-viewed CUDA tensor [count, 2].

  `first`/`count` let a rank derive only its own contiguous slice of the GLOBAL split, which is
  what makes sharded results independent of the number of GPUs."""
  device = _lib.require_cuda()
  count = num - first if count is None else count
  out = torch.empty((count, 2), dtype=torch.int32, device=device)
  _lib.check(_lib.lib().nq_random_split(_lib.key_arg(_host_key(key)), num, first, count,
                                        _lib.ptr(out), _lib.stream_arg()))
  return out


def split_host(key, num: int = 2) -> np.ndarray:
  """split() copied back to the host as uint32[num, 2] (for small driver-level splits)."""
  return keys_to_numpy(split(key, num))


def bits(key, shape) -> torch.Tensor:
  device = _lib.require_cuda()
  shape = (shape,) if np.isscalar(shape) else tuple(shape)
  size = int(np.prod(shape)) if shape else 1
  out = torch.empty(size, dtype=torch.int32, device=device)
  _lib.check(_lib.lib().nq_random_bits(_lib.key_arg(_host_key(key)), size, 0, size,
                                       _lib.ptr(out), _lib.stream_arg()))
  return out.reshape(shape)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> torch.Tensor:
  device = _lib.require_cuda()
  shape = (shape,) if np.isscalar(shape) else tuple(shape)
  size = int(np.prod(shape)) if shape else 1
  out = torch.empty(size, dtype=torch.float32, device=device)
  _lib.check(_lib.lib().nq_random_uniform(_lib.key_arg(_host_key(key)), size, 0, size,
                                          float(minval), float(maxval), _lib.ptr(out), _lib.stream_arg()))
  return out.reshape(shape)


def normal(key, shape=()) -> torch.Tensor:
  device = _lib.require_cuda()
  shape = (shape,) if np.isscalar(shape) else tuple(shape)
  size = int(np.prod(shape)) if shape else 1
  out = torch.empty(size, dtype=torch.float32, device=device)
  _lib.check(_lib.lib().nq_random_normal(_lib.key_arg(_host_key(key)), size, 0, size,
                                         _lib.ptr(out), _lib.stream_arg()))
  return out.reshape(shape)
