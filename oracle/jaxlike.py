"""Oracle (test infrastructure): the third-party semantics the reference path relies on.

The reference (`/root/reference/noneq_opt`) draws all of its randomness through
`jax.random` (un-vendored, un-pinned: setup.py:9-20), `tfp.distributions.Normal`
and `distrax.Bernoulli`.  None of them is installable in this image, so this file
restates their *published algorithms* in numpy:

  * threefry2x32-20: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
    (SC'11) -- Random123 `threefry2x32_R(20, ...)`; jax/_src/prng.py
    `threefry2x32_p` / `_threefry2x32_lowering` follow it verbatim.
  * array form, `split`, `random_bits`: jax/_src/prng.py `threefry_2x32`,
    `threefry_split`, `threefry_random_bits` with the ORIGINAL
    (`jax_threefry_partitionable=False`) counter layout, which is the default of
    every jax release the reference could run on (jax 0.2.25 .. 0.4.x).
  * uniform / normal: jax/_src/random.py `_uniform`, `_normal_real`.
  * erf_inv (f32): XLA `ErfInv32` (xla/client/lib/math.cc), M. Giles'
    single-precision polynomial.
  * Normal.sample/log_prob: tensorflow_probability/python/distributions/normal.py.
  * Bernoulli.sample/log_prob: distrax/_src/distributions/bernoulli.py.

Reference call sites: barrier_crossing.py:79,88,90,133,221; ising.py:72,115-117,
123-124,132,161,283.

Pinned by tests/test_oracle_rng.py (Random123 KATs + jax doc values).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
F32 = np.float32

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x, r):
  return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
  """threefry2x32-20 on (broadcastable) uint32 arrays. Returns (y0, y1)."""
  with np.errstate(over='ignore'):
    k0 = np.asarray(k0, U32)
    k1 = np.asarray(k1, U32)
    x0 = np.asarray(x0, U32).copy()
    x1 = np.asarray(x1, U32).copy()
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    x0 = x0 + ks[0]
    x1 = x1 + ks[1]
    for i in range(5):
      for r in _ROT[i % 2]:
        x0 = x0 + x1
        x1 = _rotl(x1, r)
        x1 = x1 ^ x0
      x0 = x0 + ks[(i + 1) % 3]
      x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
  return x0, x1


def threefry_2x32(key, counts):
  """jax/_src/prng.py `threefry_2x32(keypair, count)`: hash a flat count array.

  Odd lengths are padded with one zero; the first half of the array is x0, the
  second half x1; the output is concat(y0, y1) truncated to the input length.
  """
  key = np.asarray(key, U32)
  counts = np.asarray(counts, U32).ravel()
  n = counts.size
  odd = n % 2
  if odd:
    counts = np.concatenate([counts, np.zeros(1, U32)])
  half = counts.size // 2
  y0, y1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
  out = np.concatenate([y0, y1])
  return out[:n] if odd else out


def PRNGKey(seed: int):
  seed = int(seed)
  return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], U32)


def split(key, num: int = 2):
  """jax.random.split (original layout): threefry_2x32(key, iota(2*num)).reshape(num, 2)."""
  return threefry_2x32(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def random_bits(key, shape):
  """jax.random.bits(key, shape, uint32) (original layout)."""
  size = int(np.prod(shape)) if np.ndim(shape) else int(shape)
  return threefry_2x32(key, np.arange(size, dtype=U32)).reshape(shape)


def bits_to_unit_float(bits):
  """bitcast((bits >> 9) | 0x3F800000) - 1.0  in [0, 1)."""
  f = ((np.asarray(bits, U32) >> U32(9)) | U32(0x3F800000)).view(F32)
  return f - F32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
  """jax.random.uniform(key, shape, float32, minval, maxval)."""
  minval = F32(minval)
  maxval = F32(maxval)
  floats = bits_to_unit_float(random_bits(key, shape))
  return np.maximum(minval, floats * (maxval - minval) + minval).astype(F32)


_ERFINV_LT5 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
               0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
_ERFINV_GE5 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
               0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]


def _f32_of_f64(fn, x):
  """Transcendental evaluated in binary64 and rounded once to binary32 (oracle convention)."""
  return fn(np.asarray(x, F32).astype(np.float64)).astype(F32)


def exp32(x):
  return _f32_of_f64(np.exp, x)


def log32(x):
  with np.errstate(divide='ignore'):
    return _f32_of_f64(np.log, x)


def log1p32(x):
  with np.errstate(divide='ignore'):
    return _f32_of_f64(np.log1p, x)


def erf_inv32(x):
  """XLA ErfInv32: w = -log1p(-x*x); two degree-8 polynomials in w (Giles)."""
  x = np.asarray(x, F32)
  w = -log1p32(-(x * x))
  lt = w < F32(5.0)
  with np.errstate(invalid='ignore'):
    w2 = np.where(lt, w - F32(2.5), np.sqrt(w, dtype=F32) - F32(3.0)).astype(F32)
  p = np.where(lt, F32(_ERFINV_LT5[0]), F32(_ERFINV_GE5[0])).astype(F32)
  for i in range(1, 9):
    c = np.where(lt, F32(_ERFINV_LT5[i]), F32(_ERFINV_GE5[i])).astype(F32)
    p = (c + (p * w2).astype(F32)).astype(F32)
  res = (p * x).astype(F32)
  return np.where(np.abs(x) == F32(1.0), x * F32(np.inf), res).astype(F32)


_NORMAL_LO = np.nextafter(F32(-1.0), F32(0.0), dtype=F32)
_SQRT2 = F32(np.sqrt(2))


def normal_from_bits(bits):
  """jax.random.normal's transform of raw 32-bit words (uniform on (-1,1) -> sqrt2*erf_inv)."""
  floats = bits_to_unit_float(bits)
  span = F32(F32(1.0) - _NORMAL_LO)     # rounds to 2.0f, as in jax
  u = np.maximum(_NORMAL_LO, (floats * span).astype(F32) + _NORMAL_LO).astype(F32)
  return (_SQRT2 * erf_inv32(u)).astype(F32)


def normal(key, shape=()):
  """jax.random.normal(key, shape, float32)."""
  return normal_from_bits(random_bits(key, shape))


def sigmoid32(x):
  """jax.nn.sigmoid in f32: 1 / (1 + exp(-x))."""
  x = np.asarray(x, F32)
  return (F32(1.0) / (F32(1.0) + exp32(-x))).astype(F32)


def log_sigmoid32(x):
  """jax.nn.log_sigmoid(x) = -softplus(-x) = -logaddexp(-x, 0)."""
  x = np.asarray(x, F32)
  # logaddexp(a, 0) = max(a,0) + log1p(exp(-|a|))
  a = -x
  return (-(np.maximum(a, F32(0.0)) + log1p32(exp32(-np.abs(a))))).astype(F32)


def bernoulli_logits_sample(key, logits):
  """distrax.Bernoulli(logits).sample(seed=key): uniform(key,(1,)+shape) < sigmoid(logits)."""
  logits = np.asarray(logits, F32)
  u = uniform(key, logits.shape)
  return (u < sigmoid32(logits)).astype(np.int32)


def bernoulli_logits_log_prob(logits, value):
  """distrax.Bernoulli(logits).log_prob(value) = (1-v)*log_sigmoid(-l) + v*log_sigmoid(l)."""
  logits = np.asarray(logits, F32)
  v = np.asarray(value).astype(F32)
  lp0 = log_sigmoid32(-logits)
  lp1 = log_sigmoid32(logits)
  # math.multiply_no_nan(log_probs, 1 - v) + math.multiply_no_nan(log_probs1, v)
  t0 = np.where(v == 1, F32(0), lp0 * (F32(1) - v))
  t1 = np.where(v == 0, F32(0), lp1 * v)
  return (t0 + t1).astype(F32)


def bernoulli_probs_sample(key, p, shape):
  """distrax.Bernoulli(probs=p).sample(sample_shape=shape, seed=key)."""
  u = uniform(key, shape)
  return (u < F32(p)).astype(np.int32)
