"""Oracle (test infrastructure): restatement of noneq_opt/parameterization.py in numpy.

Follows /root/reference/noneq_opt/parameterization.py:
  chebyshev_coefficients  :137-141     Chebyshev            :144-179
  PiecewiseLinear         :82-133      Constant             :67-79
  ChangeDomain            :268-292     ConstrainEndpoints   :295-322
  AddBaseline             :325-344
and the baselines of noneq_opt/ising.py:45-55.

Each class is a plain callable `f(x) -> y` evaluated in float32 in the operation
order of the reference, plus `jacobian(x) -> [len(x), n_variables]`, the derivative
of the value wrt the flattened variables (what `jax.grad` / `value_and_jacfwd`
produce for these pytrees), obtained here by differentiating the same formulas by
hand.  `dtype=np.float64` gives the "truth" variant used for error accounting.
Spline (:182-265) has no restatement here: jax_cosmo is absent offline and not named by the north
star; the product's Spline is checked against scipy's CubicSpline in tests/test_host_logic.py.
"""
from __future__ import annotations

import numpy as np
import scipy.special as sps


def chebyshev_coefficients(degree):
  """parameterization.py:137-141 -- monic Chebyshev-T power-basis rows, highest power first."""
  return np.stack([
      np.concatenate([np.zeros(degree - j), np.asarray(sps.chebyt(j, True))])
      for j in range(degree + 1)])


class Chebyshev:
  """parameterization.py:144-179.  y(x) = sum_w weights[w] sum_p C[w,p] (2x-1)^(deg-p)."""

  def __init__(self, weights, dtype=np.float32):
    self.dtype = dtype
    self.weights = np.asarray(weights, dtype)

  @property
  def degree(self):
    return self.weights.shape[-1] - 1

  @property
  def n_variables(self):
    return self.weights.size

  def _powers(self, x):
    # :164-173 -- reverse scan of y *= x; row p holds x^(deg-p), last row ones.
    deg = self.degree
    rows = [None] * (deg + 1)
    rows[deg] = np.ones_like(x)
    y = np.ones_like(x)
    for p in range(deg - 1, -1, -1):
      y = (y * x).astype(self.dtype)
      rows[p] = y
    return np.stack(rows, 0)

  def basis(self, x):
    """[len(x), deg+1]: T~_j(2x-1) evaluated as in the reference's einsum (power basis)."""
    x = np.asarray(x, self.dtype)
    xs = (self.dtype(2) * x - self.dtype(1)).astype(self.dtype)
    pw = self._powers(xs)                       # [P, n]
    C = chebyshev_coefficients(self.degree).astype(self.dtype)   # [W, P]
    out = np.zeros((x.shape[0], self.degree + 1), self.dtype)
    for p in range(self.degree + 1):            # sequential sum over p
      out = (out + (C[:, p][None, :] * pw[p][:, None]).astype(self.dtype)).astype(self.dtype)
    return out

  def __call__(self, x):
    phi = self.basis(np.atleast_1d(x))
    y = np.zeros(phi.shape[0], self.dtype)
    for w in range(self.degree + 1):
      y = (y + (self.weights[w] * phi[:, w]).astype(self.dtype)).astype(self.dtype)
    return y

  def jacobian(self, x):
    return self.basis(np.atleast_1d(x))


class Constant:
  """parameterization.py:67-79."""

  def __init__(self, value, dtype=np.float32):
    self.dtype = dtype
    self.value = np.asarray(value, dtype)
  n_variables = 1

  def __call__(self, x):
    return (self.value * np.ones_like(np.atleast_1d(np.asarray(x, self.dtype)))).astype(self.dtype)

  def jacobian(self, x):
    return np.ones((np.atleast_1d(x).shape[0], 1), self.dtype)


class PiecewiseLinear:
  """parameterization.py:82-133 -- jnp.interp over linspace(x0,x1,d+2) knots, clamped ends."""

  def __init__(self, values, x0=0., x1=1., y0=0., y1=1., dtype=np.float32):
    self.dtype = dtype
    self.values = np.asarray(values, dtype)
    self.x0, self.x1, self.y0, self.y1 = (dtype(v) for v in (x0, x1, y0, y1))

  @property
  def n_variables(self):
    return self.values.size

  @property
  def x(self):
    return np.linspace(self.x0, self.x1, self.values.shape[-1] + 2, dtype=self.dtype)

  @property
  def y(self):
    return np.concatenate([[self.y0], self.values, [self.y1]]).astype(self.dtype)

  def __call__(self, x):
    x = np.atleast_1d(np.asarray(x, self.dtype))
    return np.interp(x, self.x, self.y, left=self.y[0], right=self.y[-1]).astype(self.dtype)

  def jacobian(self, x):
    x = np.atleast_1d(np.asarray(x, self.dtype))
    n = self.values.shape[-1]
    J = np.zeros((x.shape[0], n), self.dtype)
    for j in range(n):
      e = np.zeros(n + 2, self.dtype)
      e[j + 1] = 1
      J[:, j] = np.interp(x, self.x, e, left=e[0], right=e[-1])
    return J


class ChangeDomain:
  """parameterization.py:268-292."""

  def __init__(self, wrapped, x0, x1, wrapped_domain=(0., 1.)):
    self.wrapped, self.x0, self.x1, self.wd = wrapped, x0, x1, wrapped_domain
    self.dtype = wrapped.dtype

  @property
  def n_variables(self):
    return self.wrapped.n_variables

  def _map(self, x):
    d = self.dtype
    x = np.atleast_1d(np.asarray(x, d))
    scale = d((self.wd[1] - self.wd[0]) / (self.x1 - self.x0))
    return ((x - d(self.x0)) * scale + d(self.wd[0])).astype(d)

  def __call__(self, x):
    return self.wrapped(self._map(x))

  def jacobian(self, x):
    return self.wrapped.jacobian(self._map(x))


class ConstrainEndpoints:
  """parameterization.py:295-322 -- (x-x0)(x1-x) wrapped(x) + linear component."""

  def __init__(self, wrapped, y0, y1, domain=(0., 1.)):
    self.wrapped, self.y0, self.y1, self.domain = wrapped, y0, y1, domain
    self.dtype = wrapped.dtype

  @property
  def n_variables(self):
    return self.wrapped.n_variables

  def _envelope(self, x):
    d = self.dtype
    x0, x1 = d(self.domain[0]), d(self.domain[1])
    return ((x - x0) * (x1 - x)).astype(d)

  def __call__(self, x):
    d = self.dtype
    x = np.atleast_1d(np.asarray(x, d))
    x0, x1 = d(self.domain[0]), d(self.domain[1])
    linear = (d(self.y0) + (x - x0) / (x1 - x0) * d(self.y1 - self.y0)).astype(d)
    return (self._envelope(x) * self.wrapped(x) + linear).astype(d)

  def jacobian(self, x):
    d = self.dtype
    x = np.atleast_1d(np.asarray(x, d))
    return (self._envelope(x)[:, None] * self.wrapped.jacobian(x)).astype(d)


class AddBaseline:
  """parameterization.py:325-344."""

  def __init__(self, wrapped, baseline):
    self.wrapped, self.baseline = wrapped, baseline
    self.dtype = wrapped.dtype

  @property
  def n_variables(self):
    return self.wrapped.n_variables

  def __call__(self, x):
    d = self.dtype
    x = np.atleast_1d(np.asarray(x, d))
    return (self.wrapped(x) + np.asarray(self.baseline(x), d)).astype(d)

  def jacobian(self, x):
    return self.wrapped.jacobian(x)


def log_temp_baseline(min_temp=.69, max_temp=10., degree=1, dtype=np.float32):
  """ising.py:45-50 (log evaluated in binary64 and rounded, oracle convention)."""
  def _f(t):
    t = np.asarray(t, dtype)
    scale = dtype(max_temp - min_temp)
    shape = ((dtype(1) - t) ** degree * t ** degree * dtype(4 ** degree)).astype(dtype)
    return np.log((shape * scale + dtype(min_temp)).astype(dtype).astype(np.float64)).astype(dtype)
  return _f


def field_baseline(start_field=-1., end_field=1., dtype=np.float32):
  """ising.py:52-55."""
  def _f(t):
    t = np.asarray(t, dtype)
    return ((dtype(1) - t) * dtype(start_field) + t * dtype(end_field)).astype(dtype)
  return _f
