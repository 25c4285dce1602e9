"""Parameterised real-valued protocol functions with analytic Jacobians (host side).

Drop-in for the reference's `noneq_opt/parameterization.py`: the same classes, constructor
arguments, `variables` / `constants` / `domain` / `tree_flatten` / `tree_unflatten` surface
(parameterization.py:34-64), for
  Constant :67-79, PiecewiseLinear :82-133, Chebyshev :144-179 (monic, power basis :137-141),
  ChangeDomain :268-292, ConstrainEndpoints :295-322, AddBaseline :325-344.
`Spline` (:182-265) is built from the definition of the interpolating spline (jax_cosmo is not
available offline; see the class docstring for what that pins and what it does not).

What is different, by design: there is no tracing autodiff here, so every class also provides
`jacobian(x) -> [len(x), n_variables]`, the derivative of the value wrt its flattened variables
(in `tree_leaves` order).  That table is what the CUDA kernels contract the per-step
sensitivities with (`phi` in include/noneq_b200.h), replacing `jax.grad` through the protocol.
Values are evaluated in float32 in the reference's operation order, on the host: a schedule is
a few thousand floats, built once per gradient call.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, ClassVar, Tuple

import numpy as np
import scipy.special as sps

F32 = np.float32


def _f32(x):
  return np.asarray(x, dtype=F32)


class Parameterization:
  """Base class: subclasses are dataclasses; `_VARIABLES` names the trainable fields."""
  _VARIABLES: ClassVar[Tuple[str, ...]] = ()

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    dataclasses.dataclass(eq=False)(cls)

  # -- pytree surface (parameterization.py:39-61)
  @property
  def variables(self):
    return {n: getattr(self, n) for n in self._VARIABLES}

  @property
  def constants(self):
    return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)
            if f.name not in self._VARIABLES}

  def tree_flatten(self):
    return ((self.variables,), self.constants)

  @classmethod
  def tree_unflatten(cls, constants, variables):
    variables, = variables
    return cls(**variables, **constants)

  def __eq__(self, other):
    if type(self) is not type(other):
      return False
    for f in dataclasses.fields(self):
      a, b = getattr(self, f.name), getattr(other, f.name)
      if isinstance(a, Parameterization) or callable(a) and not isinstance(a, np.ndarray):
        if not (a == b):
          return False
      elif not np.array_equal(np.asarray(a), np.asarray(b)):
        return False
    return True

  __hash__ = None

  # -- numerics
  @property
  def domain(self):
    raise NotImplementedError

  def __call__(self, x):
    raise NotImplementedError

  def jacobian(self, x):
    """d value(x_i) / d variables_j as float32 [len(x), n_variables]."""
    raise NotImplementedError

  @property
  def n_variables(self):
    from . import tree
    return int(sum(np.asarray(l).size for l in tree.tree_leaves(self.variables)))

  @property
  def is_affine(self):
    """value(x) = jacobian(x) @ variables + const?  True for every class of the reference except a Spline with
    trainable knots; wrappers inherit it from what they wrap."""
    w = getattr(self, 'wrapped', None)
    return w.is_affine if isinstance(w, Parameterization) else True


class Constant(Parameterization):
  value: np.ndarray
  _VARIABLES = ('value',)

  @property
  def domain(self):
    return (0., 1.)

  def __call__(self, x):
    return _f32(self.value) * np.ones_like(_f32(x))

  def jacobian(self, x):
    return np.ones((np.atleast_1d(_f32(x)).shape[0], 1), F32)


class PiecewiseLinear(Parameterization):
  values: np.ndarray
  x0: np.ndarray = 0.
  x1: np.ndarray = 1.
  y0: np.ndarray = 0.
  y1: np.ndarray = 1.
  _VARIABLES = ('values',)

  @property
  def domain(self):
    return (self.x0, self.x1)

  @property
  def degree(self):
    return np.shape(self.values)[-1]

  @property
  def length(self):
    return self.degree + 2

  @property
  def x(self):
    return np.linspace(F32(self.x0), F32(self.x1), self.length, dtype=F32)

  @property
  def y(self):
    return np.concatenate([[F32(self.y0)], _f32(self.values), [F32(self.y1)]]).astype(F32)

  @classmethod
  def equidistant(cls, d, x0=0., x1=1., y0=0., y1=1.):
    values = np.linspace(F32(y0), F32(y1), d + 2, dtype=F32)[1:-1]
    return cls(values, x0, x1, y0, y1)

  def __call__(self, x):
    y = self.y
    return np.interp(_f32(x), self.x, y, left=y[0], right=y[-1]).astype(F32)

  def jacobian(self, x):
    # interp is linear in the knot values: column j is the hat function of knot j+1
    xs = np.atleast_1d(_f32(x))
    n = self.degree
    jac = np.zeros((xs.shape[0], n), F32)
    knots = self.x
    for j in range(n):
      hat = np.zeros(n + 2, F32)
      hat[j + 1] = 1
      jac[:, j] = np.interp(xs, knots, hat)
    return jac


_BASIS_CACHE = {}


def chebyshev_coefficients(degree):
  """Rows j = 0..degree of monic Chebyshev-T_j in the power basis, highest power first."""
  rows = []
  for j in range(degree + 1):
    poly = np.asarray(sps.chebyt(j, monic=True).coeffs, dtype=np.float64)
    rows.append(np.concatenate([np.zeros(degree - j), poly]))
  return np.stack(rows)


class Chebyshev(Parameterization):
  # Weights have shape `[degree + 1]` (or `[ndim, degree + 1]`).
  weights: np.ndarray
  _VARIABLES = ('weights',)

  @property
  def domain(self):
    return (0., 1.)

  @property
  def degree(self):
    return np.shape(self.weights)[-1] - 1

  @property
  def coefficients(self):
    return chebyshev_coefficients(self.degree)

  def _powers(self, x):
    # row p holds x^(degree-p): repeated float32 multiplication, last row ones
    deg = self.degree
    rows = [None] * (deg + 1)
    acc = np.ones_like(x)
    rows[deg] = acc
    for p in range(deg - 1, -1, -1):
      acc = (acc * x).astype(F32)
      rows[p] = acc
    return np.stack(rows, 0)

  def basis(self, x):
    """[len(x), degree+1] monic Chebyshev basis at 2x-1 (float32, power-basis evaluation).  The
    basis depends on (degree, x) only, not on the weights: it is cached, so that an optimisation
    loop which re-evaluates the schedule and its Jacobian on the same time grid pays once."""
    xs = np.atleast_1d(_f32(x))
    key = (self.degree, xs.shape, xs.tobytes())
    hit = _BASIS_CACHE.get(key)
    if hit is not None:
      return hit
    t = (F32(2) * xs - F32(1)).astype(F32)
    pw = self._powers(t)
    coef = self.coefficients.astype(F32)
    out = np.zeros((xs.shape[0], self.degree + 1), F32)
    for p in range(self.degree + 1):
      out = (out + (coef[:, p][None, :] * pw[p][:, None]).astype(F32)).astype(F32)
    out.setflags(write=False)
    if len(_BASIS_CACHE) >= 64:
      _BASIS_CACHE.pop(next(iter(_BASIS_CACHE)))
    _BASIS_CACHE[key] = out
    return out

  def __call__(self, x):
    w = _f32(self.weights)
    phi = self.basis(x)
    if w.ndim == 1:
      y = np.zeros(phi.shape[0], F32)
      for j in range(w.shape[0]):
        y = (y + (w[j] * phi[:, j]).astype(F32)).astype(F32)
      return y if np.ndim(x) else y[0]
    return np.stack([Chebyshev(row)(x) for row in w], 0)

  def jacobian(self, x):
    """[len(x), D] for a weight vector; for weights `[ndim, D]` (one polynomial per output row,
    parameterization.py:145) `[ndim, len(x), ndim * D]`: row i depends only on its own D weights."""
    phi = self.basis(x)
    if np.ndim(self.weights) == 1:
      return phi
    ndim, D = np.shape(self.weights)
    jac = np.zeros((ndim, phi.shape[0], ndim * D), F32)
    for i in range(ndim):
      jac[i, :, i * D:(i + 1) * D] = phi
    return jac


class ChangeDomain(Parameterization):
  """Wraps another `Parameterization`, linearly moving it to a new domain."""
  wrapped: Parameterization
  x0: np.ndarray
  x1: np.ndarray
  _VARIABLES = ('wrapped',)

  @property
  def domain(self):
    return (self.x0, self.x1)

  def _to_wrapped(self, x):
    w0, w1 = self.wrapped.domain
    scale = F32((w1 - w0) / (self.x1 - self.x0))
    return ((_f32(x) - F32(self.x0)) * scale + F32(w0)).astype(F32)

  def __call__(self, x):
    return self.wrapped(self._to_wrapped(x))

  def jacobian(self, x):
    return self.wrapped.jacobian(self._to_wrapped(x))


class ConstrainEndpoints(Parameterization):
  """(x - x0)(x1 - x) wrapped(x) + the straight line from (x0, y0) to (x1, y1)."""
  wrapped: Parameterization
  y0: np.ndarray
  y1: np.ndarray
  _VARIABLES = ('wrapped',)

  @property
  def domain(self):
    return self.wrapped.domain

  def _envelope(self, x):
    x0, x1 = (F32(v) for v in self.domain)
    return ((x - x0) * (x1 - x)).astype(F32)

  def __call__(self, x):
    xs = _f32(x)
    x0, x1 = (F32(v) for v in self.domain)
    linear = (F32(self.y0) + (xs - x0) / (x1 - x0) * F32(self.y1 - self.y0)).astype(F32)
    return (self._envelope(xs) * self.wrapped(xs) + linear).astype(F32)

  def jacobian(self, x):
    xs = np.atleast_1d(_f32(x))
    return (self._envelope(xs)[:, None] * self.wrapped.jacobian(xs)).astype(F32)


class AddBaseline(Parameterization):
  """wrapped(x) + baseline(x); the baseline is never fitted."""
  wrapped: Parameterization
  baseline: Callable
  _VARIABLES = ('wrapped',)

  @property
  def domain(self):
    return self.wrapped.domain

  def __call__(self, x):
    xs = _f32(x)
    return (self.wrapped(xs) + _f32(self.baseline(xs))).astype(F32)

  def jacobian(self, x):
    return self.wrapped.jacobian(x)


def _spline_matrix(xk, k, endpoints, xs):
  """M [len(xs), len(xk)] with spline(xs) = M @ y for the interpolating spline of degree k through
  (xk, y): piecewise polynomials on the data intervals, C^{k-1} at the interior points, closed by
  the `endpoints` condition -- 'not-a-knot' (the k-th derivative is continuous at the first and,
  for k = 3, also the last interior point) or 'natural' (second derivative zero at the ends).
  Evaluation outside [xk[0], xk[-1]] continues the end pieces.  float64 dense solve: a protocol has
  a few dozen knots and the matrix is cached by the caller."""
  xk = np.asarray(xk, np.float64)
  xs = np.atleast_1d(np.asarray(xs, np.float64))
  n = xk.shape[0]
  if k not in (1, 2, 3):
    raise ValueError(f'spline degree must be 1, 2 or 3 (got {k})')
  if endpoints not in ('not-a-knot', 'natural'):
    raise ValueError(f"endpoints must be 'not-a-knot' or 'natural' (got {endpoints!r})")
  if n < k + 1:
    raise ValueError(f'a degree-{k} spline needs at least {k + 1} points (got {n})')
  if np.any(np.diff(xk) <= 0):
    raise ValueError('spline knots must be strictly increasing')
  nint, nc = n - 1, k + 1
  h = np.diff(xk)
  rows, rhs = [], []      # rows act on the coefficient vector c[i * nc + d]; rhs rows act on y

  def deriv_row(i, t, order):
    """coefficients of d^order/dt^order p_i at offset t from x_i"""
    r = np.zeros(nint * nc)
    for d in range(order, nc):
      f = 1.0
      for q in range(order):
        f *= d - q
      r[i * nc + d] = f * t ** (d - order)
    return r

  def unit(j):
    e = np.zeros(n)
    if j is not None:
      e[j] = 1.0
    return e
  for i in range(nint):
    rows.append(deriv_row(i, 0.0, 0)); rhs.append(unit(i))
    rows.append(deriv_row(i, h[i], 0)); rhs.append(unit(i + 1))
  for i in range(nint - 1):
    for order in range(1, k):
      rows.append(deriv_row(i, h[i], order) - deriv_row(i + 1, 0.0, order)); rhs.append(unit(None))
  if k >= 2:
    if endpoints == 'not-a-knot' and nint >= 2:
      rows.append(deriv_row(0, h[0], k) - deriv_row(1, 0.0, k)); rhs.append(unit(None))
      if k == 3:
        if nint >= 3:
          rows.append(deriv_row(nint - 2, h[nint - 2], k) - deriv_row(nint - 1, 0.0, k))
        else:   # three points: the single cubic through them with zero third derivative
          rows.append(deriv_row(0, 0.0, 3))
        rhs.append(unit(None))
    else:
      rows.append(deriv_row(0, 0.0, 2)); rhs.append(unit(None))
      if k == 3:
        rows.append(deriv_row(nint - 1, h[-1], 2)); rhs.append(unit(None))
  A, Bm = np.stack(rows), np.stack(rhs)
  coef = np.linalg.solve(A, Bm)                     # [nint*nc, n]: coefficients as linear maps of y
  idx = np.clip(np.searchsorted(xk, xs, side='right') - 1, 0, nint - 1)
  t = xs - xk[idx]
  M = np.zeros((xs.shape[0], n))
  for d in range(nc):
    M += (t ** d)[:, None] * coef[idx * nc + d]
  return M


_SPLINE_CACHE = {}


class Spline(Parameterization):
  """Interpolating spline through (x0, y0), (knots, values), (x1, y1) (parameterization.py:182-265).

  The reference delegates to `jax_cosmo.scipy.interpolate.InterpolatedUnivariateSpline(x, y,
  k=spline_degree, endpoints=spline_endpoints)`, which is not available offline; this class builds
  the spline from its definition (`_spline_matrix`).  Cubic not-a-knot / natural splines are unique
  and agree with scipy's CubicSpline (tests/test_host_logic.py); for degree 2 the closing condition
  is an assumption about jax_cosmo's convention -- parity unpinned.

  `variable_knots=True` (:192, 195-198) makes the interior knots trainable too: `variables` is then
  {'knots', 'values'} (pytree leaf order: knots, values -- dict keys sort) and `jacobian` carries the columns
  d spline(x) / d knot_k in front of the value columns.  The value is NOT affine in the knots, which is fine for
  the gradient estimators (they only need the Jacobian at the current variables: chain rule) but not for the
  device-resident trainers, whose on-device schedule rebuild assumes an affine protocol (`is_affine`)."""
  knots: np.ndarray
  values: np.ndarray
  x0: np.ndarray = 0.
  x1: np.ndarray = 1.
  y0: np.ndarray = 0.
  y1: np.ndarray = 1.
  spline_degree: int = 2
  spline_endpoints: str = 'not-a-knot'
  variable_knots: bool = False
  _VARIABLES = ('values',)

  @property
  def variables(self):
    v = dict(values=self.values)
    if self.variable_knots:
      v['knots'] = self.knots
    return v

  @property
  def constants(self):
    names = set(self.variables)
    return {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.name not in names}

  @property
  def is_affine(self):
    return not self.variable_knots

  @property
  def domain(self):
    return (self.x0, self.x1)

  @property
  def degree(self):
    return np.shape(self.knots)[-1]

  @property
  def length(self):
    return self.degree + 2

  @property
  def x(self):
    return np.concatenate([[F32(self.x0)], _f32(self.knots), [F32(self.x1)]]).astype(F32)

  @property
  def y(self):
    return np.concatenate([[F32(self.y0)], _f32(self.values), [F32(self.y1)]]).astype(F32)

  @classmethod
  def equidistant(cls, d, x0=0., x1=1., y0=0., y1=1., **kwargs):
    knots = np.linspace(F32(x0), F32(x1), d + 2, dtype=F32)[1:-1]
    values = np.linspace(F32(y0), F32(y1), d + 2, dtype=F32)[1:-1]
    return cls(knots, values, x0, x1, y0, y1, **kwargs)

  @classmethod
  def from_baseline(cls, baseline, d, x0=0., x1=1., **kwargs):
    y0, y1 = baseline(_f32(x0)), baseline(_f32(x1))
    knots = np.linspace(F32(x0), F32(x1), d + 2, dtype=F32)[1:-1]
    return cls(knots, _f32(baseline(knots)), x0, x1, y0, y1, **kwargs)

  def _matrix(self, x):
    xs = np.atleast_1d(_f32(x))
    xk = self.x
    key = (xk.tobytes(), int(self.spline_degree), self.spline_endpoints, xs.shape, xs.tobytes())
    hit = _SPLINE_CACHE.get(key)
    if hit is None:
      hit = _spline_matrix(xk, int(self.spline_degree), self.spline_endpoints, xs)
      if len(_SPLINE_CACHE) >= 32:
        _SPLINE_CACHE.pop(next(iter(_SPLINE_CACHE)))
      _SPLINE_CACHE[key] = hit
    return hit

  def __call__(self, x):
    y = (self._matrix(x) @ self.y.astype(np.float64)).astype(F32)
    return y if np.ndim(x) else y[0]

  def jacobian(self, x):
    Jv = self._matrix(x)[:, 1:-1]
    if not self.variable_knots:
      return Jv.astype(F32)
    # d spline(x) / d knot_k: central differences of the float64 construction (smooth in the knots; the step is
    # 1e-6 of the neighbouring knot spacing, truncation error ~1e-12 relative), columns in front of the values'
    xs = np.atleast_1d(_f32(x)).astype(np.float64)
    xk, y = self.x.astype(np.float64), self.y.astype(np.float64)
    k, ep = int(self.spline_degree), self.spline_endpoints
    Jk = np.zeros((xs.shape[0], xk.shape[0] - 2))
    for j in range(1, xk.shape[0] - 1):
      h = 1e-6 * min(xk[j] - xk[j - 1], xk[j + 1] - xk[j])
      up, dn = xk.copy(), xk.copy()
      up[j] += h
      dn[j] -= h
      Jk[:, j - 1] = (_spline_matrix(up, k, ep, xs) @ y - _spline_matrix(dn, k, ep, xs) @ y) / (2 * h)
    return np.concatenate([Jk, Jv], 1).astype(F32)
