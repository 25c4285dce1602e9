"""Parameterised real-valued protocol functions with analytic Jacobians (host side).

Drop-in for the reference's `noneq_opt/parameterization.py`: the same classes, constructor
arguments, `variables` / `constants` / `domain` / `tree_flatten` / `tree_unflatten` surface
(parameterization.py:34-64), for
  Constant :67-79, PiecewiseLinear :82-133, Chebyshev :144-179 (monic, power basis :137-141),
  ChangeDomain :268-292, ConstrainEndpoints :295-322, AddBaseline :325-344.
`Spline` (:182-265) is out of scope for this build (it needs jax_cosmo's spline).

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
    if np.ndim(self.weights) != 1:
      raise NotImplementedError('jacobian of multi-row Chebyshev weights is not supported')
    return self.basis(x)


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


class Spline(Parameterization):
  """parameterization.py:182-265 -- not available in this build (needs jax_cosmo's spline)."""
  knots: np.ndarray
  values: np.ndarray
  _VARIABLES = ('values',)

  def __post_init__(self):
    raise NotImplementedError('Spline is out of scope for the B200 hot-path build (SURVEY N4)')
