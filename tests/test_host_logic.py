"""Host-side logic of the product package (no GPU): parameterisations (re-expressing the
reference's tests/parameterization_test.py), class tables, optimisers, tree / control-flow /
gradients utilities, and their agreement with the oracle."""
import numpy as np
import pytest
import torch

from noneq_opt_b200 import control_flow, gradients, ising, optimizers, parameterization as p10n, tree
from noneq_opt_b200 import barrier_crossing as bc
from oracle import ising_ref as I, jaxlike as J, langevin_ref as L, parameterization_ref as P

p10ns = [
    p10n.Constant(np.pi),
    p10n.PiecewiseLinear(np.linspace(-4, 6, 200)),
    p10n.Chebyshev(np.zeros(32)),
    p10n.ChangeDomain(p10n.PiecewiseLinear.equidistant(10), -11., 22.),
    p10n.ConstrainEndpoints(p10n.PiecewiseLinear.equidistant(10), -11., 22.),
    p10n.AddBaseline(p10n.Chebyshev(np.arange(5.)), np.sin),
]


class TestParameterizationProperties:     # tests/parameterization_test.py:23-49

  @pytest.mark.parametrize('p', p10ns)
  def test_finite(self, p):
    assert np.isfinite(p(np.linspace(0, 1, 100))).all()

  @pytest.mark.parametrize('p', p10ns)
  def test_flatten_unflatten(self, p):
    leaves, treedef = tree.tree_flatten(p)
    assert p == tree.tree_unflatten(treedef, leaves)

  @pytest.mark.parametrize('p', p10ns)
  def test_has_domain(self, p):
    assert isinstance(p.domain, tuple)

  @pytest.mark.parametrize('p', p10ns)
  def test_variables_and_constants(self, p):
    all_fields = set(p.__dataclass_fields__)
    variables, constants = set(p.variables), set(p.constants)
    assert variables.isdisjoint(constants)
    assert variables.union(constants) == all_fields


@pytest.mark.parametrize('x0, x1', [(17., 333.), (0., 100.)])
@pytest.mark.parametrize('p', p10ns)
def test_change_domain(p, x0, x1):            # tests/parameterization_test.py:52-61
  moved = p10n.ChangeDomain(p, x0=x0, x1=x1)
  np.testing.assert_allclose((x0, x1), moved.domain)
  np.testing.assert_allclose(moved(x0), p(p.domain[0]), rtol=1e-6)
  np.testing.assert_allclose(moved(x1), p(p.domain[1]), rtol=1e-6)


@pytest.mark.parametrize('y0, y1', [(17., 333.), (0., 100.)])
@pytest.mark.parametrize('p', p10ns)
def test_constrain_endpoints(p, y0, y1):      # tests/parameterization_test.py:64-74
  c = p10n.ConstrainEndpoints(p, y0=y0, y1=y1)
  x0, x1 = p.domain
  np.testing.assert_allclose(p.domain, c.domain)
  np.testing.assert_allclose(y0, c(x0))
  np.testing.assert_allclose(y1, c(x1))


@pytest.mark.parametrize('baseline', [lambda x: x ** 2, np.sqrt])
@pytest.mark.parametrize('p', p10ns)
def test_add_baseline(p, baseline):           # tests/parameterization_test.py:76-84
  xs = np.linspace(*p.domain, num=100).astype(np.float32)
  np.testing.assert_allclose(p(xs) + baseline(xs), p10n.AddBaseline(wrapped=p, baseline=baseline)(xs), rtol=1e-6)


def test_spline_parameterization():
  """parameterization.py:182-265.  The reference delegates to jax_cosmo (absent offline); the cubic
  interpolating spline is unique, so scipy's CubicSpline is the checker for degree 3; degrees 1 and
  2 are checked through their defining properties."""
  from scipy.interpolate import CubicSpline
  rng = np.random.RandomState(0)
  kn = np.sort(rng.uniform(0.05, 0.95, 6)).astype(np.float32)
  v = rng.normal(size=6).astype(np.float32)
  xs = np.linspace(0, 1, 101, dtype=np.float32)
  for ep in ('not-a-knot', 'natural'):
    sp = p10n.Spline(kn, v, 0., 1., 0.5, -0.5, spline_degree=3, spline_endpoints=ep)
    ref = CubicSpline(sp.x.astype(np.float64), sp.y.astype(np.float64), bc_type=ep)(xs.astype(np.float64))
    np.testing.assert_allclose(sp(xs), ref, atol=2e-6)
  sp2 = p10n.Spline(kn, v, 0., 1., 0.5, -0.5)           # defaults: degree 2, not-a-knot
  np.testing.assert_allclose(sp2(sp2.x), sp2.y, atol=1e-6)   # interpolates its points
  xk64, h = sp2.x.astype(np.float64), 1e-7                 # C^1 at an interior knot (float64 operator)
  d = np.diff(p10n._spline_matrix(xk64, 2, 'not-a-knot', [xk64[3] - h, xk64[3], xk64[3] + h]) @ sp2.y.astype(np.float64))
  assert abs(d[0] - d[1]) < 1e-3 * abs(d[0])
  quad = lambda t: 1 + 2 * t - 3 * t * t
  spq = p10n.Spline(kn, quad(kn), 0., 1., quad(0.), quad(1.))
  np.testing.assert_allclose(spq(xs), quad(xs), atol=1e-4)   # a quadratic is its own spline
  sp1 = p10n.Spline(kn, v, 0., 1., 0.5, -0.5, spline_degree=1)
  np.testing.assert_allclose(sp1(xs), np.interp(xs, sp1.x, sp1.y), atol=1e-6)
  # pytree surface and the Jacobian (the value is affine in `values`)
  assert set(sp2.variables) == {'values'} and 'knots' in sp2.constants and sp2.degree == 6 and sp2.length == 8
  rebuilt = p10n.Spline.tree_unflatten(*reversed(sp2.tree_flatten()))
  assert rebuilt == sp2
  J = sp2.jacobian(xs)
  bumped = p10n.Spline(kn, v + np.float32(0.5) * np.eye(6, dtype=np.float32)[3], 0., 1., 0.5, -0.5)
  np.testing.assert_allclose((bumped(xs) - sp2(xs)) / 0.5, J[:, 3], atol=1e-5)
  # variable_knots=True (parameterization.py:192-198): the knots join the variables; leaf order knots, values
  from noneq_opt_b200 import tree
  for k in (2, 3):
    spv = p10n.Spline(kn, v, 0., 1., 0.5, -0.5, spline_degree=k, variable_knots=True)
    assert set(spv.variables) == {'values', 'knots'} and 'knots' not in spv.constants and not spv.is_affine
    names = {f.name for f in __import__('dataclasses').fields(spv)}
    assert set(spv.variables) | set(spv.constants) == names and not set(spv.variables) & set(spv.constants)
    leaves, treedef = tree.tree_flatten(spv)
    np.testing.assert_array_equal(leaves[0], kn)
    np.testing.assert_array_equal(leaves[1], v)
    assert tree.tree_unflatten(treedef, leaves) == spv and spv.n_variables == 12
    Jv = spv.jacobian(xs)
    assert Jv.shape == (101, 12)
    np.testing.assert_array_equal(Jv[:, 6:], p10n.Spline(kn, v, 0., 1., 0.5, -0.5, spline_degree=k).jacobian(xs))
    M = lambda knots: p10n._spline_matrix(np.concatenate([[0.], knots, [1.]]), k, 'not-a-knot', xs.astype(np.float64)) @ spv.y.astype(np.float64)
    for j in (0, 3, 5):      # against a wide central difference of the float64 construction
      e = np.zeros(6); e[j] = 1e-4
      fd = (M(kn.astype(np.float64) + e) - M(kn.astype(np.float64) - e)) / 2e-4
      np.testing.assert_allclose(Jv[:, j], fd, atol=2e-5 * max(1.0, np.abs(fd).max()))
  assert p10n.AddBaseline(p10n.ConstrainEndpoints(spv, 0., 0.), lambda t: t).is_affine is False
  assert p10n.ConstrainEndpoints(p10n.Chebyshev(np.zeros(3, np.float32)), 0., 0.).is_affine
  eq = p10n.Spline.equidistant(4, 0., 2., 1., 3., spline_degree=3)
  np.testing.assert_allclose(eq(np.linspace(0, 2, 9, dtype=np.float32)), np.linspace(1, 3, 9), atol=1e-6)
  fb = p10n.Spline.from_baseline(lambda t: np.float32(2) * t, 3, 0., 1.)
  np.testing.assert_allclose(fb(xs), 2 * xs, atol=1e-6)
  with pytest.raises(ValueError):
    p10n.Spline(kn[::-1].copy(), v)(xs)


def test_product_parameterizations_equal_oracle_bits():
  w = (0.3 * np.random.RandomState(1).normal(size=13)).astype(np.float32)
  xs = np.linspace(0, 1, 257, dtype=np.float32)
  np.testing.assert_array_equal(p10n.Chebyshev(w)(xs), P.Chebyshev(w)(xs))
  np.testing.assert_array_equal(p10n.Chebyshev(w).jacobian(xs), P.Chebyshev(w).jacobian(xs))
  np.testing.assert_array_equal(p10n.chebyshev_coefficients(12), P.chebyshev_coefficients(12))
  a = p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w), 0., 0.), ising.log_temp_baseline())
  b = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w), 0., 0.), P.log_temp_baseline())
  np.testing.assert_array_equal(a(xs), b(xs))
  np.testing.assert_array_equal(a.jacobian(xs), b.jacobian(xs))
  v = np.linspace(1, 3, 6).astype(np.float32)
  np.testing.assert_array_equal(p10n.PiecewiseLinear(v)(xs), P.PiecewiseLinear(v)(xs))
  np.testing.assert_array_equal(p10n.PiecewiseLinear(v).jacobian(xs), P.PiecewiseLinear(v).jacobian(xs))


@pytest.mark.parametrize('make', [
    lambda w: p10n.Chebyshev(w), lambda w: p10n.PiecewiseLinear(w),
    lambda w: p10n.ConstrainEndpoints(p10n.Chebyshev(w), 1., 2.),
    lambda w: p10n.ChangeDomain(p10n.PiecewiseLinear(w, 0., 1., 0., 1.), -1., 3.)])
def test_jacobian_is_the_derivative(make):
  w = np.linspace(0.2, 1.1, 6)
  xs = np.linspace(0.05, 0.95, 11)
  if isinstance(make(w), p10n.ChangeDomain):
    xs = np.linspace(-0.9, 2.9, 11)
  jac = make(w.astype(np.float32)).jacobian(xs.astype(np.float32))
  for j in range(6):
    e = np.zeros(6)
    e[j] = 1e-2
    fd = (make((w + e).astype(np.float32))(xs) - make((w - e).astype(np.float32))(xs)) / 2e-2
    np.testing.assert_allclose(jac[:, j], fd, atol=2e-4)


def test_trap_schedule_matches_oracle():
  coeffs = L.fit_linear_to_cheby(1.0, 1e-3, 5, 10, 12)
  np.testing.assert_array_equal(coeffs, bc.fit_linear_to_cheby(1.0, 1e-3, 5, 10, 12))
  r_o, phi_o = L.make_schedule_chebyshev(1000, coeffs, 5., 10.)
  r_p, phi_p = bc._schedule_and_jacobian(np.arange(1001), coeffs, 5., 10.)
  np.testing.assert_array_equal(r_o, r_p)
  np.testing.assert_array_equal(phi_o, phi_p)
  np.testing.assert_array_equal(bc.MakeSchedule_chebyshev(np.arange(1001), coeffs, 5., 10.), r_o)
  assert bc.make_trap_fxn(np.arange(1001), coeffs, 5., 10.)(0) == 5 and r_p[-1] == 10


def test_class_tables_reproduce_per_site_oracle_arithmetic():
  rs = np.random.RandomState(0)
  lt = rs.uniform(-1, 2, 7).astype(np.float32)
  h = rs.uniform(-2, 2, 7).astype(np.float32)
  thr, lut = ising.class_tables(lt, h)
  spins = I.random_spins((12, 12), 0.5, J.PRNGKey(2))
  nb = I.sum_neighbors(spins)
  slot = 8 * (spins > 0) + (nb + 4) // 2
  for t in range(1, 7):
    logits = I.flip_logits(spins, lt[t], h[t])
    np.testing.assert_array_equal(lut[t - 1, 0][slot].astype(np.float32), logits)
    np.testing.assert_array_equal(lut[t - 1, 1][slot].astype(np.float32), J.log_sigmoid32(-logits))
    np.testing.assert_array_equal(lut[t - 1, 2][slot].astype(np.float32), J.log_sigmoid32(logits))
    u = J.uniform(J.PRNGKey(t), spins.shape)
    bits = J.random_bits(J.PRNGKey(t), spins.shape)
    np.testing.assert_array_equal(u < J.sigmoid32(logits), (bits >> 9) < thr[t - 1][slot])


def test_even_odd_masks_and_errors():
  even, odd = ising.even_odd_masks((2, 2))
  np.testing.assert_array_equal(even, [[0, 1], [1, 0]])
  np.testing.assert_array_equal(odd, [[1, 0], [0, 1]])
  with pytest.raises(ValueError):
    ising.even_odd_masks((3, 2))
  with pytest.raises(ValueError):
    ising.get_train_step(optimizers.adam(1e-3), None, 2, 3, mode='sideways')
  with pytest.raises(TypeError):
    ising.estimate_gradient_rev(42)


def test_adam_matches_reference_update_rule():
  opt = optimizers.adam(optimizers.exponential_decay(0.1, 10, 0.5))
  x0 = {'a': np.array([1., -2.], np.float32), 'b': np.float32(0.5)}
  state = opt.init_fn(x0)
  m = v = 0.
  x = 0.5
  for i in range(5):
    g = {'a': np.array([0.1 * (i + 1), -0.3], np.float32), 'b': np.float32(0.2 * (i - 2))}
    state = opt.update_fn(i, g, state)
    gb = 0.2 * (i - 2)
    m = 0.1 * gb + 0.9 * m
    v = 0.001 * gb * gb + 0.999 * v
    x = x - 0.1 * 0.5 ** (i / 10) * (m / (1 - 0.9 ** (i + 1))) / (np.sqrt(v / (1 - 0.999 ** (i + 1))) + 1e-8)
  np.testing.assert_allclose(opt.params_fn(state)['b'], x, rtol=1e-5)
  clipped = optimizers.clip_grads({'a': np.array([3., 4.], np.float32)}, 1.0)
  np.testing.assert_allclose(clipped['a'], [0.6, 0.8], rtol=1e-6)
  same = optimizers.clip_grads({'a': np.array([3., 4.], np.float32)}, 10.0)
  np.testing.assert_allclose(same['a'], [3., 4.])


def test_optimizer_over_schedule_pytree():
  sched = ising.IsingSchedule(p10n.Constant(1.), p10n.Chebyshev(np.ones(8)))
  opt = optimizers.adam(1e-3)
  state = opt.init_fn(sched)
  grads = tree.tree_map(np.ones_like, sched)
  state = opt.update_fn(0, grads, state)
  new = opt.params_fn(state)
  assert isinstance(new, ising.IsingSchedule) and isinstance(new.field, p10n.Chebyshev)
  np.testing.assert_allclose(new.field.weights, 1 - 1e-3, rtol=1e-5)


@pytest.mark.parametrize('every', [1, 3, 5])
def test_checkpoint_scan_equals_scan(every):    # tests/control_flow_test.py:10-30
  def step(carry, x):
    y = carry + x
    return y, y
  xs = np.arange(30.)
  _, expected = control_flow.scan(step, np.zeros((3, 4)), xs)
  _, actual = control_flow.checkpoint_scan(step, np.zeros((3, 4)), xs, every)
  np.testing.assert_allclose(expected, actual)
  with pytest.raises(ValueError):
    control_flow.checkpoint_scan(step, np.zeros((3, 4)), xs, 7)


def test_value_and_jacfwd():                     # tests/gradients_test.py:15-48
  x = torch.tensor(np.random.RandomState(0).normal(size=(3, 3)))
  val, jac = gradients.value_and_jacfwd(torch.exp)(x)
  np.testing.assert_allclose(val, torch.exp(x))
  expected = torch.autograd.functional.jacobian(torch.exp, x)
  np.testing.assert_allclose(jac, expected)
  a, b = torch.tensor([[1., 2.], [3., 4.]], dtype=torch.float64), torch.tensor([0.5, -1.], dtype=torch.float64)
  val, jac = gradients.value_and_jacfwd(lambda t: t[0] @ t[1])([a, b])
  np.testing.assert_allclose(val, a @ b)
  np.testing.assert_allclose(jac[1], a)
  val, jac = gradients.value_and_jacfwd(lambda u, y: (u * y, u / y))(a, a + 1)
  np.testing.assert_allclose(val[1], a / (a + 1))
  assert jac[0].shape == (2, 2, 2, 2)


def test_unknown_callables_are_rejected_not_emulated():
  from noneq_opt_b200 import potentials, space
  with pytest.raises(NotImplementedError):
    potentials.as_energy(lambda position, r0=0.: 0.0)
  with pytest.raises(NotImplementedError):
    space.as_shift(lambda R, dR: R + dR)
  assert potentials.V_spring(2.0)(np.array([[3.0]]), r0=1.0) == 4.0


def test_multi_row_chebyshev_value_and_jacobian():
  """Chebyshev with weights [ndim, degree + 1] (parameterization.py:145): one polynomial per row."""
  rs = np.random.RandomState(2)
  w = (0.5 * rs.normal(size=(3, 5))).astype(np.float32)
  xs = np.linspace(0, 1, 17, dtype=np.float32)
  cheb = p10n.Chebyshev(w)
  y = cheb(xs)
  assert y.shape == (3, 17)
  for i in range(3):
    np.testing.assert_array_equal(y[i], p10n.Chebyshev(w[i])(xs))
  J = cheb.jacobian(xs)
  assert J.shape == (3, 17, 15)
  np.testing.assert_allclose(np.einsum('inv,v->in', J, w.reshape(-1)), y, atol=2e-6)
  bumped = w.copy(); bumped[1, 2] += 0.25
  dy = (p10n.Chebyshev(bumped)(xs) - y) / 0.25
  np.testing.assert_allclose(dy, J[:, :, 1 * 5 + 2], atol=2e-6)
  assert cheb.degree == 4 and cheb.n_variables == 15
