"""Oracle (test infrastructure): restatement of the reference's Langevin path in numpy.

Follows /root/reference/noneq_opt/barrier_crossing.py (and its verbatim duplicates in
simulate.py / potentials.py):
  potentials  V_spring potentials.py:47-52, V_biomolecule barrier_crossing.py:176-184,
              V_biomolecule_shifted :186-194
  brownian    :28-94  (_dist :66-71, init_fn :73-75, apply_fn :77-92)
  run_brownian_opt :97-143, MakeSchedule_chebyshev :145-151, fit_linear_to_cheby :168-174
  run_brownian_stationary_trap :227-241, mapped_brownian_eqm :249-256
  estimators  :204-224 (REINFORCE), :262-285 (logP), :288-311 (reparam), :313-333

Vectorised over the trajectory axis (what `jax.vmap` does in the reference); N=1, dim=1
as the reference potentials require (`position[0][0]`).  `dtype=np.float32` is the
parity oracle (one rounding per operation, reference operation order, no FMA;
transcendentals correctly rounded from binary64); `dtype=np.float64` is the "truth"
variant used only for error accounting (it still consumes the same f32 uniforms).

The force is the hand-derived reverse-mode derivative of the energy closure, in the
order jax's VJP rules produce it (log: g/x, exp: g*ans, integer_pow 2: g*(2*u)).
Gradients are given two ways: closed form (SURVEY Appendix B.1 -- positions carry no
derivative because dR is stop_gradient'ed, barrier_crossing.py:88) and by central
finite differences of the frozen-noise surrogate in binary64 (`fd_gradient`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import jaxlike as J
from .parameterization_ref import Chebyshev

F32 = np.float32
LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


# ----------------------------------------------------------------------------- potentials
@dataclass(frozen=True)
class Potential:
  kind: str            # 'spring' | 'biomolecule' | 'biomolecule_shifted'
  k_s: float
  kappa_l: float = 0.0
  kappa_r: float = 0.0
  x_m: float = 0.0
  delta_E: float = 0.0
  beta: float = 1.0

  def constants(self, d):
    """Python-float (binary64) folded constants, cast once to the working dtype --
    what jax's weak-typed python scalars become."""
    if self.kind == 'spring':
      return dict(hk=d(self.k_s / 2))
    return dict(hk=d(self.k_s / 2),
                inv_beta=d(1. / self.beta),
                cl=d(-0.5 * self.beta * self.kappa_l),
                cr=d(0.5 * self.beta * self.kappa_r),
                bde=d(self.beta * self.delta_E),
                xm=d(self.x_m),
                two_xm=d(2. * self.x_m))


def V_spring(k_s):
  return Potential('spring', k_s)


def V_biomolecule(kappa_l, kappa_r, x_m, delta_E, k_s, beta):
  return Potential('biomolecule', k_s, kappa_l, kappa_r, x_m, delta_E, beta)


def V_biomolecule_shifted(kappa_l, kappa_r, x_m, delta_E, k_s, beta):
  return Potential('biomolecule_shifted', k_s, kappa_l, kappa_r, x_m, delta_E, beta)


def _exp(x, d):
  return J.exp32(x) if d is F32 else np.exp(x)


def _log(x, d):
  return J.log32(x) if d is F32 else np.log(x)


def _molecule_terms(pot, x, d):
  c = pot.constants(d)
  if pot.kind == 'biomolecule':
    ul = (x + c['xm']).astype(d)
    ur = (x - c['xm']).astype(d)
  else:
    ul = x
    ur = (x - c['two_xm']).astype(d)
  a = (c['cl'] * (ul * ul).astype(d)).astype(d)
  v = ((c['cr'] * (ur * ur).astype(d)).astype(d) + c['bde']).astype(d)
  A = _exp(a, d)
  B = _exp(-v, d)
  return c, ul, ur, A, B


def energy(pot, x, r0, dtype=F32):
  """E(position, r0) -- barrier_crossing.py:177-183 / potentials.py:48-51."""
  d = dtype
  x = np.asarray(x, d)
  r0 = np.asarray(r0, d)
  c = pot.constants(d)
  u = (x - r0).astype(d)
  Es = (c['hk'] * (u * u).astype(d)).astype(d)
  if pot.kind == 'spring':
    return Es
  c, ul, ur, A, B = _molecule_terms(pot, x, d)
  Em = ((-c['inv_beta']) * _log((A + B).astype(d), d)).astype(d)
  return (Em + Es).astype(d)


def force(pot, x, r0, dtype=F32):
  """F = -dE/dx  (jax_md.quantity.canonicalize_force -> -grad(E), barrier_crossing.py:60)."""
  d = dtype
  x = np.asarray(x, d)
  r0 = np.asarray(r0, d)
  c = pot.constants(d)
  u = (x - r0).astype(d)
  dEs = (c['hk'] * (d(2) * u).astype(d)).astype(d)
  if pot.kind == 'spring':
    return (-dEs).astype(d)
  c, ul, ur, A, B = _molecule_terms(pot, x, d)
  g = ((-c['inv_beta']) / (A + B).astype(d)).astype(d)          # cotangent of A+B
  ga = (g * A).astype(d)                                        # cotangent of a
  gb = (g * B).astype(d)                                        # cotangent of b = -v
  dxa = ((ga * c['cl']).astype(d) * (d(2) * ul).astype(d)).astype(d)
  dxb = (((-gb) * c['cr']).astype(d) * (d(2) * ur).astype(d)).astype(d)
  dEm = (dxa + dxb).astype(d)
  return (-(dEm + dEs).astype(d)).astype(d)


# ----------------------------------------------------------------------------- integrator
@dataclass(frozen=True)
class StepConstants:
  nu: object
  dt: object
  sigma: object
  log_norm: object


def step_constants(dt, kT, mass, gamma, dtype=F32):
  """brownian._dist :66-71 -- nu, variance, sigma; Normal's log normaliser (tfp normal.py)."""
  d = dtype
  nu = d(1) / (d(mass) * d(gamma))
  variance = ((d(2) * d(kT)) * d(dt)) * nu
  sigma = np.sqrt(d(variance))
  if d is F32:
    log_sigma = J.log32(sigma)
    log_norm = (F32(LOG_SQRT_2PI) + log_sigma).astype(F32)
  else:
    log_norm = d(LOG_SQRT_2PI) + np.log(sigma)
  return StepConstants(d(nu), d(dt), d(sigma), d(log_norm))


def apply_shift(x, dR, shift, d):
  """jax_md.space: free -> R + dR ; periodic(box) -> mod(R + dR, box)."""
  xn = (x + dR).astype(d)
  if shift is None or shift == 'free':
    return xn
  kind, box = shift
  assert kind == 'periodic'
  return np.mod(xn, d(box)).astype(d)


def brownian_apply(pot, sc, x, rng, r0, shift='free', dtype=F32):
  """One apply_fn (:77-92) for a batch: returns (x_new, rng_new, log_prob, dR, mean)."""
  d = dtype
  F = force(pot, x, r0, d)
  mean = ((F * sc.nu).astype(d) * sc.dt).astype(d)
  # key, split = jax.random.split(state.rng)   (:79)  -- 2 threefry blocks
  a0, a1 = J.threefry2x32(rng[:, 0], rng[:, 1], np.uint32(0), np.uint32(2))
  b0, b1 = J.threefry2x32(rng[:, 0], rng[:, 1], np.uint32(1), np.uint32(3))
  new_rng = np.stack([a0, b0], 1)
  # dist.sample(seed=split): jax.random.normal(split, [1,1,1]) -> block (0,0).y0
  bits, _ = J.threefry2x32(a1, b1, np.uint32(0), np.uint32(0))
  xi = J.normal_from_bits(bits).astype(d)
  dR = ((xi * sc.sigma).astype(d) + mean).astype(d)
  # Normal.log_prob: -0.5*squared_difference(x/scale, loc/scale) - (0.5 log 2pi + log scale)
  z = ((dR / sc.sigma).astype(d) - (mean / sc.sigma).astype(d)).astype(d)
  lp = ((d(-0.5) * (z * z).astype(d)).astype(d) - sc.log_norm).astype(d)
  xn = apply_shift(x, dR, shift, d)
  return xn, new_rng, lp, dR, mean


def brownian_apply_nd(pot, sc, position, rng, r0, shift='free', dtype=F32):
  """apply_fn (simulate.py:82-97 = barrier_crossing.py:77-92) on ONE rank-generic state: position [N, dim],
  rng uint32[2].  The reference's energy closures read position[0][0] only (potentials.py:27-52), so -grad E is
  zero everywhere else.  dist.sample(seed=split) = jax.random.normal(split, [1, N, dim]) * sigma + mean;
  log_prob = sum over all elements of Normal.log_prob (:90).  Returns (position', rng', log_prob, dR)."""
  d = dtype
  position = np.asarray(position, d)
  N, dim = position.shape
  F = np.zeros((N, dim), d)
  F[0, 0] = force(pot, position[0:1, 0], r0, d)[0]
  mean = ((F * sc.nu).astype(d) * sc.dt).astype(d)
  (new_rng, sub) = J.split(np.asarray(rng, np.uint32), 2)
  xi = J.normal(sub, (1, N, dim)).astype(d).reshape(N, dim)
  dR = ((xi * sc.sigma).astype(d) + mean).astype(d)
  z = ((dR / sc.sigma).astype(d) - (mean / sc.sigma).astype(d)).astype(d)
  lp = ((d(-0.5) * (z * z).astype(d)).astype(d) - sc.log_norm).astype(d)
  xn = apply_shift(position.reshape(-1), dR.reshape(-1), shift, d).reshape(N, dim)
  return xn, new_rng, d(lp.astype(np.float64).sum()), dR


def make_schedule_chebyshev(S, coeffs, r0_init, r0_final, dtype=F32):
  """MakeSchedule_chebyshev(arange(S+1), ...) :145-151 -> r[S+1]; also the [S-1, D] Jacobian.

  `coeffs` may also be any parameterisation object of oracle.parameterization_ref (callable with
  `.jacobian`), e.g. PiecewiseLinear -- the "piecewise" protocols of the north star: the reference
  builds the table as `p10n.<Class>(...)(scaled_timevec)` in exactly the same way."""
  d = dtype
  tv = np.arange(S + 1)[1:-1]
  scaled = ((tv - 1).astype(d) / d(tv[-1] - 1)).astype(d)
  proto = coeffs if hasattr(coeffs, 'jacobian') else Chebyshev(coeffs, d)
  vals = proto(scaled)
  r = np.concatenate([[d(r0_init)], vals, [d(r0_final)]]).astype(d)
  return r, proto.jacobian(scaled)


def fit_linear_to_cheby(end_time, dt, r0_init, r0_final, degree):
  """:168-174 (standard-basis fit coefficients fed to the monic basis, as the reference does)."""
  S = int(end_time / dt)
  slope = (r0_final - r0_init) / S
  vals = slope * np.arange(1, S) + r0_init
  xs = (np.arange(1, S) - 1) / (S - 2)
  p = np.polynomial.Chebyshev.fit(xs, vals, degree, domain=[0, 1])
  return np.asarray(p.coef, F32)


def trajectory_keys(seed, B):
  """seeds = jax.random.split(seed, B) (:221) then per trajectory `key, split = split(key)` (:133)."""
  seeds = J.split(seed, B)
  _, b1 = J.threefry2x32(seeds[:, 0], seeds[:, 1], np.uint32(1), np.uint32(3))
  _, a1 = J.threefry2x32(seeds[:, 0], seeds[:, 1], np.uint32(0), np.uint32(2))
  return seeds, np.stack([a1, b1], 1)     # (per-trajectory seeds, initial state.rng)


def run_brownian_opt(pot, r_table, x0, rng0, Neq, S, dt, kT, mass=1.0, gamma=1.0,
                     shift='free', dtype=F32, record=True):
  """run_brownian_opt :97-143 for a batch, given the trap table r[S+1] and state rng [B,2].

  Returns dict(positions [B,S+1] (or None), log_probs [B,S], works [B,S+1], x_final,
  dR [B,S], mean [B,S]) -- dR/mean feed the closed-form gradients.
  """
  d = dtype
  sc = step_constants(dt, kT, mass, gamma, d)
  x = np.asarray(x0, d).copy()
  rng = np.asarray(rng0, np.uint32).copy()
  B = x.shape[0]
  r = np.asarray(r_table, d)
  for _ in range(Neq):                                   # equilibrate :115-121
    x, rng, _, _, _ = brownian_apply(pot, sc, x, rng, r[0], shift, d)
  x_eq = x.copy()
  pos = np.zeros((B, S + 1), d) if record else None
  lps = np.zeros((B, S), d)
  works = np.zeros((B, S + 1), d)
  dRs = np.zeros((B, S), d)
  means = np.zeros((B, S), d)
  if record:
    pos[:, 0] = x_eq
  for step in range(1, S + 1):                           # scan_fn :127-131
    dW = (energy(pot, x, r[step], d) - energy(pot, x, r[step - 1], d)).astype(d)
    x, rng, lp, dR, mean = brownian_apply(pot, sc, x, rng, r[step], shift, d)
    works[:, step] = dW
    lps[:, step - 1] = lp
    dRs[:, step - 1] = dR
    means[:, step - 1] = mean
    if record:
      pos[:, step] = x
  return dict(positions=pos, log_probs=lps, works=works, x_final=x, x_eq=x_eq,
              dR=dRs, mean=means, sc=sc)


def reduce_sum(a, axis, dtype):
  """`.sum()` of the reference (barrier_crossing.py:208-209).  XLA's reduction order is not
  specified, so the oracle takes the exact sum of the f32 terms (binary64 accumulation)
  rounded once to `dtype` -- every f32 summation order deviates from this by at most its own
  rounding noise."""
  return np.asarray(a, dtype).astype(np.float64).sum(axis).astype(dtype)


def estimate_gradient(pot, coeffs, x0, seed, r0_init, r0_final, Neq, S, dt, kT, mass, gamma,
                      shift='free', dtype=F32, rng0=None, record=False):
  """estimate_gradient_boltzinit* :216-224/:277-285/:303-311 -- all three terms at once.

  Returns dict with grad (REINFORCE, = logP + reparam), grad_logP, grad_reparam (each
  mean over the batch, [D]), per-trajectory grads, estimator, tot_log_prob, total_work,
  and the trajectory record.
  """
  d = dtype
  B = len(x0)
  if rng0 is None:
    _, rng0 = trajectory_keys(seed, B)
  r, phi = make_schedule_chebyshev(S, coeffs, r0_init, r0_final, d)
  out = run_brownian_opt(pot, r, x0, rng0, Neq, S, dt, kT, mass, gamma, shift, d, record)
  sc = out['sc']
  total_work = reduce_sum(out['works'], 1, d)
  tot_lp = reduce_sum(out['log_probs'], 1, d)
  ks = d(pot.k_s)
  # d logp_k / d r_k = z_k/sigma * nu*dt*k_s  with z_k = (dR_k - mean_k)/sigma   (k = 1..S-1)
  z = ((out['dR'] / sc.sigma).astype(d) - (out['mean'] / sc.sigma).astype(d)).astype(d)
  a = (((z / sc.sigma).astype(d) * sc.nu).astype(d) * sc.dt * ks).astype(d)[:, :S - 1]
  # d W / d r_k = -k_s (x_{k-1} - r_k) + k_s (x_k - r_k) = k_s (x_k - x_{k-1}) = k_s dR_k, except where a
  # periodic shift wrapped the position (the trap energy itself is not periodic in the reference)
  if shift in (None, 'free'):
    step_dx = out['dR']
  else:
    if not record:
      raise ValueError('periodic gradients need record=True (positions)')
    step_dx = (out['positions'][:, 1:] - out['positions'][:, :-1]).astype(d)
  b = (ks * step_dx).astype(d)[:, :S - 1]
  A = (a.astype(np.float64) @ phi.astype(np.float64))     # [B, D]
  Bm = (b.astype(np.float64) @ phi.astype(np.float64))
  g_logp = (total_work.astype(np.float64)[:, None] * A)
  g_rep = Bm
  return dict(grad=(g_logp + g_rep).mean(0), grad_logP=g_logp.mean(0), grad_reparam=g_rep.mean(0),
              per_traj_logP=g_logp, per_traj_reparam=g_rep,
              estimator=(tot_lp * total_work + total_work).astype(d),
              tot_log_prob=tot_lp, total_work=total_work, r=r, phi=phi, traj=out)


def fd_gradient(pot, coeffs, x0, rng0, r0_init, r0_final, Neq, S, dt, kT, mass, gamma, eps=1e-4):
  """Brute-force check of the closed forms: central differences, in binary64, of the surrogate
  sum_n [ sum_k logp_k(c) * stop_grad(W_n) + W_n(c) ] / B with positions and dR frozen
  (that is what stop_gradient at barrier_crossing.py:88 makes of it)."""
  d = np.float64
  coeffs = np.asarray(coeffs, d)
  r, _ = make_schedule_chebyshev(S, coeffs, r0_init, r0_final, d)
  base = run_brownian_opt(pot, r, x0, rng0, Neq, S, dt, kT, mass, gamma, 'free', d, True)
  sc = base['sc']
  pos = base['positions']
  dR = base['dR']
  W0 = base['works'].sum(1)

  def surrogate(c, which):
    rr, _ = make_schedule_chebyshev(S, c, r0_init, r0_final, d)
    lp = np.zeros(len(x0))
    W = np.zeros(len(x0))
    for step in range(1, S + 1):
      x = pos[:, step - 1]
      W += energy(pot, x, rr[step], d) - energy(pot, x, rr[step - 1], d)
      mean = force(pot, x, rr[step], d) * sc.nu * sc.dt
      zz = dR[:, step - 1] / sc.sigma - mean / sc.sigma
      lp += -0.5 * zz * zz - sc.log_norm
    if which == 'logP':
      return (lp * W0).mean()
    if which == 'reparam':
      return W.mean()
    return (lp * W0 + W).mean()

  out = {}
  for which in ('reinforce', 'logP', 'reparam'):
    g = np.zeros_like(coeffs)
    for j in range(len(coeffs)):
      e = np.zeros_like(coeffs)
      e[j] = eps
      g[j] = (surrogate(coeffs + e, which) - surrogate(coeffs - e, which)) / (2 * eps)
    out[which] = g
  return out


def run_brownian_stationary_trap(pot, x0, r0, rng0, S, dt, kT, mass, gamma, shift='free',
                                 dtype=F32, record=False):
  """:227-241 -- same integrator, constant r0; returns final positions (and all if record)."""
  d = dtype
  sc = step_constants(dt, kT, mass, gamma, d)
  x = np.asarray(x0, d).copy()
  rng = np.asarray(rng0, np.uint32).copy()
  pos = np.zeros((x.shape[0], S), d) if record else None
  for step in range(S):
    x, rng, _, _, _ = brownian_apply(pot, sc, x, rng, d(r0), shift, d)
    if record:
      pos[:, step] = x
  return (x, pos) if record else x


def mapped_brownian_eqm(pot, B, init_position, r0_init, seed, sim_length, dt, kT, mass, gamma,
                        dtype=F32):
  """mapped_brownian_eqm :249-256 with the given dt/kT/mass/gamma (the reference ignores its
  arguments and hard-codes 1e-6,1e-5,1,1 at :250 -- callers of this oracle pass what they mean)."""
  _, rng0 = trajectory_keys(seed, B)
  x0 = np.full(B, init_position, dtype)
  return run_brownian_stationary_trap(pot, x0, r0_init, rng0, sim_length, dt, kT, mass, gamma,
                                      'free', dtype)
