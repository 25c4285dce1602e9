"""Oracle (test infrastructure): restatement of noneq_opt/ising.py in numpy.

Follows /root/reference/noneq_opt/ising.py line by line, on whole-lattice arrays exactly as
the reference does (rolls, masks, full-lattice Bernoulli draw of which half is discarded):
  random_spins :71-72      sum_neighbors :75-83    energy :85-86      log_prob :89-90
  flip_logits :93-96       even_odd_masks :99-106  masked_update :109-126
  update :129-150          simulate_ising :153-173
  total_entropy_production :239-249   total_work :252-255
  estimate_gradient_{rev,fwd}[_REINFORCE|_logP|_reparam|_track_terms] :179-236, :294-599

Arithmetic is float32 in reference order (see oracle/__init__.py for the convention);
reductions (`.sum()`, `.mean()`) are the exact sum of the f32 terms rounded once, because
XLA's reduction order is unspecified.  `dtype=np.float64` is the truth variant.

Gradients: `closed_form_gradient` implements SURVEY Appendix B.2 (spins carry no
derivative -- flips are integer samples); `fd_gradient` differentiates the frozen-flip
surrogate by central differences in binary64 as the brute-force cross-check.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import jaxlike as J

F32 = np.float32


class IsingSummary(NamedTuple):
  work: np.ndarray
  dissipated_heat: np.ndarray
  forward_log_prob: np.ndarray
  reverse_log_prob: np.ndarray
  entropy_production: np.ndarray
  magnetization: np.ndarray
  energy: np.ndarray


def random_spins(shape, p, seed):
  """:71-72."""
  n = int(np.prod(shape))
  return (2 * J.bernoulli_probs_sample(seed, p, (n,)).reshape(shape) - 1).astype(np.int32)


def sum_neighbors(spins):
  """:75-83 -- periodic sum over +-1 rolls on every axis."""
  out = 0
  for axis in range(spins.ndim):
    for sh in (-1, 1):
      out = out + np.roll(spins, sh, axis)
  return out


def _rsum(a, d):
  return d(np.asarray(a).astype(np.float64).sum())


def energy(spins, field, dtype=F32):
  """:85-86 -- (-spins * (sum_neighbors/2 + field)).sum()."""
  d = dtype
  nb = sum_neighbors(spins).astype(d) / d(2)
  t = ((-spins).astype(d) * (nb + d(field)).astype(d)).astype(d)
  return _rsum(t, d)


def _exp(x, d):
  return J.exp32(x) if d is F32 else np.exp(np.asarray(x, d))


def log_prob(spins, log_temp, field, dtype=F32):
  """:89-90."""
  d = dtype
  return d(-energy(spins, field, d) / d(_exp(d(log_temp), d)))


def flip_logits(spins, log_temp, field, dtype=F32):
  """:93-96 -- ((-2*spins) * (nbr + field)) * exp(-log_temp)."""
  d = dtype
  t = ((-2 * spins).astype(d) * (sum_neighbors(spins).astype(d) + d(field)).astype(d)).astype(d)
  return (t * d(_exp(-d(log_temp), d))).astype(d)


def even_odd_masks(shape):
  """:99-106 -- NB "even" mask is 1 where the index sum is ODD (tests/ising_test.py:47-51)."""
  for s in shape:
    if s % 2:
      raise ValueError(f'All entries in `shape` must be even; got {shape}.')
  grid = np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')
  mask_even = sum(grid) % 2
  return mask_even, 1 - mask_even


def _sigmoid(x, d):
  return J.sigmoid32(x) if d is F32 else 1.0 / (1.0 + np.exp(-x))


def _bern_log_prob(logits, v, d):
  if d is F32:
    return J.bernoulli_logits_log_prob(logits, v)
  ls = lambda x: -(np.maximum(-x, 0) + np.log1p(np.exp(-np.abs(x))))
  v = v.astype(d)
  return ls(-logits) * (1 - v) + ls(logits) * v


def masked_update(spins, log_temp, field, mask, seed, dtype=F32):
  """:109-126.  Returns (new_spins, fwd_lp, rev_lp, flips, logits)."""
  d = dtype
  logits = flip_logits(spins, log_temp, field, d)
  u = J.uniform(seed, spins.shape)                     # full-lattice draw (f32 uniforms)
  flips = (u < _sigmoid(logits.astype(F32) if d is F32 else logits, d).astype(u.dtype)
           ).astype(np.int32) * mask
  fwd = _rsum(_bern_log_prob(logits, flips, d) * mask.astype(d), d)
  new_spins = spins * (1 - 2 * flips)
  rev_logits = flip_logits(new_spins, log_temp, field, d)
  rev = _rsum(_bern_log_prob(rev_logits, flips, d) * mask.astype(d), d)
  return new_spins, fwd, rev, flips, logits


def update(spins, old_params, new_params, seed, dtype=F32, trace=None):
  """:129-150.  params = (log_temp, field).  Returns (new_spins, summary of scalars)."""
  d = dtype
  seed_even, seed_odd = J.split(seed, 2)
  mask_even, mask_odd = even_odd_masks(spins.shape)
  lt, h = new_params
  initial_energy = energy(spins, old_params[1], d)
  intermediate_energy = energy(spins, h, d)
  s1, ef, er, f_even, l_even = masked_update(spins, lt, h, mask_even, seed_even, d)
  s2, of, orv, f_odd, l_odd = masked_update(s1, lt, h, mask_odd, seed_odd, d)
  final_energy = energy(s2, h, d)
  fwd = d(ef + of)
  rev = d(er + orv)
  if trace is not None:
    trace.append(dict(spins_before=spins, flips=(f_even, f_odd), logits=(l_even, l_odd),
                      masks=(mask_even, mask_odd), spins_mid=s1, spins_after=s2))
  summary = IsingSummary(work=d(intermediate_energy - initial_energy),
                         dissipated_heat=d(intermediate_energy - final_energy),
                         forward_log_prob=fwd, reverse_log_prob=rev,
                         entropy_production=d(fwd - rev),
                         magnetization=d(np.asarray(s2).astype(np.float64).mean()),
                         energy=final_energy)
  return s2, summary


def simulate_ising(log_temp, field, initial_spins, seed, return_states=False, dtype=F32,
                   trace=None):
  """:153-173.  log_temp/field are the [T] schedule tables; T-1 sweeps with split(seed, T-1)."""
  T = len(field)
  seeds = J.split(seed, T - 1)
  spins = np.asarray(initial_spins)
  rows, states = [], []
  for t in range(1, T):
    spins, s = update(spins, (log_temp[t - 1], field[t - 1]), (log_temp[t], field[t]),
                      seeds[t - 1], dtype, trace)
    rows.append(s)
    if return_states:
      states.append(spins.copy())
  summary = IsingSummary(*[np.array([getattr(r, f) for r in rows], dtype) for f in IsingSummary._fields])
  return (spins, summary, np.stack(states)) if return_states else (spins, summary)


def total_entropy_production(initial_spins, final_spins, log_temp, field, summary, dtype=F32):
  """:239-249 -- sum(ep) + log_prob(initial; params[0]) - log_prob(final; params[T-1])."""
  d = dtype
  traj = _rsum(summary.entropy_production, d)
  endpoint = d(log_prob(initial_spins, log_temp[0], field[0], d) -
               log_prob(final_spins, log_temp[-1], field[-1], d))
  return d(endpoint + traj)


def total_work(initial_spins, final_spins, log_temp, field, summary, dtype=F32):
  """:252-255."""
  return _rsum(summary.work, dtype)


LOSSES = dict(total_entropy_production=total_entropy_production, total_work=total_work)


# ----------------------------------------------------------------------------- gradients
def closed_form_gradient(loss_name, log_temp, field, J_lt, J_h, initial_spins, seed):
  """Appendix B.2, in binary64 on the f32 trajectory.  J_lt [T, D1], J_h [T, D2] are the
  schedule Jacobians d(log_temp_t)/d(w), d(field_t)/d(w).  Returns dict with the three
  term families of ising.py:559-599 as (grad_log_temp_vars, grad_field_vars) pairs, the loss
  and the summary."""
  trace = []
  lt64 = np.asarray(log_temp, np.float64)
  h64 = np.asarray(field, np.float64)
  T = len(field)
  final, summary = simulate_ising(log_temp, field, initial_spins, seed, trace=trace)
  loss = float(LOSSES[loss_name](initial_spins, final, log_temp, field, summary))
  dfwd_dh = np.zeros(T); dfwd_dl = np.zeros(T)
  dloss_dh = np.zeros(T); dloss_dl = np.zeros(T)
  for t in range(1, T):
    tr = trace[t - 1]
    einv = np.exp(-lt64[t])
    M_prev = float(tr['spins_before'].sum())
    if loss_name == 'total_work':
      dloss_dh[t] += -M_prev
      dloss_dh[t - 1] += M_prev
    cur = tr['spins_before']
    for half in range(2):
      mask = tr['masks'][half].astype(bool)
      flips = tr['flips'][half].astype(np.float64)
      s = cur.astype(np.float64)
      n = sum_neighbors(cur).astype(np.float64)
      ell = -2 * s * (n + h64[t]) * einv
      sig = 1 / (1 + np.exp(-ell))
      dfwd_dh[t] += ((flips - sig) * (-2 * s * einv))[mask].sum()
      dfwd_dl[t] += ((flips - sig) * (-ell))[mask].sum()
      if loss_name == 'total_entropy_production':
        fl = mask & (flips > 0)
        dloss_dh[t] += (-2 * s * einv)[fl].sum()
        dloss_dl[t] += (-ell)[fl].sum()
      cur = tr['spins_mid'] if half == 0 else tr['spins_after']
  if loss_name == 'total_entropy_production':
    s0 = np.asarray(initial_spins).astype(np.float64)
    sf = np.asarray(final).astype(np.float64)
    M0, Mf = s0.sum(), sf.sum()
    B0 = 0.5 * (s0 * sum_neighbors(np.asarray(initial_spins))).sum()
    Bf = 0.5 * (sf * sum_neighbors(np.asarray(final))).sum()
    e0, ef = np.exp(-lt64[0]), np.exp(-lt64[-1])
    dloss_dh[0] += M0 * e0
    dloss_dl[0] += -(B0 + h64[0] * M0) * e0
    dloss_dh[-1] += -Mf * ef
    dloss_dl[-1] += (Bf + h64[-1] * Mf) * ef
  J_lt = np.asarray(J_lt, np.float64)
  J_h = np.asarray(J_h, np.float64)
  logP = (J_lt.T @ (loss * dfwd_dl), J_h.T @ (loss * dfwd_dh))
  rep = (J_lt.T @ dloss_dl, J_h.T @ dloss_dh)
  return dict(grad=(logP[0] + rep[0], logP[1] + rep[1]), grad_logP=logP, grad_reparam=rep,
              loss=loss, summary=summary, final_spins=final,
              per_step=dict(dfwd_dh=dfwd_dh, dfwd_dl=dfwd_dl, dloss_dh=dloss_dh, dloss_dl=dloss_dl))


def fd_gradient(loss_name, schedule_fn, w_lt, w_h, times, initial_spins, seed, eps=1e-5):
  """Central differences (binary64) of  sum(fwd_lp)(w) * stop_grad(loss) + loss(w)  with the
  flips frozen at their sampled values.  schedule_fn(w_lt, w_h, times) -> (log_temp[T], field[T])."""
  d = np.float64
  lt0, h0 = schedule_fn(np.asarray(w_lt, d), np.asarray(w_h, d), times)
  trace = []
  final, summ0 = simulate_ising(np.asarray(lt0, F32), np.asarray(h0, F32), initial_spins, seed,
                                trace=trace)
  T = len(times)

  def replay(wl, wh):
    lt, h = schedule_fn(wl, wh, times)
    rows, fwd_tot = [], 0.0
    for t in range(1, T):
      tr = trace[t - 1]
      e_init = energy(tr['spins_before'], h[t - 1], d)
      e_mid = energy(tr['spins_before'], h[t], d)
      cur, f, r = tr['spins_before'], 0.0, 0.0
      for half in range(2):
        mask = tr['masks'][half]
        flips = tr['flips'][half]
        lg = flip_logits(cur, lt[t], h[t], d)
        f += (_bern_log_prob(lg, flips, d) * mask).sum()
        new = cur * (1 - 2 * flips)
        rl = flip_logits(new, lt[t], h[t], d)
        r += (_bern_log_prob(rl, flips, d) * mask).sum()
        cur = new
      e_fin = energy(cur, h[t], d)
      rows.append(IsingSummary(e_mid - e_init, e_mid - e_fin, f, r, f - r, cur.mean(), e_fin))
      fwd_tot += f
    summ = IsingSummary(*[np.array([getattr(q, fl) for q in rows]) for fl in IsingSummary._fields])
    loss = LOSSES[loss_name](initial_spins, final, lt, h, summ, d)
    return fwd_tot, loss

  _, loss0 = replay(np.asarray(w_lt, d), np.asarray(w_h, d))

  def surrogate(wl, wh, which):
    f, l = replay(wl, wh)
    return {'reinforce': f * loss0 + l, 'logP': f * loss0, 'reparam': l}[which]

  out = {}
  for which in ('reinforce', 'logP', 'reparam'):
    gl = np.zeros(len(w_lt)); gh = np.zeros(len(w_h))
    for j in range(len(w_lt)):
      e = np.zeros(len(w_lt)); e[j] = eps
      gl[j] = (surrogate(w_lt + e, w_h, which) - surrogate(w_lt - e, w_h, which)) / (2 * eps)
    for j in range(len(w_h)):
      e = np.zeros(len(w_h)); e[j] = eps
      gh[j] = (surrogate(w_lt, w_h + e, which) - surrogate(w_lt, w_h - e, which)) / (2 * eps)
    out[which] = (gl, gh)
  return out
