"""Brownian (overdamped Langevin) simulations of a particle dragged by a moving trap, and the
gradient estimators built on them -- B200 drop-in for `noneq_opt/barrier_crossing.py`.

Same names, argument order and return nesting as the reference:
  BrownianState :24-26          brownian :28-94            run_brownian_opt :97-143
  MakeSchedule_chebyshev :145   make_trap_fxn :153         theoretical_opt :159
  make_theoretical_trap_fxn :162   fit_linear_to_cheby :168
  V_biomolecule :176            V_biomolecule_shifted :186
  single_estimate_boltzinit / estimate_gradient_boltzinit :204-224
  run_brownian_stationary_trap :227   single_brownian_eqm :243   mapped_brownian_eqm :249
  *_logP :262-285   *_reparam :288-311   *_REINFORCE_recordall :313-333

All simulation work is done by the sm_100a kernels in csrc/langevin.cu through the C ABI of
include/noneq_b200.h; arrays come back as CUDA torch tensors.  Differences from the reference,
all forced by the absence of a tracing compiler (documented in DESIGN.md):
  * `energy_fn` / `shift` must be the descriptor objects of `potentials` / `space`;
  * functions accept a batch of keys `[B, 2]` directly where the reference uses `jax.vmap`;
  * `positions` of the gradient estimators are recorded only when they fit
    `MAX_AUTO_POSITION_BYTES` (or when `record_positions=` says so) -- at config-5 scale they
    would be 400 GB.
"""
from __future__ import annotations

import collections
import ctypes as C
import functools
import os

import numpy as np
import torch

from . import _lib
from . import control_flow
from . import parameterization as p10n
from . import random as nqrandom
from .potentials import V_biomolecule, V_biomolecule_shifted, V_spring, as_energy  # noqa: F401
from .space import as_shift

MAX_AUTO_POSITION_BYTES = 1 << 30
_MATH_MODES = {'strict': _lib.NQ_MATH_STRICT, 'fast': _lib.NQ_MATH_FAST}
_math_mode = _MATH_MODES[os.environ.get('NQ_MATH_MODE', 'strict')]


def set_math_mode(mode: str):
  """'strict' (reference operation order, IEEE div/sqrt, no FMA) or 'fast' (MUFU + FMA)."""
  global _math_mode
  _math_mode = _MATH_MODES[mode]


def get_math_mode() -> str:
  return {v: k for k, v in _MATH_MODES.items()}[_math_mode]


class BrownianState(collections.namedtuple('BrownianState', 'position mass rng log_prob')):
  pass


# ------------------------------------------------------------------------------ helpers
def _desc(energy_fn, shift, steps, Neq, dt, kT, mass, gamma):
  d = _lib.NqLangevinDesc()
  as_energy(energy_fn).fill(d)
  sh = as_shift(shift)
  d.shift_kind, d.box = sh.kind, float(sh.box)
  d.math_mode = _math_mode
  d.dt, d.kT, d.gamma, d.mass = float(dt), float(kT), float(gamma), float(mass)
  d.steps, d.eq_steps = int(steps), int(Neq)
  return d


def _dev_f32(a, device):
  if isinstance(a, torch.Tensor):
    return a.to(device=device, dtype=torch.float32).contiguous()
  return torch.from_numpy(np.array(a, dtype=np.float32, order="C")).to(device)


def _is_batched_key(key):
  return np.ndim(key) == 2 if not isinstance(key, torch.Tensor) else key.dim() == 2


def _simulate(energy_fn, shift, r_table, phi, x0, seeds, S, Neq, dt, kT, mass, gamma,
              positions_stride=None, step_outputs=False):
  """Core launcher.  x0 [n] f32 cuda, seeds [n,2] int32 cuda, r_table [S+1], phi [S-1,D] or None.

  Returns dict of CUDA tensors (+ float64 `moments`, `grad_sums` when phi is given)."""
  device = x0.device
  n = x0.shape[0]
  lib = _lib.lib()
  desc = _desc(energy_fn, shift, S, Neq, dt, kT, mass, gamma)
  r_dev = _dev_f32(r_table, device)
  if r_dev.numel() != S + 1:
    raise ValueError(f'trap table must have simulation_steps+1 = {S + 1} entries, got {r_dev.numel()}')
  D = 0 if phi is None else int(phi.shape[1])
  ws_bytes = lib.nq_langevin_workspace_bytes(n, D)
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
  out = dict(work=torch.empty(n, dtype=torch.float32, device=device),
             logp=torch.empty(n, dtype=torch.float32, device=device),
             x_final=torch.empty(n, dtype=torch.float32, device=device),
             moments=torch.zeros(_lib.NQ_MOM_COUNT, dtype=torch.float64, device=device))
  st = _lib.stream_arg()
  if phi is None:
    pos = sw = sl = None
    stride = 1
    if positions_stride:
      stride = int(positions_stride)
      pos = torch.empty((S // stride + 1, n), dtype=torch.float32, device=device)
    if step_outputs:
      sw = torch.empty((S + 1, n), dtype=torch.float32, device=device)
      sl = torch.empty((S, n), dtype=torch.float32, device=device)
    _lib.check(lib.nq_langevin_forward(C.byref(desc), _lib.ptr(r_dev), _lib.ptr(x0), _lib.ptr(seeds), n,
                                       _lib.ptr(out['work']), _lib.ptr(out['logp']), _lib.ptr(out['x_final']),
                                       _lib.ptr(pos), stride, _lib.ptr(sw), _lib.ptr(sl),
                                       _lib.ptr(out['moments']), _lib.ptr(ws), ws_bytes, st))
    out.update(positions_tm=pos, step_works_tm=sw, step_logps_tm=sl)
  else:
    phi_dev = _dev_f32(phi, device)
    if phi_dev.shape[0] != S - 1:
      raise ValueError(f'phi must have simulation_steps-1 = {S - 1} rows, got {phi_dev.shape[0]}')
    out['estimator'] = torch.empty(n, dtype=torch.float32, device=device)
    # [grad sums (2 D) | moments (8)] in ONE float64 buffer: the multi-GPU combine is a single in-place all-reduce
    flat = torch.zeros(2 * D + _lib.NQ_MOM_COUNT, dtype=torch.float64, device=device)
    out['sums_flat'], out['grad_sums'], out['moments'] = flat, flat[:2 * D].view(2, D), flat[2 * D:]
    out['positions_tm'] = None
    rc = _lib.NQ_ERR_UNSUPPORTED
    if positions_stride:
      # small ensembles: the fused pass records the trajectories too (one pass instead of two)
      stride = int(positions_stride)
      pos = torch.empty((S // stride + 1, n), dtype=torch.float32, device=device)
      rc = lib.nq_langevin_grad_record(C.byref(desc), _lib.ptr(r_dev), _lib.ptr(phi_dev), D, _lib.ptr(x0), _lib.ptr(seeds), n,
                                       _lib.ptr(out['work']), _lib.ptr(out['logp']), _lib.ptr(out['estimator']),
                                       _lib.ptr(out['x_final']), _lib.ptr(pos), stride, _lib.ptr(out['grad_sums']),
                                       _lib.ptr(out['moments']), _lib.ptr(ws), ws_bytes, st)
      if rc == 0:
        out['positions_tm'] = pos
      elif rc != _lib.NQ_ERR_UNSUPPORTED:
        _lib.check(rc)
    if rc == _lib.NQ_ERR_UNSUPPORTED:
      _lib.check(lib.nq_langevin_grad(C.byref(desc), _lib.ptr(r_dev), _lib.ptr(phi_dev), D,
                                      _lib.ptr(x0), _lib.ptr(seeds), n,
                                      _lib.ptr(out['work']), _lib.ptr(out['logp']), _lib.ptr(out['estimator']),
                                      _lib.ptr(out['x_final']), _lib.ptr(out['grad_sums']),
                                      _lib.ptr(out['moments']), _lib.ptr(ws), ws_bytes, st))
  return out


# ------------------------------------------------------------------------------ schedules
def MakeSchedule_chebyshev(timevec, coeffs, r0_init, r0_final):
  """Trap table r[len(timevec)]: end points fixed, interior = Chebyshev(coeffs) on [0, 1]."""
  vals, _ = _schedule_and_jacobian(timevec, coeffs, r0_init, r0_final)
  return vals


@functools.lru_cache(maxsize=16)
def _chebyshev_grid_basis(n_times: int, n_coeffs: int):
  """Monic Chebyshev basis on the reference's scaled step grid ((k-1)/(S-2), k = 1..S-1); it does
  not depend on the coefficients, so it is built once per (S, D)."""
  tv = np.arange(n_times)[1:-1]
  scaled = ((tv - 1).astype(np.float32) / np.float32(tv[-1] - 1)).astype(np.float32)
  basis = p10n.Chebyshev(np.zeros(n_coeffs, np.float32)).basis(scaled)
  basis.setflags(write=False)
  return scaled, basis


def _schedule_and_jacobian(timevec, coeffs, r0_init, r0_final):
  timevec = np.asarray(timevec)
  if isinstance(coeffs, p10n.Parameterization):
    tv = timevec[1:-1]
    scaled = ((tv - 1).astype(np.float32) / np.float32(tv[-1] - 1)).astype(np.float32)
    inner, jac = coeffs(scaled), coeffs.jacobian(scaled)
  else:
    if isinstance(coeffs, torch.Tensor):
      coeffs = coeffs.detach().cpu().numpy()
    w = np.asarray(coeffs, np.float32)
    contiguous = timevec.ndim == 1 and timevec[0] == 0 and timevec[-1] == len(timevec) - 1
    if contiguous and w.ndim == 1:
      _, jac = _chebyshev_grid_basis(len(timevec), w.shape[0])
      inner = np.zeros(jac.shape[0], np.float32)
      for j in range(w.shape[0]):        # same float32 summation order as Chebyshev.__call__
        inner = (inner + (w[j] * jac[:, j]).astype(np.float32)).astype(np.float32)
    else:
      tv = timevec[1:-1]
      scaled = ((tv - 1).astype(np.float32) / np.float32(tv[-1] - 1)).astype(np.float32)
      proto = p10n.Chebyshev(w)
      inner, jac = proto(scaled), proto.jacobian(scaled)
  vals = np.concatenate([[np.float32(r0_init)], inner, [np.float32(r0_final)]]).astype(np.float32)
  return vals, jac


_DEVICE_BASIS = {}


def _device_schedule_and_jacobian(S, coeffs, r0_init, r0_final, device):
  """Trap table and protocol Jacobian for a Chebyshev coefficient vector, built ON the device: the
  basis (= Jacobian) is uploaded once per (S, D, device) and the table is one `nq_schedule_affine`
  launch in the float32 summation order of `_schedule_and_jacobian`, so a gradient call does no host
  arithmetic and no table upload.  Returns None when `coeffs` is not a plain coefficient vector."""
  if isinstance(coeffs, p10n.Parameterization):
    return None
  if isinstance(coeffs, torch.Tensor):
    if coeffs.dim() != 1:
      return None
    w = coeffs.detach().to(device=device, dtype=torch.float32).contiguous()
  else:
    w_np = np.asarray(coeffs, np.float32)
    if w_np.ndim != 1:
      return None
    w = torch.from_numpy(np.array(w_np, dtype=np.float32, order='C')).to(device)
  D = int(w.shape[0])
  key = (int(S), D, str(device))
  basis = _DEVICE_BASIS.get(key)
  if basis is None:
    _, host = _chebyshev_grid_basis(S + 1, D)
    basis = torch.from_numpy(np.array(host, dtype=np.float32, order='C')).to(device)
    if len(_DEVICE_BASIS) >= 8:
      _DEVICE_BASIS.pop(next(iter(_DEVICE_BASIS)))
    _DEVICE_BASIS[key] = basis
  table = torch.empty(S + 1, dtype=torch.float32, device=device)
  _lib.check(_lib.lib().nq_schedule_affine(_lib.ptr(basis), _lib.ptr(w), None, float(np.float32(r0_init)),
                                           float(np.float32(r0_final)), S - 1, D, _lib.ptr(table), _lib.stream_arg()))
  return table, basis


def make_trap_fxn(timevec, coeffs, r0_init, r0_final):
  positions = MakeSchedule_chebyshev(timevec, coeffs, r0_init, r0_final)

  def Get_r0(step):
    return positions[step]
  return Get_r0


def theoretical_opt(t, tvec, lambda_t):
  return np.interp(np.asarray(t, np.float32), np.asarray(tvec, np.float32),
                   np.asarray(lambda_t, np.float32)).astype(np.float32)


def make_theoretical_trap_fxn(sim_stepvec, dt, tvec, lambda_t):
  positions = theoretical_opt(np.asarray(sim_stepvec) * dt, tvec, lambda_t)

  def Get_r0(step):
    return positions[step]
  return Get_r0


def fit_linear_to_cheby(end_time, dt, r0_init, r0_final, degree):
  """Initial protocol coefficients: numpy Chebyshev fit of the straight-line protocol on [0, 1]."""
  simulation_steps = int(end_time / dt)
  k = np.arange(1, simulation_steps)
  vals = (r0_final - r0_init) / simulation_steps * k + r0_init
  fit = np.polynomial.Chebyshev.fit((k - 1) / (simulation_steps - 2), vals, degree, domain=[0, 1])
  return p10n.Chebyshev(np.asarray(fit.coef, np.float32)).weights


# ------------------------------------------------------------------------------ integrator API
def brownian(energy_or_force, shift, dt, kT=0.1, gamma=0.1):
  """(init_fn, apply_fn) of the Euler-Maruyama Brownian integrator (barrier_crossing.py:28-94 = simulate.py:33-99).
  States may be batched: position `[..., N, dim]`, rng `[..., 2]` -- one key per state, as in the reference, whose
  draw for a state is `jax.random.normal(split, [1, N, dim])`.  N = dim = 1 (the hot path) runs `nq_langevin_apply`,
  any other particle count / dimension `nq_langevin_apply_nd`; `log_prob` is the sum over all elements (:90)."""
  energy = as_energy(energy_or_force)
  sh = as_shift(shift)

  def init_fn(key, R, mass=np.float32(1)):
    return BrownianState(R, mass, key, 0.)

  def apply_fn(state, t=np.float32(0), **kwargs):
    device = _lib.require_cuda()
    r0 = float(kwargs.get('r0', 0.0))
    pos = _dev_f32(state.position, device)
    if pos.dim() < 2:
      raise ValueError('state.position must have shape [..., N, dim]')
    desc = _desc(energy, sh, 1, 0, dt, kT, float(np.asarray(state.mass)), gamma)
    n_part, dim = int(pos.shape[-2]), int(pos.shape[-1])
    if n_part * dim != 1:      # rank-generic state (simulate.py:66-97): one key per [N, dim] state
      batch = tuple(pos.shape[:-2])
      n_states = int(np.prod(batch)) if batch else 1
      x = pos.reshape(n_states, n_part * dim).clone()
      rng = nqrandom.as_key_tensor(state.rng, device).reshape(-1, 2).clone()
      if rng.shape[0] != n_states:
        raise ValueError('state.position and state.rng disagree on the batch size')
      lp = torch.empty(n_states, dtype=torch.float32, device=device)
      _lib.check(_lib.lib().nq_langevin_apply_nd(C.byref(desc), r0, _lib.ptr(x), _lib.ptr(rng), _lib.ptr(lp), n_states,
                                                 n_part, dim, _lib.stream_arg()))
      return BrownianState(x.reshape(pos.shape), state.mass, rng.reshape(batch + (2,)), lp.reshape(batch))
    x = pos.reshape(-1).clone()
    rng = nqrandom.as_key_tensor(state.rng, device).reshape(-1, 2).clone()
    if rng.shape[0] != x.shape[0]:
      raise ValueError('state.position and state.rng disagree on the batch size')
    lp = torch.empty_like(x)
    _lib.check(_lib.lib().nq_langevin_apply(C.byref(desc), r0, _lib.ptr(x), _lib.ptr(rng), _lib.ptr(lp),
                                            x.shape[0], _lib.stream_arg()))
    batch = pos.shape[:-2]
    return BrownianState(x.reshape(pos.shape), state.mass, rng.reshape(tuple(batch) + (2,)),
                         lp.reshape(batch))

  return init_fn, apply_fn


def _batchify(init_position, key, device):
  batched = _is_batched_key(key)
  seeds = nqrandom.as_key_tensor(key, device).reshape(-1, 2)
  n = seeds.shape[0]
  x0 = _dev_f32(init_position, device).reshape(-1)
  if x0.numel() == 1 and n > 1:
    x0 = x0.expand(n).contiguous()
  if x0.numel() != n:
    raise ValueError(f'init_position holds {x0.numel()} particles for {n} keys (N=1, dim=1 only)')
  return batched, seeds, x0


def run_brownian_opt(energy_fn, coeffs, init_position, r0_init, r0_final, Neq, shift, key,
                     simulation_steps, dt=1e-5, kT=1e-5, mass=1.0, gamma=1.0):
  """One (key `[2]`) or a batch (keys `[B, 2]`, init_position `[B, 1, 1]`) of dragged-trap
  trajectories.  Returns (positions `[.., S+1, 1, 1]`, log_probs `[.., S]`, works `[.., S+1]`)."""
  device = _lib.require_cuda()
  S = int(simulation_steps)
  batched, seeds, x0 = _batchify(init_position, key, device)
  r, _ = _schedule_and_jacobian(np.arange(S + 1), coeffs, r0_init, r0_final)
  out = _simulate(energy_fn, shift, r, None, x0, seeds, S, Neq, dt, kT, mass, gamma,
                  positions_stride=1, step_outputs=True)
  positions = out['positions_tm'].t().unsqueeze(-1).unsqueeze(-1)   # [B, S+1, 1, 1] view
  log_probs = out['step_logps_tm'].t()
  works = out['step_works_tm'].t()
  if not batched:
    return positions[0], log_probs[0], works[0]
  return positions, log_probs, works


def run_brownian_stationary_trap(energy_fn, init_position, r0, shift, key, simulation_steps, dt, kT, mgamma):
  """Positions `[.., S, 1, 1]` of the integrator in a stationary trap at r0 (mass = 1)."""
  device = _lib.require_cuda()
  S = int(simulation_steps)
  batched, seeds, x0 = _batchify(init_position, key, device)
  n = x0.shape[0]
  desc = _desc(energy_fn, shift, S, 0, dt, kT, 1.0, mgamma)
  pos = torch.empty((S, n), dtype=torch.float32, device=device)
  _lib.check(_lib.lib().nq_langevin_stationary(C.byref(desc), float(r0), _lib.ptr(x0), _lib.ptr(seeds), n,
                                               None, _lib.ptr(pos), _lib.stream_arg()))
  positions = pos.t().unsqueeze(-1).unsqueeze(-1)
  return positions if batched else positions[0]


def _stationary_final(energy_fn, init_position, r0, shift, seeds, S, dt, kT, gamma):
  device = seeds.device
  n = seeds.shape[0]
  x0 = _dev_f32(init_position, device).reshape(-1)
  if x0.numel() == 1:
    x0 = x0.expand(n).contiguous()
  desc = _desc(energy_fn, shift, S, 0, dt, kT, 1.0, gamma)
  xf = torch.empty(n, dtype=torch.float32, device=device)
  _lib.check(_lib.lib().nq_langevin_stationary(C.byref(desc), float(r0), _lib.ptr(x0), _lib.ptr(seeds), n,
                                               _lib.ptr(xf), None, _lib.stream_arg()))
  return xf.reshape(n, 1, 1)


def single_brownian_eqm(energy_fn, init_position, r0_init, shift_fn, sim_length, dt=1e-6, kT=1e-5,
                        mass=1.0, gamma=1.0):
  def _single_brownian_eqm(seed):
    device = _lib.require_cuda()
    batched = _is_batched_key(seed)
    seeds = nqrandom.as_key_tensor(seed, device).reshape(-1, 2)
    xf = _stationary_final(energy_fn, init_position, r0_init, shift_fn, seeds, int(sim_length), dt, kT, gamma)
    return xf if batched else xf[0]
  return _single_brownian_eqm


def mapped_brownian_eqm(batch_size, energy_fn, init_position, r0_init, shift_fn, sim_length, dt=1e-6,
                        kT=1e-5, mass=1.0, gamma=1.0, *, reference_quirk=True):
  """Boltzmann-initial positions `[batch_size, 1, 1]`.

  The reference ignores its dt/kT/mass/gamma arguments here and hard-codes 1e-6, 1e-5, 1, 1
  (barrier_crossing.py:250); `reference_quirk=True` (default, drop-in) reproduces that,
  `reference_quirk=False` uses the arguments as passed."""
  if reference_quirk:
    dt, kT, mass, gamma = 1e-6, 1e-5, 1.0, 1.0
  eqm = single_brownian_eqm(energy_fn, init_position, r0_init, shift_fn, sim_length, dt, kT, mass, gamma)

  def _mapped_brownian_eqm(seed):
    return eqm(nqrandom.split(seed, batch_size))
  return _mapped_brownian_eqm


# ------------------------------------------------------------------------------ estimators
_TERMS = {'reinforce': (1.0, 1.0), 'logP': (1.0, 0.0), 'reparam': (0.0, 1.0)}


def gradient_terms(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift, simulation_steps,
                   dt, temperature, mass, gamma, positions_stride=None):
  """One fused pass over the trajectories keyed by `seeds [n, 2]`: un-normalised float64 sums of
  the log-prob (score) and reparameterisation gradient terms plus work moments.  This is the
  rank-local piece that `distributed.allreduce_gradient` combines."""
  device = _lib.require_cuda()
  S = int(simulation_steps)
  seeds = nqrandom.as_key_tensor(seeds, device).reshape(-1, 2)
  x0 = _dev_f32(init_pos, device).reshape(-1)
  if x0.numel() == 1:
    x0 = x0.expand(seeds.shape[0]).contiguous()
  on_device = _device_schedule_and_jacobian(S, coeffs, r0_init, r0_final, device) if S >= 3 else None
  r, phi = on_device if on_device is not None else _schedule_and_jacobian(np.arange(S + 1), coeffs, r0_init, r0_final)
  return _simulate(energy_fn, shift, r, phi, x0, seeds, S, Neq, dt, temperature, mass, gamma,
                   positions_stride=positions_stride)


def _estimator_factory(term):
  w_logp, w_rep = _TERMS[term]

  def single(energy_fn, r0_init, r0_final, Neq, shift, simulation_steps, dt, temperature, mass, gamma,
             record_positions=None):
    def _single_estimate(coeffs, init_pos, seed):
      """value_and_grad(has_aux): ((objective, (positions, tot_log_prob, total_work)), grad)."""
      batched = _is_batched_key(seed)
      out = gradient_terms(energy_fn, coeffs, init_pos, seed, r0_init, r0_final, Neq, shift,
                           simulation_steps, dt, temperature, mass, gamma)
      n = out['work'].shape[0]
      if n != 1 and not batched:
        raise ValueError('single estimate called with several keys')
      if batched:
        raise NotImplementedError('use estimate_gradient_* for batches (per-trajectory gradients '
                                  'are reduced on the device)')
      grad = (w_logp * out['grad_sums'][0] + w_rep * out['grad_sums'][1]).to(torch.float32)
      positions = _maybe_positions(energy_fn, coeffs, init_pos, seed, r0_init, r0_final, Neq, shift,
                                   simulation_steps, dt, temperature, mass, gamma, record_positions, 1)
      W, lp = out['work'][0], out['logp'][0]
      value = w_logp * lp * W + w_rep * W
      return (value, (None if positions is None else positions[0], lp, W)), grad
    return _single_estimate

  def batch(batch_size, energy_fn, r0_init, r0_final, Neq, shift, simulation_steps, dt, temperature, mass,
            gamma, record_positions=None):
    def _estimate_gradient(coeffs, init_pos, seed):
      """(mean grad [D], (objective [B], (positions [B,S+1,1,1] | None, tot_log_prob [B], total_work [B])))."""
      seeds = nqrandom.split(seed, batch_size)
      want_pos = _wants_positions(record_positions, batch_size, simulation_steps)
      out = gradient_terms(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift,
                           simulation_steps, dt, temperature, mass, gamma, positions_stride=1 if want_pos else None)
      grad = ((w_logp * out['grad_sums'][0] + w_rep * out['grad_sums'][1]) / batch_size).to(torch.float32)
      if out.get('positions_tm') is not None:     # recorded by the same pass (small ensembles)
        positions = out['positions_tm'].t().unsqueeze(-1).unsqueeze(-1)
      else:
        positions = _maybe_positions(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift,
                                     simulation_steps, dt, temperature, mass, gamma, want_pos, batch_size)
      W, lp = out['work'], out['logp']
      if term == 'reinforce':
        value = out['estimator']
      else:
        value = w_logp * lp * W + w_rep * W
      _estimate_gradient.last = out
      return grad, (value, (positions, lp, W))
    return _estimate_gradient

  return single, batch


def _wants_positions(record_positions, n, S):
  if record_positions is None:
    return 4 * n * (int(S) + 1) <= MAX_AUTO_POSITION_BYTES
  return bool(record_positions)


def _maybe_positions(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift, S, dt, kT, mass,
                     gamma, record_positions, n):
  S = int(S)
  if not _wants_positions(record_positions, n, S):
    return None
  device = _lib.require_cuda()
  seeds = nqrandom.as_key_tensor(seeds, device).reshape(-1, 2)
  x0 = _dev_f32(init_pos, device).reshape(-1)
  if x0.numel() == 1:
    x0 = x0.expand(seeds.shape[0]).contiguous()
  r, _ = _schedule_and_jacobian(np.arange(S + 1), coeffs, r0_init, r0_final)
  rec = _simulate(energy_fn, shift, r, None, x0, seeds, S, Neq, dt, kT, mass, gamma, positions_stride=1)
  return rec['positions_tm'].t().unsqueeze(-1).unsqueeze(-1)


single_estimate_boltzinit, estimate_gradient_boltzinit = _estimator_factory('reinforce')
single_estimate_boltzinit_logP, estimate_gradient_boltzinit_logP = _estimator_factory('logP')
single_estimate_boltzinit_reparam, estimate_gradient_boltzinit_reparam = _estimator_factory('reparam')
(single_estimate_boltzinit_REINFORCE_recordall,
 estimate_gradient_boltzinit_REINFORCE_recordall) = _estimator_factory('reinforce')


# ------------------------------------------------------------------------------ checkpoint + replay
def forward_checkpointed(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift, simulation_steps,
                         dt, temperature, mass, gamma, checkpoint_every):
  """Forward pass that stores one (x, rng) carry per block of `checkpoint_every` steps -- the memory
  profile of `control_flow.checkpoint_scan` (control_flow.py:16-31) for this integrator.  Returns the
  trajectory totals plus an opaque `tape` for `replay_vjp`."""
  control_flow.check_divides(int(simulation_steps), int(checkpoint_every))
  device = _lib.require_cuda()
  S, K = int(simulation_steps), int(checkpoint_every)
  seeds = nqrandom.as_key_tensor(seeds, device).reshape(-1, 2)
  n = seeds.shape[0]
  x0 = _dev_f32(init_pos, device).reshape(-1)
  if x0.numel() == 1:
    x0 = x0.expand(n).contiguous()
  r, phi = _schedule_and_jacobian(np.arange(S + 1), coeffs, r0_init, r0_final)
  lib = _lib.lib()
  desc = _desc(energy_fn, shift, S, Neq, dt, temperature, mass, gamma)
  r_dev = _dev_f32(r, device)
  n_ckpt = lib.nq_langevin_ckpt_count(S, K)
  ckpt_x = torch.empty((n_ckpt, n), dtype=torch.float32, device=device)
  ckpt_rng = torch.empty((n_ckpt, 4, n), dtype=torch.int32, device=device)
  ws_bytes = lib.nq_langevin_workspace_bytes(n, 0)
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
  out = dict(work=torch.empty(n, dtype=torch.float32, device=device),
             logp=torch.empty(n, dtype=torch.float32, device=device),
             x_final=torch.empty(n, dtype=torch.float32, device=device),
             moments=torch.zeros(_lib.NQ_MOM_COUNT, dtype=torch.float64, device=device))
  _lib.check(lib.nq_langevin_forward_ckpt(C.byref(desc), _lib.ptr(r_dev), _lib.ptr(x0), _lib.ptr(seeds), n, K,
                                          _lib.ptr(out['work']), _lib.ptr(out['logp']), _lib.ptr(out['x_final']),
                                          _lib.ptr(ckpt_x), _lib.ptr(ckpt_rng), _lib.ptr(out['moments']),
                                          _lib.ptr(ws), ws_bytes, _lib.stream_arg()))
  out['tape'] = dict(desc=desc, r_dev=r_dev, phi=phi, ckpt_x=ckpt_x, ckpt_rng=ckpt_rng, K=K, S=S, n=n)
  return out


def replay_vjp(tape, lbar, wbar):
  """Reverse ("adjoint") pass: regenerate every block from its checkpoint and return
  gbar[k] = sum_n lbar_n dlogp_n/dr_k + wbar_n dW_n/dr_k for k = 0..S (float64 CUDA tensor)."""
  device = tape['r_dev'].device
  lbar = _dev_f32(lbar, device).reshape(-1)
  wbar = _dev_f32(wbar, device).reshape(-1)
  gbar = torch.empty(tape['S'] + 1, dtype=torch.float64, device=device)
  _lib.check(_lib.lib().nq_langevin_replay(C.byref(tape['desc']), _lib.ptr(tape['r_dev']), tape['n'], tape['K'],
                                           _lib.ptr(tape['ckpt_x']), _lib.ptr(tape['ckpt_rng']), _lib.ptr(lbar),
                                           _lib.ptr(wbar), _lib.ptr(gbar), None, 0, _lib.stream_arg()))
  return gbar


def estimate_gradient_boltzinit_checkpointed(batch_size, energy_fn, r0_init, r0_final, Neq, shift, simulation_steps,
                                             dt, temperature, mass, gamma, checkpoint_every, term='reinforce'):
  """estimate_gradient_boltzinit (:216-224) as forward-with-checkpoints + replay: the two-pass
  reverse-mode structure of jax.grad through a checkpointed scan.  Same result as the fused
  single-pass estimator; useful for arbitrary per-trajectory cotangents via `replay_vjp`."""
  w_logp, w_rep = _TERMS[term]

  def _estimate_gradient(coeffs, init_pos, seed):
    seeds = nqrandom.split(seed, batch_size)
    fwd = forward_checkpointed(energy_fn, coeffs, init_pos, seeds, r0_init, r0_final, Neq, shift, simulation_steps,
                               dt, temperature, mass, gamma, checkpoint_every)
    # objective = logp * stop_grad(W) + W :  d/dlogp = W,  d/dW = 1
    gbar = replay_vjp(fwd['tape'], w_logp * fwd['work'], torch.full_like(fwd['work'], w_rep))
    phi = torch.from_numpy(np.asarray(fwd['tape']['phi'], np.float64)).to(gbar.device)
    grad = ((gbar[1:-1] @ phi) / batch_size).to(torch.float32)
    W, lp = fwd['work'], fwd['logp']
    return grad, (w_logp * lp * W + w_rep * W, (None, lp, W)), gbar
  return _estimate_gradient
