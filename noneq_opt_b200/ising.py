"""Running and optimising Ising simulations under a time-varying field and temperature --
B200 drop-in for `noneq_opt/ising.py`.

Same names, argument order and return nesting as the reference:
  IsingParameters / IsingSchedule / IsingState / IsingSummary :14-40
  log_temp_baseline :45   field_baseline :52   seed_stream :57   map_slice :63   map_stack :67
  random_spins :71   sum_neighbors :75   energy :85   log_prob :89   flip_logits :93
  even_odd_masks :99   masked_update :109 (as part of update)   update :129   simulate_ising :153
  estimate_gradient_rev / _fwd :179-236, total_entropy_production :239, total_work :252
  get_train_step :263 and the _REINFORCE / _logP / _reparam / _track_terms families :294-630

The sweeps run in csrc/ising.cu (bit-packed lattice resident in shared memory, jax-matched
threefry uniforms, integer acceptance thresholds); this module builds the per-sweep class
tables from the schedule, launches the kernel and assembles losses and gradients from the
kernel's per-sweep sufficient statistics (closed forms of SURVEY Appendix B.2 -- flips are
integer samples, so spins carry no derivative and no tape is needed).

Differences from the reference: lattices have rank 1-3 (square 2-D lattices with
`L % 64 == 0` take the tuned kernel); a batch of seeds `[R, 2]` may be passed wherever the reference
uses `jax.vmap`; loss functions must be `total_work` or
`total_entropy_production` (the two the reference ships); arrays are CUDA torch tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from . import control_flow
from . import random as nqrandom
from . import tree

F32 = np.float32


class IsingParameters(NamedTuple):
  log_temp: object
  field: object


class IsingSchedule(NamedTuple):
  log_temp: Callable
  field: Callable

  def __call__(self, x):
    return IsingParameters(log_temp=self.log_temp(x), field=self.field(x))


class IsingState(NamedTuple):
  spins: object
  params: IsingParameters


class IsingSummary(NamedTuple):
  work: object
  dissipated_heat: object
  forward_log_prob: object
  reverse_log_prob: object
  entropy_production: object
  magnetization: object
  energy: object


# ------------------------------------------------------------------------------ small helpers
def log_temp_baseline(min_temp=.69, max_temp=10., degree=1):
  def _log_temp_baseline(t):
    t = np.asarray(t, F32)
    scale = F32(max_temp - min_temp)
    shape = ((F32(1) - t) ** degree * t ** degree * F32(4 ** degree)).astype(F32)
    return np.log((shape * scale + F32(min_temp)).astype(np.float64)).astype(F32)
  return _log_temp_baseline


def field_baseline(start_field=-1., end_field=1.):
  def _field_baseline(t):
    t = np.asarray(t, F32)
    return ((F32(1) - t) * F32(start_field) + t * F32(end_field)).astype(F32)
  return _field_baseline


def seed_stream(seed):
  key = nqrandom.PRNGKey(seed)
  while True:
    key, _ = nqrandom.split_host(key)
    yield key


def map_slice(x, idx):
  return tree.tree_map(lambda y: y[idx], x)


def map_stack(xs, axis=0):
  def stack(*args):
    if isinstance(args[0], torch.Tensor):
      return torch.stack(args, axis)
    return np.stack([np.asarray(a) for a in args], axis)
  return tree.tree_map(stack, *xs)


def _as_cuda(x, dtype=None):
  device = _lib.require_cuda()
  if isinstance(x, torch.Tensor):
    t = x.to(device)
  else:
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(x))).to(device)
  return t if dtype is None else t.to(dtype)


def _pack(spins, lattice_ndim: int = 2) -> torch.Tensor:
  """+-1 spins `[..., *lattice]` (int or float) -> bit-packed int32 words `[..., ceil(sites/32)]`."""
  t = _as_cuda(spins)
  lead, n = t.shape[:t.dim() - lattice_ndim], int(np.prod(t.shape[t.dim() - lattice_ndim:]))
  words = (n + 31) // 32
  is_float = t.is_floating_point()
  flat = t.to(torch.float32 if is_float else torch.int32).reshape(-1, n).contiguous()
  out = torch.empty((flat.shape[0], words), dtype=torch.int32, device=t.device)
  lib = _lib.lib()
  if n % 32 == 0:   # lattices are word aligned: one launch for the whole batch
    _lib.check(lib.nq_ising_pack(_lib.ptr(flat), int(is_float), flat.numel(), _lib.ptr(out), _lib.stream_arg()))
  else:
    for r in range(flat.shape[0]):
      _lib.check(lib.nq_ising_pack(_lib.ptr(flat[r]), int(is_float), n, _lib.ptr(out[r]), _lib.stream_arg()))
  return out.reshape(tuple(lead) + (words,))


def _shape_of(L):
  return (int(L), int(L)) if np.isscalar(L) else tuple(int(v) for v in L)


def _unpack(bits: torch.Tensor, L, like) -> torch.Tensor:
  """bit-packed `[..., words]` -> +-1 spins `[..., *lattice]` with the dtype family of `like`
  (`L`: side of a square 2-D lattice, or the lattice shape)."""
  shape = _shape_of(L)
  is_float = bool(like.is_floating_point()) if isinstance(like, torch.Tensor) else \
      np.issubdtype(np.asarray(like).dtype, np.floating)
  n = int(np.prod(shape))
  flat = bits.reshape(-1, bits.shape[-1]).contiguous()
  out = torch.empty((flat.shape[0], n), dtype=torch.float32 if is_float else torch.int32, device=bits.device)
  lib = _lib.lib()
  if n % 32 == 0:
    _lib.check(lib.nq_ising_unpack(_lib.ptr(flat), out.numel(), int(is_float), _lib.ptr(out), _lib.stream_arg()))
  else:
    for r in range(flat.shape[0]):
      _lib.check(lib.nq_ising_unpack(_lib.ptr(flat[r]), n, int(is_float), _lib.ptr(out[r]), _lib.stream_arg()))
  return out.reshape(tuple(bits.shape[:-1]) + shape)


def random_spins(shape, p, seed):
  """2 * Bernoulli(p).sample(shape) - 1 with jax's uniform stream; int32 CUDA tensor (1-3 dims)."""
  shape = (int(shape),) if np.isscalar(shape) else tuple(int(v) for v in shape)
  if not 1 <= len(shape) <= 3:
    raise NotImplementedError('lattices must have 1, 2 or 3 dimensions')
  device = _lib.require_cuda()
  n = int(np.prod(shape))
  bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=device)
  _lib.check(_lib.lib().nq_ising_random_spins(_lib.key_arg(nqrandom._host_key(seed)), n, float(p),
                                              _lib.ptr(bits), _lib.stream_arg()))
  return _unpack(bits, shape, torch.zeros((), dtype=torch.int32))


def sum_neighbors(spins):
  """Sum over the neighbours of each point in a periodic grid (any rank); torch on the GPU."""
  t = _as_cuda(spins)
  out = torch.zeros_like(t)
  for axis in range(t.dim()):
    for shift in (-1, 1):
      out = out + torch.roll(t, shift, axis)
  return out


def _param(x):
  """A schedule value as a float32 CUDA scalar; torch tensors pass through (keeps autograd / vmap)."""
  device = _lib.require_cuda()
  if isinstance(x, torch.Tensor):
    return x.to(device=device, dtype=torch.float32)
  return torch.tensor(float(np.asarray(x)), dtype=torch.float32, device=device)


def energy(state: IsingState):
  s = _as_cuda(state.spins).to(torch.float32)
  return (-s * (sum_neighbors(s) / 2 + _param(state.params.field))).sum()


def log_prob(state: IsingState):
  return -energy(state) / torch.exp(_param(state.params.log_temp))


def flip_logits(spins, params: IsingParameters):
  """Log odds of a flip at each site under Glauber dynamics."""
  s = _as_cuda(spins).to(torch.float32)
  return -2 * s * (sum_neighbors(s) + _param(params.field)) * torch.exp(-_param(params.log_temp))


def even_odd_masks(shape):
  for s in shape:
    if s % 2:
      raise ValueError(f'All entries in `shape` must be even; got {shape}.')
  grid = np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')
  mask_even = sum(grid) % 2
  return mask_even, 1 - mask_even


def _to_host(x):
  return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


# ------------------------------------------------------------------------------ class tables
def _exp32(x):
  return np.exp(np.asarray(x, F32).astype(np.float64)).astype(F32)


def _log1p32(x):
  return np.log1p(np.asarray(x, F32).astype(np.float64)).astype(F32)


def _log_sigmoid32(x):
  a = (-np.asarray(x, F32)).astype(F32)
  return (-(np.maximum(a, F32(0)) + _log1p32(_exp32(-np.abs(a))))).astype(F32)


def class_tables(log_temp, field, ndim: int = 2):
  """Per-sweep tables for the 2 x (2*ndim+1) (spin, #up-neighbours) site classes (2x5 in 2-D).

  Class slot = 8*(spin up) + n_up.  For sweep t = 1..T-1 (parameters[t]):
    logits  : ((-2 s) (nsum + h_t)) exp(-log_temp_t) in float32, the reference's operation order
              (ising.py:96); transcendental = binary64 evaluation rounded once to float32
    thr     : ceil(sigmoid32(logit) * 2^23) -- `uniform < probs` of distrax.Bernoulli as an
              integer compare on the 23 mantissa bits jax keeps of each threefry word
    lut     : [logit, log_sigmoid32(-logit), log_sigmoid32(logit), sigmoid(logit) (binary64)]
  Returns (thr uint32 [T-1,16], lut float64 [T-1,4,16])."""
  lt = np.asarray(log_temp, F32)[1:]
  h = np.asarray(field, F32)[1:]
  n_sweeps = lt.shape[0]
  thr = np.zeros((n_sweeps, 16), np.uint32)
  lut = np.zeros((n_sweeps, 4, 16), np.float64)
  einv = _exp32(-lt)
  # all 2 x (2 ndim + 1) classes at once (the same elementwise float32 operations as a loop over the classes; one
  # pass of numpy calls instead of ten: this runs on the host in front of every estimator call)
  ncls = 2 * ndim + 1
  slots = np.array([8 * up + n_up for up in (0, 1) for n_up in range(ncls)])
  s = np.array([1 if up else -1 for up in (0, 1) for _ in range(ncls)], F32)
  nsum = np.array([2 * n_up - 2 * ndim for _ in (0, 1) for n_up in range(ncls)], F32)
  logit = (((F32(-2) * s)[None, :] * (nsum[None, :] + h[:, None]).astype(F32)).astype(F32) * einv[:, None]).astype(F32)
  prob = (F32(1) / (F32(1) + _exp32(-logit))).astype(F32)
  thr[:, slots] = np.ceil(prob.astype(np.float64) * 8388608.0).astype(np.uint32)
  lut[:, 0, slots] = logit
  lut[:, 1, slots] = _log_sigmoid32(-logit)
  lut[:, 2, slots] = _log_sigmoid32(logit)
  lut[:, 3, slots] = 1.0 / (1.0 + np.exp(-logit.astype(np.float64)))
  return thr, lut


# ------------------------------------------------------------------------------ simulation
class _Run(NamedTuple):
  final_bits: torch.Tensor   # [R, words]
  summary: torch.Tensor      # [7, R, T-1] float32
  partials: Optional[torch.Tensor]  # [4, R, T-1]
  counts: Optional[torch.Tensor]
  states: Optional[torch.Tensor]
  MB0: torch.Tensor          # [R or 1, 2] int64 spin sum and bond sum of the initial lattice


def _run_kernel(log_temp, field, spins0_bits, broadcast, seeds, L, want_partials=False, want_counts=False,
                want_states=False, direct_keys=False) -> _Run:
  device = seeds.device
  lib = _lib.lib()
  R, T = int(seeds.shape[0]), int(np.asarray(field).shape[0])
  shape = _shape_of(L)
  ndim = len(shape)
  if any(v % 2 for v in shape):
    raise ValueError(f'All entries in `shape` must be even; got {shape}.')
  thr, lut = class_tables(log_temp, field, ndim)
  thr_d = torch.from_numpy(thr.view(np.int32)).to(device)
  lut_d = torch.from_numpy(lut).to(device)
  field_d = torch.from_numpy(np.ascontiguousarray(np.asarray(field, F32))).to(device)
  lt_d = torch.from_numpy(np.ascontiguousarray(np.asarray(log_temp, F32))).to(device)
  words = (int(np.prod(shape)) + 31) // 32
  dims = (C.c_int32 * 3)(*(list(shape) + [1] * (3 - ndim)))
  square = ndim == 2 and shape[0] == shape[1]
  desc = _lib.NqIsingDesc(L=shape[0], T=T, R=R, spins0_broadcast=int(broadcast), sweep_key_mode=int(direct_keys),
                          ndim=0 if square else ndim, dims=dims)
  final_bits = torch.empty((R, words), dtype=torch.int32, device=device)
  summary = torch.zeros((_lib.NQ_IS_NSUMMARY, R, max(T - 1, 0)), dtype=torch.float32, device=device)
  partials = torch.zeros((_lib.NQ_IS_NPARTIAL, R, T - 1), dtype=torch.float32, device=device) if want_partials else None
  counts = torch.zeros((R, T - 1, 2, 2, 16), dtype=torch.int32, device=device) if want_counts else None
  states = torch.empty((R, T - 1, words), dtype=torch.int32, device=device) if want_states else None
  n0 = 1 if broadcast else R
  MB0 = torch.empty((n0, 2), dtype=torch.int64, device=device)
  _lib.check(lib.nq_ising_measure(_lib.ptr(spins0_bits), ndim, dims, n0, _lib.ptr(MB0), _lib.stream_arg()))
  ws_bytes = lib.nq_ising_workspace_bytes(C.byref(desc))
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device) if ws_bytes else None
  if T > 1:
    _lib.check(lib.nq_ising_run(C.byref(desc), _lib.ptr(thr_d), _lib.ptr(lut_d), _lib.ptr(field_d), _lib.ptr(lt_d),
                                _lib.ptr(spins0_bits), _lib.ptr(seeds), _lib.ptr(final_bits), _lib.ptr(summary),
                                _lib.ptr(partials), _lib.ptr(counts), _lib.ptr(states), _lib.ptr(ws), ws_bytes,
                                _lib.stream_arg()))
  else:
    final_bits.copy_(spins0_bits.expand(R, words))
  return _Run(final_bits, summary, partials, counts, states, MB0)


def _prepare(parameters: IsingParameters, initial_spins, seed):
  device = _lib.require_cuda()
  log_temp = np.asarray(_to_host(parameters.log_temp), F32).reshape(-1)
  field = np.asarray(_to_host(parameters.field), F32).reshape(-1)
  if log_temp.shape != field.shape:
    raise ValueError('log_temp and field schedules must have the same length')
  spins = _as_cuda(initial_spins)
  batched = np.ndim(seed) == 2 if not isinstance(seed, torch.Tensor) else seed.dim() == 2
  seeds = nqrandom.as_key_tensor(seed, device).reshape(-1, 2)
  if not 1 <= spins.dim() <= 3:
    raise NotImplementedError('initial_spins must be ONE periodic lattice of 1, 2 or 3 dimensions (as in the '
                              'reference, replicas differ by their seeds only)')
  shape = tuple(int(v) for v in spins.shape)
  if any(v % 2 for v in shape):
    raise ValueError(f'All entries in `shape` must be even; got {shape}.')
  L = shape[0] if len(shape) == 2 and shape[0] == shape[1] else shape
  broadcast = True
  bits = _pack(spins, len(shape)).reshape(1, -1)
  return log_temp, field, spins, L, batched, seeds, broadcast, bits


def _summary_from(run: _Run, batched: bool) -> IsingSummary:
  s = run.summary
  return IsingSummary(*[(s[f] if batched else s[f, 0]) for f in range(_lib.NQ_IS_NSUMMARY)])


def masked_update(state: IsingState, new_params: IsingParameters, mask, seed):
  """Random update of the sites selected by an arbitrary `mask` (ising.py:109-126), returning
  (new_state, forward_log_prob, reverse_log_prob).  This is the reference's building block of
  `update`; with a caller-supplied mask it cannot use the checkerboard kernel, so it is evaluated
  with torch ops on the GPU from the same jax-matched uniforms (`random.uniform`).  The hot path is
  `update` / `simulate_ising`."""
  spins = _as_cuda(state.spins)
  s = spins.to(torch.float32)
  m = _as_cuda(mask).to(torch.float32)
  logits = flip_logits(s, new_params)
  u = nqrandom.uniform(seed, tuple(s.shape))
  flips = (u < torch.sigmoid(logits)).to(torch.float32) * m
  log_sig = torch.nn.functional.logsigmoid

  def bern_log_prob(lg):
    return (1 - flips) * log_sig(-lg) + flips * log_sig(lg)

  forward_log_prob = (bern_log_prob(logits) * m).sum()
  new_spins = s * (1 - 2 * flips)
  reverse_log_prob = (bern_log_prob(flip_logits(new_spins, new_params)) * m).sum()
  return IsingState(new_spins.to(spins.dtype), new_params), forward_log_prob, reverse_log_prob


def update(state: IsingState, new_params: IsingParameters, seed):
  """One sweep: both checkerboard half-sweeps with (seed_even, seed_odd) = split(seed, 2)."""
  device = _lib.require_cuda()
  old_lt = float(np.asarray(_to_host(state.params.log_temp)))
  old_h = float(np.asarray(_to_host(state.params.field)))
  new_lt = float(np.asarray(_to_host(new_params.log_temp)))
  new_h = float(np.asarray(_to_host(new_params.field)))
  params = IsingParameters(np.array([old_lt, new_lt], F32), np.array([old_h, new_h], F32))
  log_temp, field, spins, L, batched, seeds, broadcast, bits = _prepare(params, state.spins, seed)
  # update() consumes `seed` itself as the sweep key (ising.py:132), whereas simulate_ising
  # derives sweep keys with split(seed, T-1) (:161): sweep_key_mode=1 tells the kernel so.
  run = _run_kernel(log_temp, field, bits, broadcast, seeds, L, direct_keys=True)
  new_spins = _unpack(run.final_bits, L, spins)
  summary = IsingSummary(*[(run.summary[f, :, 0] if batched else run.summary[f, 0, 0])
                           for f in range(_lib.NQ_IS_NSUMMARY)])
  return IsingState(new_spins if batched else new_spins[0], new_params), summary


def simulate_ising(parameters: IsingParameters, initial_spins, seed, checkpoint_every: Optional[int] = None,
                   return_states: bool = False):
  """T-1 sweeps under `parameters` (arrays of length T).  Returns (final_state, summary) or
  (final_state, (summary, states)) like `lax.scan` over `update` in the reference."""
  log_temp, field, spins, L, batched, seeds, broadcast, bits = _prepare(parameters, initial_spins, seed)
  T = field.shape[0]
  if checkpoint_every is not None:
    control_flow.check_divides(T - 1, checkpoint_every)
  run = _run_kernel(log_temp, field, bits, broadcast, seeds, L, want_states=return_states)
  final_spins = _unpack(run.final_bits, L, spins)
  final_state = IsingState(final_spins if batched else final_spins[0],
                           IsingParameters(log_temp[-1], field[-1]))
  summary = _summary_from(run, batched)
  if not return_states:
    return final_state, summary
  st = _unpack(run.states, L, spins)
  states = IsingState(st if batched else st[0], IsingParameters(log_temp[1:], field[1:]))
  return final_state, (summary, states)


# ------------------------------------------------------------------------------ losses
# A `LossFn` maps (initial state, final_state, trajectory summary) to a scalar loss.
LossFn = Callable[[IsingState, IsingState, IsingSummary], object]


def _sum32(x, dim=-1):
  """`.sum()` of a float32 array as the exact sum of its terms rounded once (XLA's reduction
  order is unspecified; see oracle/__init__.py)."""
  return x.to(torch.float64).sum(dim).to(torch.float32)


def total_entropy_production(initial_state: IsingState, final_state: IsingState, summary: IsingSummary):
  trajectory = _sum32(_as_cuda(summary.entropy_production))
  endpoint = log_prob(initial_state) - log_prob(final_state)
  return endpoint + trajectory


def total_work(initial_state: IsingState, final_state: IsingState, summary: IsingSummary):
  return _sum32(_as_cuda(summary.work))


_LOSS_KINDS = {total_work: 'work', total_entropy_production: 'entropy'}


def _loss_kind(loss_function):
  """'work' / 'entropy' for the two losses the reference ships (closed-form fast path), 'generic' for any
  other callable `loss(initial_state, final_state, summary)` written with torch ops."""
  if not callable(loss_function):
    raise TypeError('loss_function must be callable')
  try:
    return _LOSS_KINDS.get(loss_function, 'generic')
  except TypeError:
    return 'generic'


def _generic_loss_sensitivities(loss_function, run: _Run, log_temp, field, L, initial_spins, final_spins):
  """Arbitrary LossFn (ising.py:177): the loss is evaluated per replica with torch (vmap) on the
  summary and the two end-point states; torch autograd gives d loss / d summary and
  d loss / d end-point parameters, and the chain rule to the schedule tables uses the same closed-form
  per-sweep sensitivities as the built-in losses:
     work_t = -(h_t - h_{t-1}) M_{t-1}          heat_t = dB_t + h_t dM_t          energy_t = -(B_t + h_t M_t)
     d fwd_lp_t, d rev_lp_t / d(h_t, lt_t) from the kernel's partials; entropy_production = fwd - rev."""
  n_sites = float(np.prod(_shape_of(L)))
  S32 = run.summary                                              # [7, R, T-1]
  R, Tm1 = S32.shape[1], S32.shape[2]
  T = Tm1 + 1
  dev = S32.device
  lt = torch.from_numpy(np.asarray(log_temp, F32)).to(dev)
  h = torch.from_numpy(np.asarray(field, F32)).to(dev)
  spins0 = _as_cuda(initial_spins).to(torch.float32)

  def per_replica(summ, fin, lt0, h0, ltf, hf):
    initial_state = IsingState(spins0, IsingParameters(lt0, h0))
    final_state = IsingState(fin, IsingParameters(ltf, hf))
    out = loss_function(initial_state, final_state, IsingSummary(*summ.unbind(0)))
    return out, out

  grad_fn = torch.func.vmap(torch.func.grad(per_replica, argnums=(0, 2, 3, 4, 5), has_aux=True),
                            in_dims=(1, 0, None, None, None, None))
  (g_summ, g_lt0, g_h0, g_ltf, g_hf), loss = grad_fn(S32, final_spins.to(torch.float32), lt[0], h[0], lt[-1], h[-1])
  g = g_summ.to(torch.float64)                                   # [R, 7, T-1]
  P = run.partials.to(torch.float64)
  M0 = run.MB0[:, 0].to(torch.float64).expand(R)
  M_after = torch.round(S32[_lib.NQ_IS_MAG].to(torch.float64) * n_sites)
  M_before = torch.cat([M0[:, None], M_after[:, :-1]], 1)
  dL_dh = torch.zeros((R, T), dtype=torch.float64, device=dev)
  dL_dl = torch.zeros((R, T), dtype=torch.float64, device=dev)
  gW, gQ, gF, gR, gE, gEn = (g[:, i] for i in (_lib.NQ_IS_WORK, _lib.NQ_IS_HEAT, _lib.NQ_IS_FWD, _lib.NQ_IS_REV,
                                              _lib.NQ_IS_EP, _lib.NQ_IS_ENERGY))
  dL_dh[:, 1:] += -gW * M_before + gQ * (M_after - M_before) - gEn * M_after
  dL_dh[:, :-1] += gW * M_before
  dL_dh[:, 1:] += (gF + gE) * P[_lib.NQ_IS_DFWD_DH] + (gR - gE) * P[_lib.NQ_IS_DREV_DH]
  dL_dl[:, 1:] += (gF + gE) * P[_lib.NQ_IS_DFWD_DL] + (gR - gE) * P[_lib.NQ_IS_DREV_DL]
  dL_dh[:, 0] += g_h0.to(torch.float64)
  dL_dl[:, 0] += g_lt0.to(torch.float64)
  dL_dh[:, -1] += g_hf.to(torch.float64)
  dL_dl[:, -1] += g_ltf.to(torch.float64)
  return loss.to(torch.float32), dL_dh, dL_dl


def _loss_and_sensitivities(kind, run: _Run, log_temp, field, L):
  """Per replica: loss [R] (float32) and d loss / d field_t, d loss / d log_temp_t as [R, T] float64.
  The torch form of what `nq_ising_grad_assemble` computes for the two closed-form losses; kept as the
  checker of that kernel (tests/test_gpu_ising.py), not called by the estimators."""
  n_sites = float(np.prod(_shape_of(L)))
  S = run.summary.to(torch.float64)                 # [7, R, T-1]
  R, Tm1 = S.shape[1], S.shape[2]
  T = Tm1 + 1
  dev = S.device
  lt = torch.from_numpy(np.asarray(log_temp, np.float64)).to(dev)
  h = torch.from_numpy(np.asarray(field, np.float64)).to(dev)
  dL_dh = torch.zeros((R, T), dtype=torch.float64, device=dev)
  dL_dl = torch.zeros((R, T), dtype=torch.float64, device=dev)
  M0 = run.MB0[:, 0].to(torch.float64).expand(R)
  B0 = run.MB0[:, 1].to(torch.float64).expand(R)
  M_after = torch.round(S[_lib.NQ_IS_MAG] * n_sites)              # spin sum after each sweep (exact integers)
  M_before = torch.cat([M0[:, None], M_after[:, :-1]], 1)         # before sweep t = 1..T-1
  if kind == 'work':
    loss = _sum32(run.summary[_lib.NQ_IS_WORK])
    # work_t = -(h_t - h_{t-1}) M_{t-1}
    dL_dh[:, 1:] -= M_before
    dL_dh[:, :-1] += M_before
  else:
    P = run.partials.to(torch.float64)
    dL_dh[:, 1:] += P[_lib.NQ_IS_DFWD_DH] - P[_lib.NQ_IS_DREV_DH]
    dL_dl[:, 1:] += P[_lib.NQ_IS_DFWD_DL] - P[_lib.NQ_IS_DREV_DL]
    # endpoint term log_prob(initial; params[0]) - log_prob(final; params[T-1]), log_prob = (B + h M) e^{-lt}
    e0, ef = torch.exp(-lt[0]), torch.exp(-lt[-1])
    E_fin = S[_lib.NQ_IS_ENERGY][:, -1]                           # = -(B_f + h_{T-1} M_f)
    M_f = M_after[:, -1]
    dL_dh[:, 0] += M0 * e0
    dL_dl[:, 0] += -(B0 + h[0] * M0) * e0
    dL_dh[:, -1] += -M_f * ef
    dL_dl[:, -1] += -E_fin * ef
    # loss value in the reference's float32 arithmetic: -energy / exp(log_temp) at both ends
    e_init32 = (-(B0 + h[0] * M0)).to(torch.float32)
    # tensor / tensor: IEEE division (torch turns division by a host scalar into a reciprocal multiply)
    exp_ends = torch.tensor([np.exp(np.float64(F32(log_temp[0]))), np.exp(np.float64(F32(log_temp[-1])))],
                            dtype=torch.float32, device=dev)
    lp_init = -e_init32 / exp_ends[0]
    lp_fin = -run.summary[_lib.NQ_IS_ENERGY][:, -1] / exp_ends[1]
    loss = (lp_init - lp_fin) + _sum32(run.summary[_lib.NQ_IS_EP])
  return loss, dL_dh, dL_dl


def _schedule_tables(schedule: IsingSchedule, times):
  times = np.asarray(_to_host(times), F32)
  params = schedule(times)
  J_lt = np.asarray(schedule.log_temp.jacobian(times), np.float64)
  J_h = np.asarray(schedule.field.jacobian(times), np.float64)
  return times, np.asarray(params.log_temp, F32), np.asarray(params.field, F32), J_lt, J_h


def _pack_grad_like(schedule: IsingSchedule, g_lt: torch.Tensor, g_h: torch.Tensor, batched: bool):
  """Gradient with the pytree structure of `schedule` (what jax.grad returns); leaves are float32
  CUDA tensors `[n]` (or `[R, n]` for a batch of seeds)."""
  def like(p10n_obj, g):
    leaves, treedef = tree.tree_flatten(p10n_obj)
    out, o = [], 0
    for leaf in leaves:
      shape = np.shape(leaf)
      size = int(np.prod(shape)) if shape else 1
      piece = g[..., o:o + size].to(torch.float32)
      out.append(piece.reshape(tuple(g.shape[:-1]) + tuple(shape)) if batched else piece.reshape(shape))
      o += size
    return tree.tree_unflatten(treedef, out)
  return IsingSchedule(like(schedule.log_temp, g_lt), like(schedule.field, g_h))


_JACOBIAN_CACHE = {}


def _jacobian_on_device(J, dev):
  """Schedule Jacobian [T, D] as a float64 device tensor; uploaded once per distinct table (an
  optimisation loop evaluates the same Jacobian every step: every Parameterization is affine)."""
  J = np.ascontiguousarray(J, np.float64)
  key = (J.shape, str(dev), hash(J.tobytes()))
  hit = _JACOBIAN_CACHE.get(key)
  if hit is None:
    hit = torch.from_numpy(J.copy()).to(dev)
    if len(_JACOBIAN_CACHE) >= 16:
      _JACOBIAN_CACHE.pop(next(iter(_JACOBIAN_CACHE)))
    _JACOBIAN_CACHE[key] = hit
  return hit


def _gradient_terms(loss_function, schedule, times, initial_spins, seed, checkpoint_every=None):
  kind = _loss_kind(loss_function)
  times, log_temp, field, J_lt, J_h = _schedule_tables(schedule, times)
  params = IsingParameters(log_temp, field)
  _, _, spins, L, batched, seeds, broadcast, bits = _prepare(params, initial_spins, seed)
  if checkpoint_every is not None:
    control_flow.check_divides(len(times) - 1, checkpoint_every)
  run = _run_kernel(log_temp, field, bits, broadcast, seeds, L, want_partials=True)
  dev = run.summary.device
  R, T = run.summary.shape[1], len(times)
  Jl = _jacobian_on_device(J_lt, dev)
  Jh = _jacobian_on_device(J_h, dev)
  Dl, Dh = Jl.shape[1], Jh.shape[1]
  ext = (None, None, None)
  if kind == 'generic':
    final_spins = _unpack(run.final_bits, L, torch.zeros((), dtype=torch.float32))
    loss_ext, dL_dh, dL_dl = _generic_loss_sensitivities(loss_function, run, log_temp, field, L, spins, final_spins)
    ext = (dL_dh.to(torch.float64).contiguous(), dL_dl.to(torch.float64).contiguous(),
           loss_ext.to(torch.float32).contiguous())
  if Dl > 32 or Dh > 32:
    raise ValueError(f'at most 32 variables per schedule component (got {Dl}, {Dh})')
  loss = torch.empty(R, dtype=torch.float32, device=dev)
  g = [torch.empty((R, d), dtype=torch.float64, device=dev) for d in (Dl, Dh, Dl, Dh)]
  ends = (C.c_float * 2)(float(F32(log_temp[0])), float(F32(log_temp[-1])))
  n_sites = int(np.prod(_shape_of(L)))
  _lib.check(_lib.lib().nq_ising_grad_assemble(
      {'work': 0, 'entropy': 1, 'generic': 2}[kind], T, R, n_sites, int(run.MB0.shape[0] == 1),
      ends, float(F32(field[0])), _lib.ptr(run.summary), _lib.ptr(run.partials), _lib.ptr(run.MB0),
      _lib.ptr(Jl), Dl, _lib.ptr(Jh), Dh, _lib.ptr(ext[0]), _lib.ptr(ext[1]), _lib.ptr(ext[2]),
      _lib.ptr(loss), _lib.ptr(g[0]), _lib.ptr(g[1]), _lib.ptr(g[2]), _lib.ptr(g[3]), _lib.stream_arg()))
  logP = (g[0], g[1])          # grad of sum(fwd_lp) * stop_grad(loss)
  reparam = (g[2], g[3])       # grad of loss
  if not batched:
    logP = tuple(x[0] for x in logP)
    reparam = tuple(x[0] for x in reparam)
  return dict(logP=logP, reparam=reparam, loss=loss if batched else loss[0],
              summary=_summary_from(run, batched), batched=batched, schedule=schedule, run=run)


def _make_estimator(which, track_terms=False):
  def factory(loss_function: LossFn, checkpoint_every: Optional[int] = None):
    _loss_kind(loss_function)

    def _estimate_gradient(schedule: IsingSchedule, times, initial_spins, seed):
      t = _gradient_terms(loss_function, schedule, times, initial_spins, seed, checkpoint_every)
      full = tuple(a + b for a, b in zip(t['logP'], t['reparam']))
      pick = {'reinforce': full, 'logP': t['logP'], 'reparam': t['reparam']}[which]
      grad = _pack_grad_like(schedule, pick[0], pick[1], t['batched'])
      if track_terms:
        return (grad, _pack_grad_like(schedule, *t['logP'], t['batched']),
                _pack_grad_like(schedule, *t['reparam'], t['batched']), t['summary'])
      return grad, t['summary']
    return _estimate_gradient
  return factory


def _fwd_only(factory):
  def fwd_factory(loss_function):
    return factory(loss_function)
  return fwd_factory


# Reverse- and forward-mode estimators compute the same numbers (tests/ising_test.py:129-145); with
# closed-form sensitivities there is one implementation behind both names.
estimate_gradient_rev = _make_estimator('reinforce')
estimate_gradient_fwd = _fwd_only(estimate_gradient_rev)
estimate_gradient_rev_REINFORCE = _make_estimator('reinforce')
estimate_gradient_fwd_REINFORCE = _fwd_only(estimate_gradient_rev_REINFORCE)
estimate_gradient_rev_logP = _make_estimator('logP')
estimate_gradient_fwd_logP = _fwd_only(estimate_gradient_rev_logP)
estimate_gradient_rev_reparam = _make_estimator('reparam')
estimate_gradient_fwd_reparam = _fwd_only(estimate_gradient_rev_reparam)
estimate_gradient_fwd_REINFORCE_track_terms = _fwd_only(_make_estimator('reinforce', track_terms=True))


# ------------------------------------------------------------------------------ training steps
def _mean_over_batch(grads):
  return tree.tree_map(lambda x: x.to(torch.float64).mean(0).to(torch.float32).cpu().numpy(), grads)


def _make_train_step(rev_factory, fwd_factory, n_extra):
  def get_train_step(optimizer, initial_spins, batch_size: int, time_steps: int,
                     loss_function: LossFn = total_entropy_production, mode: str = 'rev', max_grad=1e4):
    """optimizer = (init_fn, update_fn, params_fn) as in jax.example_libraries.optimizers
    (see noneq_opt_b200.optimizers).  Returns train_step(opt_state, step, seed)."""
    from . import optimizers
    if mode == 'rev':
      gradient_estimator = rev_factory(loss_function)
    elif mode == 'fwd':
      gradient_estimator = fwd_factory(loss_function)
    else:
      raise ValueError(f'`mode` must be either "fwd" or "rev"; got {mode}.')
    times = np.linspace(0, 1, time_steps, dtype=F32)
    update_fn, params_fn = optimizer[1], optimizer[2]

    def _train_step(opt_state, step, seed):
      seeds = nqrandom.split(seed, batch_size)
      schedule = params_fn(opt_state)
      out = gradient_estimator(schedule, times, initial_spins, seeds)
      grads, summary = out[0], out[-1]
      mean_grad = _mean_over_batch(grads)
      opt_state = update_fn(step, optimizers.clip_grads(mean_grad, max_grad), opt_state)
      extras = (mean_grad,)[:min(n_extra, 1)]
      if n_extra == 3:
        if len(out) == 4:
          extras = (mean_grad, _mean_over_batch(out[1]), _mean_over_batch(out[2]))
        else:   # mode == 'rev' returns only the full gradient (ising.py:611)
          extras = (mean_grad,)
      return (opt_state, summary) + tuple(extras)
    return _train_step
  return get_train_step


get_train_step = _make_train_step(estimate_gradient_rev, estimate_gradient_fwd, 0)
get_train_step_REINFORCE = _make_train_step(estimate_gradient_rev_REINFORCE, estimate_gradient_fwd_REINFORCE, 1)
get_train_step_logP = _make_train_step(estimate_gradient_rev_logP, estimate_gradient_fwd_logP, 1)
get_train_step_reparam = _make_train_step(estimate_gradient_rev_reparam, estimate_gradient_fwd_reparam, 1)
get_train_step_REINFORCE_track_terms = _make_train_step(estimate_gradient_rev_REINFORCE,
                                                        estimate_gradient_fwd_REINFORCE_track_terms, 3)
