"""Device-resident optimisation loop of the barrier-crossing driver.

`run_barrier_crossing_opt.py:113-137` performs, once per training step and from Python:
`key, k_eq, k_sim = split(key, 3)`; Boltzmann initial positions (`mapped_brownian_eqm`);
`estimate_gradient_boltzinit(...)(coeffs, init_pos, k_sim)`; the Adam update.  Under XLA each of
those is a separate dispatch with host round trips in between; for small ensembles (config 1:
10^3 x 10^3) that latency is the whole step.  `LangevinTrainer` keeps every operand of the step in
device memory -- chain key, seeds, trap table, gradient sums, Adam moments, step counter -- issues
the same sequence through the C ABI once while a CUDA graph is being captured, and then replays
that graph: one `cudaGraphLaunch` per training step, no host synchronisation, no allocation.

The arithmetic is the same as the step-by-step path (`barrier_crossing.estimate_gradient_*` +
`optimizers.adam`): same key stream, same float32 schedule evaluation order, same kernels, the
Adam update in binary32 with one rounding per operation.  tests/test_gpu_training.py checks the
coefficients after several steps against that path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, barrier_crossing as bc, parameterization as p10n, random as nqrandom, tree


def _affine_form(coeffs, S):
  """(variables [D], basis [S-1, D], offset [S-1] | None): the trap protocol on the reference's
  scaled step grid as an affine function of its variables (every Parameterization in scope is)."""
  if isinstance(coeffs, p10n.Parameterization):
    if not coeffs.is_affine:
      raise NotImplementedError('the device-resident trainer rebuilds the protocol table on the device as basis @ variables '
                                '+ offset: the protocol must be affine in its variables (Spline(variable_knots=True) is not; '
                                'use the step-by-step estimators)')
    tv = np.arange(S + 1)[1:-1]
    scaled = ((tv - 1).astype(np.float32) / np.float32(tv[-1] - 1)).astype(np.float32)
    leaves = [np.asarray(l, np.float32).reshape(-1) for l in tree.tree_leaves(coeffs.variables)]
    v0 = np.concatenate(leaves).astype(np.float32)
    jac = np.asarray(coeffs.jacobian(scaled), np.float32)
    offset = (np.asarray(coeffs(scaled), np.float64) - jac.astype(np.float64) @ v0.astype(np.float64)).astype(np.float32)
    return v0, jac, (offset if np.any(offset != 0) else None)
  if isinstance(coeffs, torch.Tensor):
    coeffs = coeffs.detach().cpu().numpy()
  w = np.asarray(coeffs, np.float32).reshape(-1)
  _, basis = bc._chebyshev_grid_basis(S + 1, w.shape[0])
  return w.copy(), np.asarray(basis, np.float32), None


class LangevinTrainer:
  """One CUDA graph per training step of `run_barrier_crossing_opt.py:113-137`.

  Args (names of the reference driver / estimator factory):
    batch_size, energy_fn, coeffs (Chebyshev coefficient vector or a Parameterization), r0_init,
    r0_final, Neq, shift, simulation_steps, dt, temperature, mass, gamma: as
      `estimate_gradient_boltzinit` (barrier_crossing.py:216-224).
    key: the driver's chain key (uint32[2]).
    eqm: None (every step starts from `init_position`) or a dict(sim_length=..., dt=..., kT=...,
      gamma=...) for the `mapped_brownian_eqm` pre-pass of :127 (no reference quirk: the values
      given are the values used).
    term: 'reinforce' | 'logP' | 'reparam' (barrier_crossing.py:204-333).
    step_size, decay_steps, decay_rate, b1, b2, eps: `adam(exponential_decay(...))` of :113.
    max_grad: optional `clip_grads` bound.
  """

  def __init__(self, batch_size, energy_fn, coeffs, r0_init, r0_final, Neq, shift, simulation_steps, dt,
               temperature, mass, gamma, key, *, init_position=None, eqm=None, term='reinforce',
               step_size=0.1, decay_steps=0.0, decay_rate=1.0, b1=0.9, b2=0.999, eps=1e-8, max_grad=None,
               use_graph=True):
    self.device = dev = _lib.require_cuda()
    self.lib = _lib.lib()
    self.B, self.S = int(batch_size), int(simulation_steps)
    v0, basis, offset = _affine_form(coeffs, self.S)
    self.D = D = int(v0.shape[0])
    if D > 32:
      raise ValueError(f'the fused gradient kernel handles up to 32 protocol variables (got {D})')
    f32 = dict(dtype=torch.float32, device=dev)
    self.variables = torch.from_numpy(v0).to(dev)
    self.basis = torch.from_numpy(np.array(basis, dtype=np.float32, order='C')).to(dev)
    self.offset = None if offset is None else torch.from_numpy(offset).to(dev)
    self.r0_init, self.r0_final = float(r0_init), float(r0_final)
    self.r_table = torch.empty(self.S + 1, **f32)
    self.adam_m, self.adam_v = torch.zeros(D, **f32), torch.zeros(D, **f32)
    self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
    self.grad = torch.zeros(D, **f32)
    self.key = nqrandom.as_key_tensor(key, dev).reshape(2).clone()
    self._keys3 = torch.empty((3, 2), dtype=self.key.dtype, device=dev)
    self.seeds = torch.empty((self.B, 2), dtype=self.key.dtype, device=dev)
    self._seeds_eq = torch.empty((self.B, 2), dtype=self.key.dtype, device=dev) if eqm else None
    x_init = self.r0_init if init_position is None else float(np.asarray(init_position).reshape(-1)[0])
    self._x_init = torch.full((self.B,), x_init, **f32)
    self.x0 = self._x_init.clone()
    self.work, self.logp = torch.empty(self.B, **f32), torch.empty(self.B, **f32)
    self.estimator, self.x_final = torch.empty(self.B, **f32), torch.empty(self.B, **f32)
    self.grad_sums = torch.zeros((2, D), dtype=torch.float64, device=dev)
    self.moments = torch.zeros(_lib.NQ_MOM_COUNT, dtype=torch.float64, device=dev)
    self._ws_bytes = self.lib.nq_langevin_workspace_bytes(self.B, D)
    self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=dev)
    self._desc = bc._desc(energy_fn, shift, self.S, Neq, dt, temperature, mass, gamma)
    self._desc_eq = None
    if eqm:
      self._desc_eq = bc._desc(energy_fn, shift, int(eqm['sim_length']), 0, eqm['dt'], eqm['kT'], 1.0, eqm['gamma'])
    w_logp, w_rep = bc._TERMS[term]
    self._adam = _lib.NqAdamDesc(float(step_size), float(decay_steps), float(decay_rate), float(b1), float(b2),
                                 float(eps), float(max_grad or 0.0), w_logp, w_rep, float(self.B))
    self._graph = None
    self._use_graph = bool(use_graph)
    self.launches_per_step = None

  # -- the launch sequence of one step (captured once, then replayed)
  def _enqueue(self):
    lib, st, ptr, check = self.lib, _lib.stream_arg(), _lib.ptr, _lib.check
    check(lib.nq_random_split_dev(ptr(self.key), 3, 0, 3, ptr(self._keys3), st))         # :125
    self.key.copy_(self._keys3[0])
    if self._desc_eq is not None:                                                          # :127
      check(lib.nq_random_split_dev(ptr(self._keys3[1]), self.B, 0, self.B, ptr(self._seeds_eq), st))
      check(lib.nq_langevin_stationary(C.byref(self._desc_eq), self.r0_init, ptr(self._x_init),
                                       ptr(self._seeds_eq), self.B, ptr(self.x0), None, st))
    check(lib.nq_random_split_dev(ptr(self._keys3[2]), self.B, 0, self.B, ptr(self.seeds), st))   # barrier_crossing.py:221
    check(lib.nq_schedule_affine(ptr(self.basis), ptr(self.variables), ptr(self.offset), self.r0_init,
                                 self.r0_final, self.S - 1, self.D, ptr(self.r_table), st))       # :145-151
    check(lib.nq_langevin_grad(C.byref(self._desc), ptr(self.r_table), ptr(self.basis), self.D, ptr(self.x0),
                               ptr(self.seeds), self.B, ptr(self.work), ptr(self.logp), ptr(self.estimator),
                               ptr(self.x_final), ptr(self.grad_sums), ptr(self.moments), ptr(self._ws),
                               self._ws_bytes, st))                                               # :204-224
    check(lib.nq_adam_step(C.byref(self._adam), self.D, ptr(self.grad_sums[0]), ptr(self.grad_sums[1]), None,
                           ptr(self.variables), ptr(self.adam_m), ptr(self.adam_v), ptr(self.step_count),
                           ptr(self.grad), st))                                                   # :129

  def _capture(self):
    before = self.lib.nq_launch_count()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    state = [t.clone() for t in (self.key, self.variables, self.adam_m, self.adam_v, self.step_count)]
    with torch.cuda.stream(side):   # warm-up outside capture (sets kernel attributes, loads modules)
      self._enqueue()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    self.launches_per_step = int(self.lib.nq_launch_count() - before)
    for t, s in zip((self.key, self.variables, self.adam_m, self.adam_v, self.step_count), state):
      t.copy_(s)
    if self._use_graph:
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph):
        self._enqueue()
      self._graph = graph

  def step(self):
    """One training step; returns nothing and does not synchronise."""
    if self.launches_per_step is None:
      self._capture()
    if self._graph is not None:
      self._graph.replay()
    else:
      self._enqueue()

  def run(self, opt_steps, record_works=False):
    """`opt_steps` steps back to back.  Returns dict(mean_work [opt_steps] f64, coeffs [opt_steps+1, D]
    (row j = variables before step j), works [opt_steps, B] if record_works) as device tensors."""
    n = int(opt_steps)
    mom = torch.empty((n, _lib.NQ_MOM_COUNT), dtype=torch.float64, device=self.device)
    hist = torch.empty((n + 1, self.D), dtype=torch.float32, device=self.device)
    works = torch.empty((n, self.B), dtype=torch.float32, device=self.device) if record_works else None
    for j in range(n):
      hist[j].copy_(self.variables)
      self.step()
      mom[j].copy_(self.moments)
      if works is not None:
        works[j].copy_(self.work)
    hist[n].copy_(self.variables)
    out = dict(mean_work=mom[:, 1] / mom[:, 0], coeffs=hist)
    if works is not None:
      out['works'] = works
    return out

  @property
  def coeffs(self):
    return self.variables


# ================================================================================ Ising
def _ising_component(param, times):
  """(variables [D] f32, basis [T, D] f32, scale | None, offset1 | None, offset2 | None) such that
  value(times) = ((scale * (basis @ variables)) + offset1) + offset2 in the float32 operation order of the
  parameterisation itself -- exact (bit-equal to `param(times)`) for the drivers' composition
  AddBaseline(ConstrainEndpoints(Chebyshev)) and its sub-forms; any other affine Parameterization goes
  through its Jacobian and value at zero (equal up to float32 rounding)."""
  F32 = np.float32
  times = np.asarray(times, F32)
  offset2 = None
  inner = param
  if isinstance(inner, p10n.AddBaseline):
    offset2 = np.asarray(inner.baseline(times), F32)
    inner = inner.wrapped
  scale = offset1 = None
  if isinstance(inner, p10n.ConstrainEndpoints) and isinstance(inner.wrapped, p10n.Chebyshev):
    x0, x1 = (F32(v) for v in inner.domain)
    scale = inner._envelope(times)
    offset1 = (F32(inner.y0) + (times - x0) / (x1 - x0) * F32(inner.y1 - inner.y0)).astype(F32)
    inner = inner.wrapped
  if isinstance(inner, p10n.Chebyshev) and np.ndim(inner.weights) == 1:
    return np.asarray(inner.weights, F32).copy(), np.asarray(inner.jacobian(times), F32), scale, offset1, offset2
  # generic affine form of the whole component
  if not param.is_affine:
    raise NotImplementedError('IsingTrainer rebuilds the schedule tables on the device as basis @ variables + offset: every '
                              'schedule component must be affine in its variables (Spline(variable_knots=True) is not)')
  leaves = [np.asarray(l, F32).reshape(-1) for l in tree.tree_leaves(param.variables)]
  v0 = np.concatenate(leaves).astype(F32)
  jac = np.asarray(param.jacobian(times), F32)
  off = (np.asarray(param(times), np.float64) - jac.astype(np.float64) @ v0.astype(np.float64)).astype(F32)
  return v0, jac, None, off, None


class IsingTrainer:
  """One CUDA graph per training step of the Ising drivers (`run_ising_opt.py:77-102` around
  `ising.get_train_step`, ising.py:263-290): per step

      seed              next(seed_stream)                                     (:57-61)
      seeds           = split(seed, batch_size)                               (:283)
      schedule tables = params_fn(opt_state)(times)                           (:284, parameterization.py)
      class tables    -> checkerboard sweeps -> per-replica REINFORCE / logP / reparam gradient (:179-236)
      mean over the batch, clip_grads, Adam                                   (:286-288)

  with every operand resident in HBM: chain key, replica keys, schedule variables and tables, the class
  tables (`nq_ising_class_tables`), summaries, per-replica gradients, Adam moments and step counter.  The launch
  sequence is issued once through the C ABI under CUDA-graph capture and replayed: no host synchronisation, no
  host arithmetic, no allocation per step (the step-by-step `get_train_step` builds its tables with numpy and
  reads the batch mean back).

  Args: schedule (ising.IsingSchedule of Parameterizations), initial_spins (one lattice), batch_size, time_steps,
    key (uint32[2]; `PRNGKey(seed)` of seed_stream), loss_function (ising.total_entropy_production |
    ising.total_work), term ('reinforce' | 'logP' | 'reparam'), max_grad (clip_grads bound of get_train_step),
    step_size / decay_steps / decay_rate / b1 / b2 / eps (adam(exponential_decay(...))),
    seed_source: 'seed_stream' -- the reference's generator, which yields the NEW CHAIN KEY itself (:57-61) --
      or 'split' -- `key, sub = split(key)` and the step runs on `sub`.
  """

  def __init__(self, schedule, initial_spins, batch_size, time_steps, key, *,
               loss_function=None, term='reinforce', max_grad=1e4, step_size=0.1, decay_steps=0.0, decay_rate=1.0,
               b1=0.9, b2=0.999, eps=1e-8, seed_source='seed_stream', use_graph=True):
    from . import ising
    self.device = dev = _lib.require_cuda()
    self.lib = _lib.lib()
    loss_function = ising.total_entropy_production if loss_function is None else loss_function
    kind = ising._loss_kind(loss_function)
    if kind == 'generic':
      raise NotImplementedError('IsingTrainer runs the two closed-form losses (total_work, total_entropy_production); '
                                'use ising.get_train_step for an arbitrary LossFn')
    if seed_source not in ('seed_stream', 'split'):
      raise ValueError("seed_source must be 'seed_stream' or 'split'")
    self._kind = {'work': 0, 'entropy': 1}[kind]
    self._seed_row = 0 if seed_source == 'seed_stream' else 1
    self.R, self.T = int(batch_size), int(time_steps)
    self.schedule = schedule
    self.times = np.linspace(0, 1, self.T, dtype=np.float32)
    f32 = dict(dtype=torch.float32, device=dev)
    comps = [_ising_component(schedule.log_temp, self.times), _ising_component(schedule.field, self.times)]
    self.Dl, self.Dh = int(comps[0][0].shape[0]), int(comps[1][0].shape[0])
    if self.Dl > 32 or self.Dh > 32:
      raise ValueError(f'at most 32 variables per schedule component (got {self.Dl}, {self.Dh})')
    P = self.Dl + self.Dh
    self.variables = torch.from_numpy(np.concatenate([comps[0][0], comps[1][0]])).to(dev)

    def up(a):
      return None if a is None else torch.from_numpy(np.array(a, dtype=np.float32, order="C")).to(dev)
    self._comp = [tuple(up(a) for a in c[1:]) for c in comps]      # (basis, scale, offset1, offset2) per component
    self._J = [torch.from_numpy(np.ascontiguousarray(np.asarray(p.jacobian(self.times), np.float64))).to(dev)
               for p in (schedule.log_temp, schedule.field)]
    self.log_temp, self.field = torch.empty(self.T, **f32), torch.empty(self.T, **f32)
    # lattice: packed once, its spin / bond sums measured once (the initial state is the same every step)
    spins = ising._as_cuda(initial_spins)
    shape = tuple(int(v) for v in spins.shape)
    if any(v % 2 for v in shape):
      raise ValueError(f'All entries in `shape` must be even; got {shape}.')
    self._shape, ndim = shape, len(shape)
    self._bits = ising._pack(spins, ndim).reshape(1, -1)
    words = (int(np.prod(shape)) + 31) // 32
    dims = (C.c_int32 * 3)(*(list(shape) + [1] * (3 - ndim)))
    square = ndim == 2 and shape[0] == shape[1]
    self._desc = _lib.NqIsingDesc(L=shape[0], T=self.T, R=self.R, spins0_broadcast=1, sweep_key_mode=0,
                                  ndim=0 if square else ndim, dims=dims)
    self._ndim = ndim
    self._MB0 = torch.empty((1, 2), dtype=torch.int64, device=dev)
    _lib.check(self.lib.nq_ising_measure(_lib.ptr(self._bits), ndim, dims, 1, _lib.ptr(self._MB0), _lib.stream_arg()))
    R, Tm1 = self.R, self.T - 1
    self._thr = torch.empty((Tm1, 16), dtype=torch.int32, device=dev)
    self._lut = torch.empty((Tm1, 4, 16), dtype=torch.float64, device=dev)
    self.final_bits = torch.empty((R, words), dtype=torch.int32, device=dev)
    self.summary_raw = torch.zeros((_lib.NQ_IS_NSUMMARY, R, Tm1), **f32)
    self._partials = torch.zeros((_lib.NQ_IS_NPARTIAL, R, Tm1), **f32)
    self._ws_bytes = self.lib.nq_ising_workspace_bytes(C.byref(self._desc))
    self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=dev) if self._ws_bytes else None
    self.loss = torch.empty(R, **f32)
    self._g = [torch.empty((R, d), dtype=torch.float64, device=dev) for d in (self.Dl, self.Dh, self.Dl, self.Dh)]
    self.grad = torch.zeros(P, **f32)          # batch-mean gradient (before clipping)
    self.grad_clipped = torch.zeros(P, **f32)
    self.adam_m, self.adam_v = torch.zeros(P, **f32), torch.zeros(P, **f32)
    self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
    self.key = nqrandom.as_key_tensor(key, dev).reshape(2).clone()
    self._keys2 = torch.empty((2, 2), dtype=self.key.dtype, device=dev)
    self.seeds = torch.empty((R, 2), dtype=self.key.dtype, device=dev)
    w_logp, w_rep = {'reinforce': (1.0, 1.0), 'logP': (1.0, 0.0), 'reparam': (0.0, 1.0)}[term]
    self._w = (w_logp, w_rep)
    self._adam = _lib.NqAdamDesc(float(step_size), float(decay_steps), float(decay_rate), float(b1), float(b2),
                                 float(eps), float(max_grad or 0.0), 1.0, 0.0, 1.0)
    self._n_sites = int(np.prod(shape))
    self._graph = None
    self._use_graph = bool(use_graph)
    self.launches_per_step = None

  def _enqueue(self):
    lib, st, ptr, check = self.lib, _lib.stream_arg(), _lib.ptr, _lib.check
    check(lib.nq_random_split_dev(ptr(self.key), 2, 0, 2, ptr(self._keys2), st))                       # seed_stream :57-61
    self.key.copy_(self._keys2[0])
    check(lib.nq_random_split_dev(ptr(self._keys2[self._seed_row]), self.R, 0, self.R, ptr(self.seeds), st))   # :283
    for tab, (basis, scale, off1, off2), o, D in ((self.log_temp, self._comp[0], 0, self.Dl),
                                                  (self.field, self._comp[1], self.Dl, self.Dh)):
      check(lib.nq_schedule_compose(ptr(basis), ptr(self.variables[o:o + D]), ptr(scale), ptr(off1), ptr(off2), self.T, D,
                                    ptr(tab), st))                                                      # :284
    check(lib.nq_ising_class_tables(ptr(self.log_temp), ptr(self.field), self.T, self._ndim, ptr(self._thr),
                                    ptr(self._lut), st))
    check(lib.nq_ising_run(C.byref(self._desc), ptr(self._thr), ptr(self._lut), ptr(self.field), ptr(self.log_temp),
                           ptr(self._bits), ptr(self.seeds), ptr(self.final_bits), ptr(self.summary_raw),
                           ptr(self._partials), None, None, ptr(self._ws), self._ws_bytes, st))         # :153-173
    check(lib.nq_ising_grad_assemble_dev(self._kind, self.T, self.R, self._n_sites, 1, ptr(self.log_temp),
                                         ptr(self.field), ptr(self.summary_raw), ptr(self._partials), ptr(self._MB0),
                                         ptr(self._J[0]), self.Dl, ptr(self._J[1]), self.Dh, ptr(self.loss),
                                         ptr(self._g[0]), ptr(self._g[1]), ptr(self._g[2]), ptr(self._g[3]), st))   # :179-198
    check(lib.nq_ising_grad_mean(ptr(self._g[0]), ptr(self._g[1]), ptr(self._g[2]), ptr(self._g[3]), self.R, self.Dl,
                                 self.Dh, self._w[0], self._w[1], ptr(self.grad), st))                  # :286
    check(lib.nq_adam_step(C.byref(self._adam), self.Dl + self.Dh, None, None, ptr(self.grad), ptr(self.variables),
                           ptr(self.adam_m), ptr(self.adam_v), ptr(self.step_count), ptr(self.grad_clipped), st))  # :288

  _capture = LangevinTrainer._capture
  step = LangevinTrainer.step

  @property
  def summary(self):
    """IsingSummary [batch, T-1] of the LAST step (device tensors; views of the trainer's buffer)."""
    from . import ising
    return ising.IsingSummary(*[self.summary_raw[f] for f in range(_lib.NQ_IS_NSUMMARY)])

  def current_schedule(self):
    """The schedule pytree with the current variables (one device-to-host copy of Dl + Dh floats)."""
    w = self.variables.cpu().numpy()
    leaves, treedef = tree.tree_flatten(self.schedule)
    out, o = [], 0
    for leaf in leaves:
      size = int(np.prod(np.shape(leaf))) if np.shape(leaf) else 1
      out.append(w[o:o + size].reshape(np.shape(leaf)).copy())
      o += size
    return tree.tree_unflatten(treedef, out)

  def run(self, training_steps, record_summaries=False):
    """`training_steps` steps back to back, no host synchronisation.  Returns dict(mean_loss [n],
    variables [n + 1, Dl + Dh] (row j = before step j), summaries [n, 7, batch, T-1] if asked) as device tensors."""
    n = int(training_steps)
    hist = torch.empty((n + 1, self.Dl + self.Dh), dtype=torch.float32, device=self.device)
    mean_loss = torch.empty(n, dtype=torch.float32, device=self.device)
    summaries = (torch.empty((n,) + tuple(self.summary_raw.shape), dtype=torch.float32, device=self.device)
                 if record_summaries else None)
    for j in range(n):
      hist[j].copy_(self.variables)
      self.step()
      mean_loss[j] = self.loss.mean()
      if summaries is not None:
        summaries[j].copy_(self.summary_raw)
    hist[n].copy_(self.variables)
    out = dict(mean_loss=mean_loss, variables=hist)
    if summaries is not None:
      out['summaries'] = summaries
    return out
