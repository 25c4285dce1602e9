#!/usr/bin/env python
"""Benchmark of the noneq_opt ensemble hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port, all host threads)

Headline workload (BASELINE.json configs[1]): barrier_crossing -- double well with moving trap,
100 000 trajectories x 10 000 steps, forward + gradient of the REINFORCE work objective wrt the 13
Chebyshev protocol coefficients.  One "step" = one such gradient evaluation per GPU (1e9
trajectory-steps); with N GPUs every rank runs its own 100 000-trajectory slice of the global batch
(weak scaling) and ONE NCCL all-reduce combines the gradient and work moments.

Also reported in the same line (`ising`): spin-updates/s of the Glauber sweep at configs[2]
(64x64 x 4096 replicas x 1000 sweeps, with the REINFORCE gradient) and configs[3] (1024x1024
bit-packed x 256 replicas; throughput per sweep is independent of T, the bench runs 100 sweeps).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- config 2: barrier_crossing
KT = 4.183
BETA = 1.0 / KT
X_M = 14.0
KAPPA = 21.3863 / (BETA * X_M ** 2)
K_TRAP = 1.5
GAMMA = KT / 1e7
DT = 1e-8
C2_COEFFS = np.array([2.00000000e+01, 1.99599600e+01, -2.63245988e-15, 8.41484236e-16, 0.00000000e+00,
                      -7.47209997e-15, -1.69086360e-14, 1.27892999e-14, 3.02606960e-15, -6.05805446e-15,
                      -6.36600761e-16, -5.53701086e-15, -1.14938537e-14], np.float32)
C2_B, C2_S = 100_000, 10_000
C5_B = 10_000_000                # configs[4]: the same path, 10^7 trajectories sharded over the GPUs
INSTR_PER_TRAJ_STEP = 326.0      # 3 threefry blocks (219 INT) + ~75 FP32 + ~6 SFU + 2*13 FFMA
INSTR_PER_SPIN_UPDATE = 47.0     # half a threefry block (36.5) + ~10 logic/compare/popc
ISSUE_PEAK = 148 * 128 * 1.965e9  # thread-instructions/s at max SM clock
L2_FLUSH_BYTES = 256 << 20


def profiled_dram_traffic(math_mode):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the Langevin kernel, from the newest
  committed `ncu --set full` summary under profiles/ (None if there is none)."""
  import csv
  import glob
  files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles',
                                        f'r*_langevin*_{math_mode}_full.csv')))
  if not files:
    return None, None
  scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  total = 0.0
  for row in csv.reader(open(files[-1])):
    if len(row) >= 3 and row[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
      total += float(row[2]) * scale.get(row[1], 1.0)
  return total, os.path.relpath(files[-1], os.path.dirname(os.path.abspath(__file__)))


def measured_peaks():
  try:
    return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
  except Exception:
    return None


class ClockSampler:
  """SM clock / throttle reasons sampled DURING the timed region: NVML polled from a thread (starts
  instantly, ~5 ms period); `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')
  NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

  def __init__(self, index):
    self.index, self.rows, self.proc, self.thread = index, [], None, None
    self.stop_flag = threading.Event()
    self.source = None

  def _nvml_handle(self):
    import pynvml
    import torch
    pynvml.nvmlInit()
    try:
      uuid = str(torch.cuda.get_device_properties(self.index).uuid)
      return pynvml, pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid if not uuid.startswith('GPU-') else uuid).encode())
    except Exception:
      visible = os.environ.get('CUDA_VISIBLE_DEVICES')
      phys = int(visible.split(',')[self.index]) if visible and visible.split(',')[self.index].isdigit() else self.index
      return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

  def _sample_nvml(self, nv, h, mx):
    masks = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
             nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
    try:
      bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
      self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx),
                        str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                       ['Active' if bits & m else 'Not Active' for m in masks])
    except Exception:
      pass

  def _poll_nvml(self, nv, h):
    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
    while not self.stop_flag.is_set():
      self._sample_nvml(nv, h, mx)
      time.sleep(0.005)

  def sample_now(self):
    """One synchronous sample from the calling thread (the bench calls it in a loop while the device works through the
    already enqueued timed steps).  Host-side only: the steps are timed by CUDA events."""
    if self.source == 'nvml' and getattr(self, '_nvml', None):
      nv, h = self._nvml
      try:
        self._sample_nvml(nv, h, nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
      except Exception:
        pass

  def prepare(self):
    """NVML initialisation and a first query, done BEFORE the warm-up so that starting the sampler at
    the top of the timed region costs nothing."""
    try:
      self._nvml = self._nvml_handle()
      nv, h = self._nvml
      nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
      nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
      self._nvml = None
    return self

  def start(self, polling_thread=True):
    """polling_thread=False: no background thread; the caller samples with sample_now() while it waits for the device
    (bench: after every timed step has been enqueued).  A thread that queries NVML during the enqueue phase delays
    rank 0's launches, and at N > 1 every other rank then waits for rank 0 inside the step's all-reduce (measured at
    8 GPUs: 6.8e11 with per-step queries on rank 0 against 7.6e11 without)."""
    try:
      nv, h = self._nvml if getattr(self, '_nvml', None) else self._nvml_handle()
      self._nvml = (nv, h)
      if polling_thread:
        self.thread = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
        self.thread.start()
      else:
        self.thread = None
      self.source = 'nvml'
      return
    except Exception:
      self.thread = None
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                    '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
      self.source = 'nvidia-smi'
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.source is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvml and nvidia-smi unavailable'])
    if self.source == 'nvml':
      self.stop_flag.set()
      if self.thread is not None:
        self.thread.join(timeout=2)
    else:
      time.sleep(0.12)
      self.proc.terminate()
      try:
        self.proc.wait(timeout=5)
      except Exception:
        self.proc.kill()
    sm, mx, reasons, power = [], [], set(), []
    for r in list(self.rows):
      try:
        sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
        for nm, v in zip(self.NAMES, r[3:7]):
          if v.lower().startswith('active'):
            reasons.add(nm)
      except Exception:
        pass
    if not sm:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'], source=self.source)
    busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
    return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                samples=len(sm), power_w_max=float(max(power)), source=self.source)


# =============================================================================== reference arm
def run_reference(args):
  """The reference's CPU implementation of the path.  JAX cannot be installed in this image, so this
  is the oracle PORT (oracle/c, OpenMP on all host threads), on a bounded sample of config 2."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from oracle import c_oracle, jaxlike as J, langevin_ref as L
  # every core of the host, at every --gpus N: torchrun exports OMP_NUM_THREADS=1 to its workers, which made the
  # N > 1 reference arms of round 1 run on ONE core.  The OpenMP team is set explicitly instead.
  cores = c_oracle.use_all_cores()
  pot = L.V_biomolecule_shifted(KAPPA, KAPPA, X_M, 0., K_TRAP, BETA)
  r, phi = L.make_schedule_chebyshev(C2_S, C2_COEFFS, 0., 40.)
  n = 1024 * cores
  seeds = J.split(J.PRNGKey(0), C2_B)[:n]
  x0 = np.zeros(n, np.float32)
  times = []
  for i in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    c_oracle.langevin(pot, r, phi, x0, seeds, C2_S, 0, DT, KT, GAMMA, 1.0)
    if i >= args.warmup:
      times.append(time.perf_counter() - t0)
  ms = 1e3 * float(np.mean(times))
  value = n * C2_S / (ms * 1e-3)
  assert c_oracle.max_threads() == cores
  line = dict(impl='reference', metric='langevin_traj_steps_per_s', value=value, unit='traj-steps/s',
              n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
              scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
              config=workload_config(args.gpus),
              cpu_baseline=dict(value=value, unit='traj-steps/s', cores=cores, kind='port',
                                sample=f'{n} of {C2_B} trajectories x {C2_S} steps per step, C/OpenMP oracle port '
                                       '(JAX itself is not installable here)',
                                normalisation='whole host: the same all-core CPU rate at every --gpus N (the host does '
                                              'not grow with N); divide the native value by n_gpus for a per-GPU ratio'),
              e2e=dict(value=value, unit='traj-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  emit(line)


def workload_config(n_gpus):
  return dict(workload='barrier_crossing: V_biomolecule_shifted double well + moving trap, '
                       f'{C2_B} trajectories/GPU x {C2_S} steps, forward + REINFORCE work gradient wrt 13 Chebyshev coeffs',
              trajectories_per_gpu=C2_B, steps_per_trajectory=C2_S, coefficients=13, kT=KT, dt=DT,
              parallelism=f'ensemble-sharded x{n_gpus} + 1 allreduce', l2='flushed between timed iterations '
              f'({L2_FLUSH_BYTES >> 20} MiB write); working set is registers + 24 B/trajectory + the 32 MB normal-draw '
              'table, which every timed iteration therefore reads cold',
              normal_draws='every step hashes its own threefry key and maps the 23 bits jax.random.normal keeps through a '
                           'table of that function over all 2^23 patterns (32 MB per device, written once by the kernels\' '
                           'own device function on the first call, i.e. during the warm-up): a tabulated function, not cached results')


# =============================================================================== native arm
def run_native(args):
  import torch
  import torch.distributed as dist
  from noneq_opt_b200 import _lib, barrier_crossing as bc, distributed as nqd, random as nqr, space
  rank, ws = nqd.init_from_env('nccl')
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  lib = _lib.lib()   # raises when the extension is missing: no fallback
  bc.set_math_mode(args.math)
  _, shift = space.free()
  pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, X_M, 0., K_TRAP, BETA)
  B_global = C2_B * ws
  first, count = nqd.shard_range(B_global, rank, ws)
  seed = nqr.PRNGKey(0)
  flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

  # ---- device-resident inputs (the `value` leg): seeds, start positions, tables already in HBM
  seeds = nqr.split(seed, B_global, first=first, count=count)
  # Boltzmann start: 1000 stationary-trap steps at r0 = 0 with the SAME dt/kT/gamma
  eq_key = nqr.split_host(seed, 3)[1]
  x0 = bc._stationary_final(pot, np.zeros((1, 1), np.float32), 0.0, shift, nqr.split(eq_key, B_global, first, count),
                            1000, DT, KT, GAMMA).reshape(-1).contiguous()
  r_np, phi_np = bc._schedule_and_jacobian(np.arange(C2_S + 1), C2_COEFFS, 0., 40.)
  r_dev = torch.from_numpy(r_np).to(dev)
  phi_dev = torch.from_numpy(np.array(phi_np)).to(dev)

  def step_resident():
    out = bc._simulate(pot, shift, r_dev, phi_dev, x0, seeds, C2_S, 0, DT, KT, 1.0, GAMMA)
    if ws > 1:
      dist.all_reduce(out['sums_flat'], op=dist.ReduceOp.SUM)   # [grad sums | moments]: one in-place collective
    gs, mom = out['grad_sums'], out['moments']
    return (gs[0] + gs[1]) / B_global, mom

  # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
  coeffs_host = torch.from_numpy(C2_COEFFS.copy()).pin_memory()
  x0_host = x0.cpu().pin_memory()
  out_host = dict(grad=torch.empty(13, dtype=torch.float32).pin_memory(),
                  work=torch.empty(count, dtype=torch.float32).pin_memory(),
                  logp=torch.empty(count, dtype=torch.float32).pin_memory(),
                  est=torch.empty(count, dtype=torch.float32).pin_memory())
  grad_fn = nqd.langevin_gradient_sharded(B_global, pot, 0., 40., 0, shift, C2_S, DT, KT, 1.0, GAMMA)
  h2d = coeffs_host.numel() * 4 + x0_host.numel() * 4 + 8
  d2h = sum(t.numel() * 4 for t in out_host.values())

  def step_e2e():
    coeffs = coeffs_host.to(dev, non_blocking=True)
    xs = x0_host.to(dev, non_blocking=True)
    grad, aux = grad_fn(coeffs, xs, seed)
    out_host['grad'].copy_(grad, non_blocking=True)
    out_host['work'].copy_(aux['local']['work'], non_blocking=True)
    out_host['logp'].copy_(aux['local']['logp'], non_blocking=True)
    out_host['est'].copy_(aux['local']['estimator'], non_blocking=True)
    torch.cuda.synchronize()
    return out_host['grad']

  def timed(fn, steps, warmup, sample_clocks=False):
    sampler = ClockSampler(local).prepare() if sample_clocks else None
    for _ in range(warmup):
      fn()
    torch.cuda.synchronize()
    if ws > 1:
      dist.barrier()
    if sampler:
      sampler.start(polling_thread=False)
    launches0 = lib.nq_launch_count()
    total_ms, result, events = 0.0, None, []
    # Every step is bracketed by its own pair of CUDA events; the host does NOT wait for a step before enqueueing the
    # next (one synchronise after the last), so that a late host thread on one rank -- Python jitter -- is absorbed by
    # the stream queue instead of stalling the other ranks inside that step's all-reduce (with 8 ranks and a host
    # synchronise per step, two such stalls in ten steps cost 5 % of the max-over-ranks time).  An end-to-end step
    # (step_e2e) reads its results on the host and therefore synchronises by itself.
    for _ in range(steps):
      flush.fill_(1)                       # L2 flush between timed iterations (not timed)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      result = fn()
      e1.record()
      events.append((e0, e1))
    if sampler:   # clocks / throttle reasons while the device works through the timed steps; the host only waits here
      while not events[-1][1].query():
        sampler.sample_now()
        time.sleep(0.002)
    torch.cuda.synchronize()
    for e0, e1 in events:
      total_ms += e0.elapsed_time(e1)
      if os.environ.get('NQ_BENCH_DEBUG'):
        print(f'[bench rank {rank}] {fn.__name__}: {e0.elapsed_time(e1):.3f} ms', file=sys.stderr)
    launches = lib.nq_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if ws > 1:
      dist.barrier()
      dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks
    return float(t.item()) / steps, launches, clocks, result

  ms, launches, clocks, (grad, mom) = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
  units = float(B_global) * C2_S
  value = units / (ms * 1e-3)
  if args.no_extras:     # kernel variants (scripts/variant_strict.sh): the device-resident leg only
    if rank == 0:
      emit(dict(metric='langevin_traj_steps_per_s', value=value, ms_per_step=ms, math_mode=args.math,
                n_gpus=ws, mean_work=float(mom[1] / mom[0])))
    if ws > 1:
      dist.barrier()
      dist.destroy_process_group()
    return
  ms_e2e, _, _, grad_e2e = timed(step_e2e, max(3, args.steps // 2), 2)
  e2e_value = units / (ms_e2e * 1e-3)
  other = 'fast' if args.math == 'strict' else 'strict'
  bc.set_math_mode(other)
  ms_other, _, _, (grad_other, mom_other) = timed(step_resident, max(3, args.steps // 2), 2)
  bc.set_math_mode(args.math)

  # ---- configs[4]: 10^7 trajectories x 10^4 steps SHARDED over the ranks (strong scaling), one all-reduce
  c5 = None
  if not args.no_config5:
    B5 = C5_B
    f5, n5 = nqd.shard_range(B5, rank, ws)
    seeds5 = nqr.split(seed, B5, first=f5, count=n5)
    x5 = bc._stationary_final(pot, np.zeros((1, 1), np.float32), 0.0, shift, nqr.split(eq_key, B5, f5, n5), 1000, DT, KT,
                              GAMMA).reshape(-1).contiguous()

    def step_c5():
      out = bc._simulate(pot, shift, r_dev, phi_dev, x5, seeds5, C2_S, 0, DT, KT, 1.0, GAMMA)
      if ws > 1:
        dist.all_reduce(out['sums_flat'], op=dist.ReduceOp.SUM)
      gs, mom5 = out['grad_sums'], out['moments']
      return (gs[0] + gs[1]) / B5, mom5
    ms5, _, _, (grad5, mom5) = timed(step_c5, 2, 1)
    c5 = dict(workload=f'configs[4]: {B5} trajectories x {C2_S} steps sharded contiguously over {ws} GPU(s) '
                       f'({n5} per rank), positions not recorded, one all-reduce of [2 x 13 gradient sums | 8 moments]',
              scaling='strong', value=float(B5) * C2_S / (ms5 * 1e-3), unit='traj-steps/s', ms_per_step=ms5,
              per_gpu=float(B5) * C2_S / (ms5 * 1e-3) / ws, steps=2, warmup=1,
              check=dict(n=float(mom5[0]), mean_work=float(mom5[1] / mom5[0]), grad0=float(grad5[0])))
    del seeds5, x5

  line = None
  if rank == 0:
    per_gpu = value / ws
    peaks = measured_peaks()
    hbm_peak = peaks['hbm_gbs'] if peaks else 6650.0
    # per-trajectory I/O + trap table + protocol Jacobian + the 2^23-entry normal-draw table (read once per launch when
    # cold -- the bench flushes L2 between timed iterations -- and L2-resident from then on)
    alg_bytes = 24.0 * C2_B + 4.0 * (C2_S + 1) + 4.0 * 13 * (C2_S - 1) + 4.0 * (1 << 23)
    traffic, traffic_src = profiled_dram_traffic(args.math)
    measured = langevin_instr_per_step().get(args.math, {})
    alu_peak = 148 * 64 * 1.965e9   # the ALU pipe (LOP3 / SHF / IADD3: the threefry rotates and xors) is half rate
    roofline = dict(
        bound='issue', kernel='langevin_persistent_kernel<BIOMOLECULE_SHIFTED, D=13> (FIFO-scheduled form of langevin_kernel)',
        achieved=per_gpu * INSTR_PER_TRAJ_STEP / 1e12, peak=ISSUE_PEAK / 1e12, unit='Tinstr/s (thread-instructions)',
        frac=per_gpu * INSTR_PER_TRAJ_STEP / ISSUE_PEAK,
        note='instruction-issue bound (3 threefry2x32-20 blocks per step): 326 thread-instr/traj-step (SURVEY 8d) x '
             'measured traj-steps/s per GPU / (148 SM x 128 lanes x 1.965 GHz); not HBM- or tensor-bound.  The kernel '
             'EXECUTES fewer than the 326 algorithmic instructions (packed f32x2 pairs, the normal draw as a table load): '
             'executed_* are from the committed ncu capture, alu_pipe is the binding half-rate pipe',
        executed_instr_per_step=measured.get('instr_per_step'),
        executed_alu_instr_per_step=measured.get('alu_instr_per_step'),
        alu_pipe=(dict(achieved=per_gpu * measured['alu_instr_per_step'] / 1e12, peak=alu_peak / 1e12,
                       unit='Tinstr/s on the half-rate ALU pipe', frac=per_gpu * measured['alu_instr_per_step'] / alu_peak)
                  if measured.get('alu_instr_per_step') else None),
        executed_source=measured.get('source'),
        traffic=traffic, traffic_source=traffic_src,
        hbm=dict(bound='hbm', achieved=alg_bytes / (ms * 1e-3) / 1e9, peak=hbm_peak, unit='GB/s',
                 frac=alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                 peak_source='MEASURED_PEAKS.json (of measured)' if peaks else 'fallback',
                 algorithmic_bytes_per_launch=alg_bytes))
    line = dict(metric='langevin_traj_steps_per_s', value=value, unit='traj-steps/s', n_gpus=ws, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', config=workload_config(ws), math_mode=args.math,
                per_gpu=per_gpu, target_per_gpu=1e11, clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=e2e_value, unit='traj-steps/s', ms_per_step=ms_e2e, h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=d2h, api='noneq_opt_b200.distributed.langevin_gradient_sharded '
                         '(= barrier_crossing.estimate_gradient_boltzinit per rank)'),
                roofline=roofline,
                check=dict(mean_work=float(mom[1] / mom[0]), grad0=float(grad[0]), grad0_e2e=float(grad_e2e[0])),
                config5=c5,
                other_math_mode={'math_mode': other, 'value': units / (ms_other * 1e-3), 'ms_per_step': ms_other,
                                 'mean_work': float(mom_other[1] / mom_other[0]), 'grad0': float(grad_other[0]),
                                 'note': 'strict = reference operation order for positions, closed-form work increment '
                                         '(parity-certified, default); fast = MUFU/FMA landscape, same exact normal '
                                         'draws, parity statistical (DESIGN.md 4.1, 5)'})

  # ---- secondary: Ising spin-updates/s (rank-local replicas; no collective on the data path)
  ising_res = None if args.no_ising else bench_ising(args, dev, rank, ws)
  if rank == 0:
    line['ising'] = ising_res
    if ws == 1 and not args.no_ising:
      line['config1_training'] = bench_config1_training(dev)
      line['ising_training'] = bench_ising_training(dev)
    if ws == 1 and not args.no_cpu:
      line['cpu_baseline'] = cpu_baseline()
      if not args.no_ising:
        cpu_i = cpu_baseline_ising()
        line['ising']['c3_64x64_4096rep_1000sweeps_grad']['cpu_baseline'] = cpu_i['c3']
        line['ising']['c4_1024x1024_256rep']['cpu_baseline'] = cpu_i['c4']
    emit(line)
  if ws > 1:
    dist.barrier()
    dist.destroy_process_group()


def bench_config1_training(dev):
  """configs[0] (10^3 trajectories x 10^3 steps, spring, the reference's CPU-sized case): a latency-bound
  regime.  One whole optimisation step (key split, equilibration-free start, gradient, Adam) through
  the step-by-step host loop and as one CUDA-graph replay (training.LangevinTrainer)."""
  import torch
  from noneq_opt_b200 import barrier_crossing as bc, optimizers, random as nqr, space, training
  _, shift = space.free()
  B, S, Neq, dt, n = 1000, 1000, 100, 1e-3, 100
  energy = bc.V_spring(1.0)
  coeffs0 = bc.fit_linear_to_cheby(1.0, dt, 5., 10., 12)
  key = np.array([0, 7], np.uint32)
  opt = optimizers.adam(optimizers.exponential_decay(0.1, n, 0.001))
  grad_fn = bc.estimate_gradient_boltzinit(B, energy, 5., 10., Neq, shift, S, dt, 1.0, 1.0, 1.0, record_positions=False)
  x0 = np.full((B, 1, 1), 5., np.float32)

  def host_loop(k, steps):
    state = opt.init_fn(coeffs0)
    for j in range(steps):
      k, _, k2 = nqr.split_host(k, 3)
      grad, _ = grad_fn(opt.params_fn(state), x0, k2)
      state = opt.update_fn(j, grad, state)
    torch.cuda.synchronize()
  host_loop(key, 5)
  t0 = time.perf_counter()
  host_loop(key, n)
  host_ms = 1e3 * (time.perf_counter() - t0) / n
  tr = training.LangevinTrainer(B, energy, coeffs0, 5., 10., Neq, shift, S, dt, 1.0, 1.0, 1.0, key,
                                step_size=0.1, decay_steps=n, decay_rate=0.001)
  tr.run(5)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  e0.record()
  for _ in range(n):
    tr.step()
  e1.record()
  torch.cuda.synchronize()
  wall_ms = 1e3 * (time.perf_counter() - t0) / n
  return dict(workload='spring, 1000 trajectories x (100 + 1000) steps, 13 coefficients, one Adam step per gradient',
              host_loop_ms_per_step=host_ms, graph_replay_ms_per_step=wall_ms,
              graph_replay_device_ms_per_step=e0.elapsed_time(e1) / n, launches_per_step=tr.launches_per_step,
              traj_steps_per_s=B * (S + Neq) / (wall_ms * 1e-3),
              note='latency-bound: 32 CTAs of one producer + one consumer warp (DESIGN.md 4.1, 4.5)')


def bench_ising_training(dev):
  """One Ising training step (`run_ising_opt.py:77-102` around `get_train_step`, `ising.py:263-290`) through the host
  `get_train_step` and as one CUDA-graph replay of `training.IsingTrainer` (zero host synchronisations per step), at
  the reference driver's own shape (its argv comments: size 32, 51 time steps, degree 32, batch 256; a 32 x 32 lattice
  takes the byte-per-spin generic kernel) and at configs[2]'s shape (64^2, 4096 replicas, 1001 time steps, degree 16)."""
  import torch
  from noneq_opt_b200 import ising, optimizers, parameterization as p10n, training
  out = {}
  for name, (L, T, deg, B, n) in dict(reference_driver_32x32_T51_deg32_batch256=(32, 51, 32, 256, 50),
                                      config3_64x64_T1001_deg16_batch4096=(64, 1001, 16, 4096, 5)).items():
    sched = ising.IsingSchedule(
        p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(np.zeros(deg, np.float32)), 0., 0.), ising.log_temp_baseline(0.65)),
        p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(np.zeros(deg, np.float32)), 0., 0.), ising.field_baseline()))
    spins0 = -np.ones((L, L), np.float32)
    opt = optimizers.adam(optimizers.exponential_decay(0.1, n, 0.01))
    step_fn = ising.get_train_step(opt, spins0, B, T, ising.total_entropy_production, 'fwd', max_grad=1.0)
    state, stream = opt.init_fn(sched), ising.seed_stream(0)
    for j in range(2):
      state, _ = step_fn(state, j, next(stream))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(n):
      state, _ = step_fn(state, j, next(stream))
    torch.cuda.synchronize()
    host_ms = 1e3 * (time.perf_counter() - t0) / n
    tr = training.IsingTrainer(sched, spins0, B, T, np.array([0, 0], np.uint32), loss_function=ising.total_entropy_production,
                               max_grad=1.0, step_size=0.1, decay_steps=n, decay_rate=0.01)
    for _ in range(3):
      tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
      tr.step()
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / n
    out[name] = dict(host_get_train_step_ms_per_step=host_ms, graph_replay_ms_per_step=dev_ms,
                     launches_per_step=int(tr.launches_per_step), spin_updates_per_s=float(B) * (T - 1) * L * L / (dev_ms * 1e-3),
                     mean_loss=float(tr.loss.mean().item()))
  return out


def langevin_instr_per_step():
  """Executed thread-instructions per trajectory-step of the headline Langevin kernel per math mode, from the committed
  `ncu --set full --import-source on` captures (profiles/langevin_instr_per_step.json, written from the opcode tables
  scripts/summarize_ncu.py produces)."""
  try:
    return json.load(open(os.path.join(ROOT, 'profiles', 'langevin_instr_per_step.json')))
  except Exception:
    return {}


def ising_instr_per_update():
  """Executed thread-instructions per spin-update of the two Ising kernels, from the committed `ncu --set full`
  summaries (profiles/ising_instr_per_update.json is written next to them by scripts/summarize_ncu.py)."""
  try:
    return json.load(open(os.path.join(ROOT, 'profiles', 'ising_instr_per_update.json')))
  except Exception:
    return {}


def bench_ising(args, dev, rank, ws):
  """Spin-updates/s at configs[2] (with the REINFORCE gradient) and configs[3]; per config: the device-resident
  rate, the end-to-end rate through the public API with HOST buffers, and the issue roofline."""
  import torch
  import torch.distributed as dist
  from noneq_opt_b200 import ising, parameterization as p10n, random as nqr
  out = {}
  measured = ising_instr_per_update()

  def schedule(D, weights):
    return ising.IsingSchedule(
        p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(weights[0]), 0., 0.), ising.log_temp_baseline(0.69, 10, 1)),
        p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(weights[1]), 0., 0.), ising.field_baseline(-1, 1)))

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    torch.cuda.synchronize()
    if ws > 1:
      dist.barrier()
    per_step = []
    for _ in range(steps):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      fn()
      e1.record()
      torch.cuda.synchronize()
      per_step.append(e0.elapsed_time(e1))
    # These legs call the public API, whose host part (schedule and class tables, ~1 ms) sits between the two events:
    # an occasional host stall (inside this process about one call in ten blocks 60-100 ms; the same calls in a process
    # of their own never do, and neither disabling nor freezing the garbage collector changes it) lands in the device
    # interval.  The MEDIAN step is reported, the mean is kept beside it (ms_per_step_mean).
    t = torch.tensor([float(np.median(per_step)), float(np.mean(per_step))], dtype=torch.float64, device=dev)
    if ws > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    timed.last_mean = float(t[1].item())
    return float(t[0].item())

  def roofline(per_gpu, key):
    m = measured.get(key, {})
    return dict(bound='issue', kernel=m.get('kernel', 'ising_fast_kernel'), achieved=per_gpu * INSTR_PER_SPIN_UPDATE / 1e12,
                peak=ISSUE_PEAK / 1e12, unit='Tinstr/s (thread-instructions)', frac=per_gpu * INSTR_PER_SPIN_UPDATE / ISSUE_PEAK,
                algorithmic_instr_per_update=INSTR_PER_SPIN_UPDATE, executed_instr_per_update=m.get('instr_per_update'),
                executed_alu_instr_per_update=m.get('alu_instr_per_update'), source=m.get('source'),
                note='issue-bound (half a threefry2x32-20 block per spin-update); lattice resident in shared memory, '
                     'HBM touched at start / end and for 44 B of summaries per replica-sweep')

  # config 3: 64x64, 4096 replicas/GPU, T = 1001, REINFORCE gradient of total entropy production
  L, R, T = 64, 4096, 1001
  w = (0.1 * np.random.RandomState(0).normal(size=(2, 16))).astype(np.float32)
  sched = schedule(16, w)
  times = np.linspace(0, 1, T, dtype=np.float32)
  spins0 = -torch.ones((L, L), dtype=torch.float32, device=dev)
  seeds = nqr.split(nqr.PRNGKey(0), R * ws, first=rank * R, count=R)
  est = ising.estimate_gradient_rev(ising.total_entropy_production)
  ms = timed(lambda: est(sched, times, spins0, seeds), 10, 3)
  ms_mean = timed.last_mean
  upd = float(R * ws) * (T - 1) * L * L
  # end to end: host lattice + host keys in, host gradient + per-replica loss terms out
  spins0_h = -np.ones((L, L), np.float32)
  seeds_h = seeds.cpu().numpy()
  h2d = spins0_h.nbytes + seeds_h.nbytes + 2 * 16 * 4

  def c3_e2e():
    grad, summ = est(sched, times, spins0_h, seeds_h)
    g = np.concatenate([grad.log_temp.wrapped.wrapped.weights.cpu().numpy().reshape(-1),
                        grad.field.wrapped.wrapped.weights.cpu().numpy().reshape(-1)])
    return g, summ.entropy_production.sum(-1).cpu().numpy()
  g3, ep3 = c3_e2e()
  d2h = g3.nbytes + ep3.nbytes
  ms_e = timed(c3_e2e, 10, 2)
  out['c3_64x64_4096rep_1000sweeps_grad'] = dict(
      spin_updates_per_s=upd / (ms * 1e-3), ms_per_step=ms, ms_per_step_mean=ms_mean, per_gpu=upd / ws / (ms * 1e-3),
      steps=10, warmup=3, statistic='median of the timed steps (the mean is ms_per_step_mean)',
      e2e=dict(value=upd / (ms_e * 1e-3), unit='spin-updates/s', ms_per_step=ms_e, h2d_bytes_per_step=h2d,
               d2h_bytes_per_step=d2h, api='noneq_opt_b200.ising.estimate_gradient_rev(total_entropy_production) with numpy '
               'lattice and keys in, per-replica gradients and entropy production back on the host'),
      roofline=roofline(upd / ws / (ms * 1e-3), 'c3'), roofline_frac_issue=upd / ws / (ms * 1e-3) * INSTR_PER_SPIN_UPDATE / ISSUE_PEAK)

  # config 4: 1024x1024 bit-packed, 256 replicas/GPU, 1000 of the 10 000 sweeps per step (throughput is per sweep)
  L, R, T = 1024, 256, 1001
  sched0 = schedule(16, np.zeros((2, 16), np.float32))
  times = np.linspace(0, 1, 10001, dtype=np.float32)[:T]
  params = sched0(times)
  spins0 = ising.random_spins([L, L], 0.5, nqr.PRNGKey(1))
  seeds = nqr.split(nqr.PRNGKey(0), R * ws, first=rank * R, count=R)
  log_temp, field, spins, L_, batched, seeds_t, broadcast, bits = ising._prepare(params, spins0, seeds)
  ms = timed(lambda: ising._run_kernel(log_temp, field, bits, broadcast, seeds_t, L), 5, 1)
  upd = float(R * ws) * (T - 1) * L * L
  spins0_h = spins0.cpu().numpy()
  seeds_h = seeds.cpu().numpy()

  def c4_e2e():
    final, summ = ising.simulate_ising(params, spins0_h, seeds_h)
    return summ.energy[:, -1].cpu().numpy(), summ.magnetization[:, -1].cpu().numpy()
  e4, m4 = c4_e2e()
  ms_e = timed(c4_e2e, 3, 0)
  peaks = measured_peaks()
  hbm_peak = peaks['hbm_gbs'] if peaks else 6650.0
  out['c4_1024x1024_256rep'] = dict(
      spin_updates_per_s=upd / (ms * 1e-3), ms_per_step=ms, sweeps_per_step=T - 1, per_gpu=upd / ws / (ms * 1e-3),
      steps=5, warmup=1, statistic='median of the timed steps',
      e2e=dict(value=upd / (ms_e * 1e-3), unit='spin-updates/s', ms_per_step=ms_e,
               h2d_bytes_per_step=spins0_h.nbytes + seeds_h.nbytes + 2 * T * 4, d2h_bytes_per_step=e4.nbytes + m4.nbytes,
               api='noneq_opt_b200.ising.simulate_ising with a numpy lattice and keys in (unpacked [R, L, L] final '
                   'lattices stay on the device), final energy and magnetisation back on the host'),
      roofline=roofline(upd / ws / (ms * 1e-3), 'c4'),
      roofline_frac_issue=upd / ws / (ms * 1e-3) * INSTR_PER_SPIN_UPDATE / ISSUE_PEAK,
      hbm_gbs_if_streamed=upd / ws / (ms * 1e-3) * 0.375 / 1e9,
      hbm_frac_if_streamed=upd / ws / (ms * 1e-3) * 0.375 / 1e9 / hbm_peak,
      note='lattice is shared-memory resident for the whole trajectory: HBM is touched '
           'only at start/end (0.375 B/update is the streamed-design figure of SURVEY 8d)')
  out['unit'] = 'spin-updates/s'
  out['target_per_gpu'] = 1e11
  return out


def cpu_baseline_ising():
  """Oracle port (C/OpenMP over replicas) on bounded samples of configs 3 and 4, all host cores."""
  from oracle import c_oracle, ising_ref as I, jaxlike as J, parameterization_ref as P
  cores = c_oracle.use_all_cores()
  out = {}
  # config 3: 64^2, T = 1001, 4 replicas per core
  L, T, R = 64, 1001, 4 * cores
  w = (0.1 * np.random.RandomState(0).normal(size=(2, 16))).astype(np.float32)
  times = np.linspace(0, 1, T, dtype=np.float32)
  lt = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[0]), 0., 0.), P.log_temp_baseline(0.69, 10, 1))(times)
  h = P.AddBaseline(P.ConstrainEndpoints(P.Chebyshev(w[1]), 0., 0.), P.field_baseline(-1, 1))(times)
  seeds = J.split(J.PRNGKey(0), 4096)[:R]
  s0 = -np.ones((L, L), np.int8)
  t0 = time.perf_counter()
  c_oracle.ising(lt, h, s0, seeds)
  dt = time.perf_counter() - t0
  out['c3'] = dict(value=R * (T - 1) * L * L / dt, unit='spin-updates/s', cores=cores, kind='port',
                   sample=f'{R} of 4096 replicas x {T - 1} sweeps of 64^2 (forward + summaries), {dt:.1f} s')
  # config 4: 1024^2, one replica per core, 6 sweeps
  L, T, R = 1024, 7, cores
  times = np.linspace(0, 1, 10001, dtype=np.float32)[:T]
  lt, h = P.log_temp_baseline(0.69, 10, 1)(times), P.field_baseline(-1, 1)(times)
  s0 = I.random_spins((L, L), .5, J.PRNGKey(1)).astype(np.int8)
  seeds = J.split(J.PRNGKey(0), 256)[:R]
  t0 = time.perf_counter()
  c_oracle.ising(lt, h, s0, seeds)
  dt = time.perf_counter() - t0
  out['c4'] = dict(value=R * (T - 1) * L * L / dt, unit='spin-updates/s', cores=cores, kind='port',
                   sample=f'{R} of 256 replicas x {T - 1} of 10000 sweeps of 1024^2, {dt:.1f} s')
  return out


def cpu_baseline():
  """Oracle port (C/OpenMP) on a bounded sample of config 2, timed on this box's host cores."""
  from oracle import c_oracle, jaxlike as J, langevin_ref as L
  c_oracle.use_all_cores()
  pot = L.V_biomolecule_shifted(KAPPA, KAPPA, X_M, 0., K_TRAP, BETA)
  r, phi = L.make_schedule_chebyshev(C2_S, C2_COEFFS, 0., 40.)
  n = min(C2_B, 4096 * c_oracle.max_threads())     # ~10 s of CPU work
  seeds = J.split(J.PRNGKey(0), C2_B)[:n]
  c_oracle.langevin(pot, r, phi, np.zeros(64, np.float32), seeds[:64], C2_S, 0, DT, KT, GAMMA, 1.0)   # warm
  t0 = time.perf_counter()
  c_oracle.langevin(pot, r, phi, np.zeros(n, np.float32), seeds, C2_S, 0, DT, KT, GAMMA, 1.0)
  dt = time.perf_counter() - t0
  return dict(value=n * C2_S / dt, unit='traj-steps/s', cores=c_oracle.max_threads(), kind='port',
              sample=f'{n} of {C2_B} trajectories x {C2_S} steps (config 2 slice), {dt:.1f} s of C/OpenMP oracle port; '
                     'the reference itself (JAX-CPU) is not installable in this image')


def _json_stdout():
  """The one JSON line goes to the REAL stdout; everything libraries print on fd 1 meanwhile (NCCL's version banner,
  for one) is sent to stderr, so that stdout holds the line and nothing else."""
  sys.stdout.flush()
  real = os.fdopen(os.dup(1), 'w')
  os.dup2(2, 1)
  return real


OUT = None


def emit(line):
  print(json.dumps(line), file=OUT or sys.stdout, flush=True)


def main():
  global OUT
  OUT = _json_stdout()
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='native', choices=['native', 'reference'])
  ap.add_argument('--math', default=os.environ.get('NQ_MATH_MODE', 'strict'), choices=['strict', 'fast'])
  ap.add_argument('--no-ising', action='store_true')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--no-config5', action='store_true')
  ap.add_argument('--no-extras', action='store_true', help='device-resident Langevin leg only (kernel variant scans)')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_native(args)


if __name__ == '__main__':
  main()
