"""ctypes front-end of the C oracle (oracle/c/noneq_oracle.c).  TEST INFRASTRUCTURE ONLY: used by
tests/ to cross-check the numpy oracle and by bench.py's `cpu_baseline` / `--impl reference` arms."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'c', '_build', 'libnoneq_oracle.so')
_lib = None


def build(force=False):
  if force or not os.path.exists(LIB):
    subprocess.run(['make', '-C', os.path.join(HERE, 'c')] + (['-B'] if force else []), check=True,
                   stdout=subprocess.DEVNULL)
  return LIB


def lib():
  global _lib
  if _lib is None:
    _lib = C.CDLL(build())
    _lib.oracle_max_threads.restype = C.c_int
  return _lib


def max_threads() -> int:
  return int(lib().oracle_max_threads())


def use_all_cores() -> int:
  """OpenMP team = every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
  fn = lib().oracle_set_threads
  fn.restype = C.c_int
  return int(fn(C.c_int(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else 0)))


def _p(a, t):
  return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def langevin(pot, r, phi, x0, seeds, S, Neq, dt, kT, gamma, mass, shift='free', record=False):
  """pot: oracle.langevin_ref.Potential.  Returns dict(work, logp, grad_sums [2, D], positions)."""
  kind = {'spring': 0, 'biomolecule': 1, 'biomolecule_shifted': 2}[pot.kind]
  n = len(x0)
  r = np.ascontiguousarray(r, np.float32)
  x0 = np.ascontiguousarray(x0, np.float32)
  seeds = np.ascontiguousarray(seeds, np.uint32)
  D = 0 if phi is None else phi.shape[1]
  phi_c = None if phi is None else np.ascontiguousarray(phi, np.float32)
  work = np.zeros(n, np.float32)
  logp = np.zeros(n, np.float32)
  grad = np.zeros((2, max(D, 1)), np.float64)
  pos = np.zeros((n, S + 1), np.float32) if record else None
  periodic, box = (0, 0.0) if shift in (None, 'free') else (1, float(shift[1]))
  lib().oracle_langevin(C.c_int(kind), C.c_int(periodic), C.c_double(pot.k_s), C.c_double(pot.kappa_l),
                        C.c_double(pot.kappa_r), C.c_double(pot.x_m), C.c_double(pot.delta_E), C.c_double(pot.beta),
                        C.c_double(dt), C.c_double(kT), C.c_double(gamma), C.c_double(mass), C.c_double(box),
                        C.c_long(S), C.c_long(Neq), _p(r, C.c_float), _p(phi_c, C.c_float), C.c_int(D),
                        _p(x0, C.c_float), _p(seeds, C.c_uint32), C.c_long(n), _p(work, C.c_float),
                        _p(logp, C.c_float), _p(grad, C.c_double), _p(pos, C.c_float))
  return dict(work=work, logp=logp, grad_sums=grad[:, :D], positions=pos)


def ising(log_temp, field, spins0, seeds, snapshots=None):
  """spins0 [L, L] (shared) or [R, L, L]; seeds [R, 2].  Returns (final int8 [R, L, L], summary [R, 7, T-1]);
  with snapshots = ascending sweep indices t (1 .. T-1) also the lattices after those sweeps, int8 [R, len, L, L]."""
  log_temp = np.ascontiguousarray(log_temp, np.float32)
  field = np.ascontiguousarray(field, np.float32)
  seeds = np.ascontiguousarray(seeds, np.uint32).reshape(-1, 2)
  R, T = seeds.shape[0], len(field)
  s0 = np.ascontiguousarray(spins0, np.int8)
  L = s0.shape[-1]
  final = np.zeros((R, L, L), np.int8)
  summary = np.zeros((R, 7, T - 1), np.float32)
  snap_t = None if snapshots is None else np.ascontiguousarray(snapshots, np.int32)
  snaps = None if snapshots is None else np.zeros((R, len(snap_t), L, L), np.int8)
  lib().oracle_ising(C.c_int(L), C.c_int(T), C.c_long(R), _p(log_temp, C.c_float), _p(field, C.c_float),
                     _p(s0, C.c_int8), C.c_int(int(s0.ndim == 2)), _p(seeds, C.c_uint32), _p(final, C.c_int8),
                     _p(summary, C.c_float), C.c_int(0 if snap_t is None else len(snap_t)), _p(snap_t, C.c_int),
                     _p(snaps, C.c_int8))
  return (final, summary) if snapshots is None else (final, summary, snaps)
