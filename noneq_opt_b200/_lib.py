"""ctypes binding of libnoneq_b200.so (include/noneq_b200.h).

There is no CPU fallback: if the shared object is missing or a call fails this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NQ_B200_LIB: development override (scripts/build_variants.sh: A/B runs of kernel build variants); never a fallback
LIB_PATH = os.environ.get('NQ_B200_LIB') or os.path.join(HERE, 'libnoneq_b200.so')

NQ_POT_SPRING, NQ_POT_BIOMOLECULE, NQ_POT_BIOMOLECULE_SHIFTED = 0, 1, 2
NQ_SHIFT_FREE, NQ_SHIFT_PERIODIC = 0, 1
NQ_MATH_STRICT, NQ_MATH_FAST = 0, 1
NQ_MOM_COUNT = 8
NQ_ERR_INVALID_ARGUMENT, NQ_ERR_UNSUPPORTED, NQ_ERR_CUDA, NQ_ERR_WORKSPACE = -1, -2, -3, -4
NQ_IS_NSUMMARY, NQ_IS_NPARTIAL = 7, 4
(NQ_IS_WORK, NQ_IS_HEAT, NQ_IS_FWD, NQ_IS_REV, NQ_IS_EP, NQ_IS_MAG, NQ_IS_ENERGY) = range(7)
(NQ_IS_DFWD_DH, NQ_IS_DFWD_DL, NQ_IS_DREV_DH, NQ_IS_DREV_DL) = range(4)


class NqLangevinDesc(C.Structure):
  _fields_ = [('potential', C.c_int32), ('shift_kind', C.c_int32), ('math_mode', C.c_int32),
              ('reserved', C.c_int32),
              ('k_s', C.c_double), ('kappa_l', C.c_double), ('kappa_r', C.c_double),
              ('x_m', C.c_double), ('delta_E', C.c_double), ('beta', C.c_double),
              ('dt', C.c_double), ('kT', C.c_double), ('gamma', C.c_double), ('mass', C.c_double),
              ('box', C.c_double),
              ('steps', C.c_int64), ('eq_steps', C.c_int64)]


class NqIsingDesc(C.Structure):
  _fields_ = [('L', C.c_int32), ('T', C.c_int32), ('R', C.c_int64),
              ('spins0_broadcast', C.c_int32), ('sweep_key_mode', C.c_int32),
              ('ndim', C.c_int32), ('dims', C.c_int32 * 3)]


class NqAdamDesc(C.Structure):
  _fields_ = [(n, C.c_double) for n in ('step_size', 'decay_steps', 'decay_rate', 'b1', 'b2', 'eps', 'max_norm',
                                        'w0', 'w1', 'denom')]


_P = C.c_void_p
_KEY = C.POINTER(C.c_uint32)

# name -> (restype, argtypes); must list EVERY symbol include/noneq_b200.h declares
SIGNATURES = {
    'nq_version': (C.c_int, []),
    'nq_last_error': (C.c_char_p, []),
    'nq_launch_count': (C.c_uint64, []),
    'nq_random_split': (C.c_int, [_KEY, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    'nq_langevin_keys': (C.c_int, [_KEY, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    'nq_random_bits': (C.c_int, [_KEY, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    'nq_random_uniform': (C.c_int, [_KEY, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_float, _P, _P]),
    'nq_random_normal': (C.c_int, [_KEY, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    'nq_langevin_workspace_bytes': (C.c_size_t, [C.c_int64, C.c_int32]),
    'nq_langevin_prepare': (C.c_int, [_P]),
    'nq_langevin_forward': (C.c_int, [C.POINTER(NqLangevinDesc), _P, _P, _P, C.c_int64, _P, _P, _P,
                                      _P, C.c_int64, _P, _P, _P, _P, C.c_size_t, _P]),
    'nq_langevin_grad': (C.c_int, [C.POINTER(NqLangevinDesc), _P, _P, C.c_int32, _P, _P, C.c_int64,
                                   _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'nq_langevin_grad_record': (C.c_int, [C.POINTER(NqLangevinDesc), _P, _P, C.c_int32, _P, _P, C.c_int64,
                                          _P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    'nq_langevin_ckpt_count': (C.c_size_t, [C.c_int64, C.c_int64]),
    'nq_langevin_forward_ckpt': (C.c_int, [C.POINTER(NqLangevinDesc), _P, _P, _P, C.c_int64, C.c_int64,
                                           _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'nq_langevin_replay': (C.c_int, [C.POINTER(NqLangevinDesc), _P, C.c_int64, C.c_int64, _P, _P, _P, _P,
                                     _P, _P, C.c_size_t, _P]),
    'nq_langevin_stationary': (C.c_int, [C.POINTER(NqLangevinDesc), C.c_float, _P, _P, C.c_int64, _P, _P, _P]),
    'nq_langevin_apply': (C.c_int, [C.POINTER(NqLangevinDesc), C.c_float, _P, _P, _P, C.c_int64, _P]),
    'nq_langevin_apply_nd': (C.c_int, [C.POINTER(NqLangevinDesc), C.c_float, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    'nq_ising_words_per_lattice': (C.c_size_t, [C.c_int32]),
    'nq_ising_workspace_bytes': (C.c_size_t, [C.POINTER(NqIsingDesc)]),
    'nq_ising_run': (C.c_int, [C.POINTER(NqIsingDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _P, C.c_size_t, _P]),
    'nq_ising_random_spins': (C.c_int, [_KEY, C.c_int64, C.c_float, _P, _P]),
    'nq_ising_pack': (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    'nq_ising_unpack': (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P]),
    'nq_ising_measure': (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.c_int64, _P, _P]),
    'nq_ising_grad_assemble': (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_float),
                                         C.c_float, _P, _P, _P, _P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P, _P, _P,
                                         _P, _P, _P]),
    'nq_ising_grad_assemble_dev': (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P, _P, _P,
                                             _P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    'nq_schedule_compose': (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P, _P]),
    'nq_ising_class_tables': (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    'nq_ising_grad_mean': (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, _P, _P]),
    'nq_random_split_dev': (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    'nq_schedule_affine': (C.c_int, [_P, _P, _P, C.c_float, C.c_float, C.c_int64, C.c_int32, _P, _P]),
    'nq_adam_step': (C.c_int, [C.POINTER(NqAdamDesc), C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'nq_allreduce_f64': (C.c_int, [_P, _P, C.c_int64, _P]),
}

_lib = None


def lib() -> C.CDLL:
  """The loaded library; raises (never falls back) when it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(
          f'{LIB_PATH} is missing: build it with `python -m noneq_opt_b200.build` '
          '(nvcc, sm_100a). noneq_opt_b200 has no CPU fallback.')
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(handle, name)    # AttributeError if the header and the library disagree
      fn.restype = res
      fn.argtypes = args
    _lib = handle
  return _lib


def check(rc: int):
  if rc != 0:
    msg = lib().nq_last_error().decode('utf-8', 'replace')
    raise RuntimeError(f'libnoneq_b200 error {rc}: {msg}')


def key_arg(key):
  """host uint32[2] -> ctypes array"""
  import numpy as np
  k = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(2))
  return (C.c_uint32 * 2)(int(k[0]), int(k[1]))


def ptr(t):
  """device pointer of a torch tensor (or None)"""
  if t is None:
    return None
  if not t.is_cuda:
    raise RuntimeError('noneq_opt_b200 kernels take CUDA tensors only (no CPU fallback)')
  if not t.is_contiguous():
    raise RuntimeError('tensor must be contiguous')
  return C.c_void_p(t.data_ptr())


def stream_arg():
  import torch
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
  import torch
  if not torch.cuda.is_available():
    raise RuntimeError('noneq_opt_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
  return torch.device('cuda', torch.cuda.current_device())
