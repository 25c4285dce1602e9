"""ctypes binding of libnoneq_b200_dev.so (include/noneq_b200_dev.h): developer tools only -- the issue-rate
microbenchmark and the bit-identity checker of the STRICT main-path math.  Used by scripts/ and tests/."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libnoneq_b200_dev.so')
_P = C.c_void_p
SIGNATURES = {
    'nqdev_microbench': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    'nqdev_math_check': (C.c_int, [C.c_int32, C.c_uint64, C.c_uint64, C.c_float, _P, _P]),
}
_lib = None


def lib() -> C.CDLL:
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -m noneq_opt_b200.build`')
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(handle, name)
      fn.restype = res
      fn.argtypes = args
    _lib = handle
  return _lib
