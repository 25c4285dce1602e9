"""nq_allreduce_f64 (include/noneq_b200.h): the one collective of the sharded path, called through the
C ABI on a communicator the caller owns.  With one visible GPU the communicator has a single rank
(the sum is the identity); with two or more, every device of a ncclCommInitAll clique contributes."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nccl():
  try:
    return C.CDLL('libnccl.so.2')
  except OSError as e:  # pragma: no cover
    pytest.skip(f'libnccl.so.2 not loadable: {e}')


def test_allreduce_f64_over_all_visible_devices(cuda):
  from noneq_opt_b200 import _lib
  nccl, lib = _nccl(), _lib.lib()
  ndev = min(torch.cuda.device_count(), 2)
  comms = (C.c_void_p * ndev)()
  devs = (C.c_int * ndev)(*range(ndev))
  nccl.ncclCommInitAll.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]
  assert nccl.ncclCommInitAll(comms, ndev, devs) == 0
  try:
    bufs = [torch.arange(37, dtype=torch.float64, device=f'cuda:{d}') * (d + 1) + 0.125 for d in range(ndev)]
    nccl.ncclGroupStart()
    for d in range(ndev):
      with torch.cuda.device(d):
        _lib.check(lib.nq_allreduce_f64(C.c_void_p(comms[d]), _lib.ptr(bufs[d]), 37, _lib.stream_arg()))
    nccl.ncclGroupEnd()
    want = sum(np.arange(37, dtype=np.float64) * (d + 1) + 0.125 for d in range(ndev))
    for d in range(ndev):
      torch.cuda.synchronize(d)
      np.testing.assert_array_equal(bufs[d].cpu().numpy(), want)
  finally:
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    for d in range(ndev):
      nccl.ncclCommDestroy(C.c_void_p(comms[d]))


def test_allreduce_f64_rejects_null_communicator(cuda):
  from noneq_opt_b200 import _lib
  buf = torch.zeros(4, dtype=torch.float64, device='cuda')
  rc = _lib.lib().nq_allreduce_f64(None, _lib.ptr(buf), 4, _lib.stream_arg())
  assert rc == -1 and b'communicator' in _lib.lib().nq_last_error()
