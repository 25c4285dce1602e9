import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def built_lib():
  """Path of libnoneq_b200.so, building it (nvcc cross-compiles without a GPU) when absent."""
  from noneq_opt_b200 import _lib, build
  if not os.path.exists(_lib.LIB_PATH):
    build.build()
  return _lib.LIB_PATH


@pytest.fixture(scope='session')
def cuda():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from noneq_opt_b200 import _lib
  _lib.lib()   # raises loudly if the extension is missing: GPU tests never fall back
  return torch.device('cuda')
