"""GPU: the C ABI used from plain C (examples/c_client.c: cudaMalloc'd buffers, no Python, no torch)
gives the numbers the Python host gets for the same trap table, Jacobian and seeds."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_matches_the_python_host(cuda, tmp_path):
  nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not os.path.exists(nvcc):
    pytest.skip('nvcc not available to build the C client')
  exe = str(tmp_path / 'c_client')
  libdir = os.path.join(ROOT, 'noneq_opt_b200')
  subprocess.run([nvcc, os.path.join(ROOT, 'examples', 'c_client.c'), '-I' + os.path.join(ROOT, 'include'), '-L' + libdir,
                  '-lnoneq_b200', '-Xlinker', '-rpath=' + libdir, '-Wno-deprecated-gpu-targets', '-o', exe],
                 check=True, capture_output=True, timeout=300)
  out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
  mean_work = float(re.search(r'mean_work ([-0-9.e+]+)', out).group(1))
  grads = np.array([[float(v) for v in m] for m in re.findall(r'logP-term ([-0-9.e+]+) reparam-term ([-0-9.e+]+)', out)])
  assert 'n 1000 ' in out and grads.shape == (2, 2)

  import torch
  from noneq_opt_b200 import barrier_crossing as bc, random as nqr, space
  B, S, Neq = 1000, 1000, 100
  k = np.arange(1, S, dtype=np.float32)
  t = (k - 1) / np.float32(S - 2)
  a, b = np.float32(5.) + np.float32(5.) / np.float32(S), np.float32(5.) * np.float32(S - 2) / np.float32(S)
  r = np.concatenate([[np.float32(5.)], (a + b * t).astype(np.float32), [np.float32(10.)]]).astype(np.float32)
  phi = np.stack([np.ones_like(t), t], 1).astype(np.float32)
  seeds = nqr.split(np.array([0, 0], np.uint32), B)
  g = bc._simulate(bc.V_spring(1.0), space.free()[1], r, phi, torch.full((B,), 5.0, device='cuda'), seeds, S, Neq,
                   1e-3, 1.0, 1.0, 1.0)
  m = g['moments'].cpu().numpy()
  assert abs(mean_work - m[1] / m[0]) < 1e-8 * abs(mean_work)
  np.testing.assert_allclose(grads.T, g['grad_sums'].cpu().numpy() / B, rtol=1e-8)
  assert abs(mean_work - 25 / np.e) < 0.4          # the linear-protocol anchor of BASELINE.md
