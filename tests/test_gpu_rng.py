"""GPU: the threefry kernels reproduce jax.random's streams (oracle/jaxlike.py) bit for bit."""
import json
import os

import numpy as np
import pytest

from oracle import jaxlike as J

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_split_bits_uniform_bit_exact(cuda):
  from noneq_opt_b200 import random as nqr
  for seed, n in [(0, 2), (1701, 1001), (42, 7), ((3 << 32) + 9, 4096)]:
    key = J.PRNGKey(seed)
    np.testing.assert_array_equal(nqr.keys_to_numpy(nqr.split(key, n)), J.split(key, n))
  for size in (1, 2, 3, 777, 4096, 100001):
    key = J.PRNGKey(size)
    np.testing.assert_array_equal(nqr.bits(key, (size,)).cpu().numpy().view(np.uint32), J.random_bits(key, (size,)))
    np.testing.assert_array_equal(nqr.uniform(key, (size,)).cpu().numpy(), J.uniform(key, (size,)))
  np.testing.assert_array_equal(nqr.uniform(J.PRNGKey(1), (64, 64)).cpu().numpy(), J.uniform(J.PRNGKey(1), (64, 64)))


def test_split_slices_are_slices_of_the_global_split(cuda):
  from noneq_opt_b200 import random as nqr
  key = J.PRNGKey(5)
  full = J.split(key, 1000)
  for first, count in [(0, 10), (125, 125), (990, 10), (499, 2)]:
    np.testing.assert_array_equal(nqr.keys_to_numpy(nqr.split(key, 1000, first=first, count=count)),
                                  full[first:first + count])


def test_normal_within_one_ulp_of_oracle(cuda):
  from noneq_opt_b200 import random as nqr
  key = J.PRNGKey(1701)
  g = nqr.normal(key, (200001,)).cpu().numpy()
  o = J.normal(key, (200001,))
  # same uniforms and polynomial; only log1pf may differ in its last bit (tolerance: 4 ulp)
  assert np.max(np.abs(g - o) / np.maximum(np.abs(o), 1e-30)) < 5e-7
  assert (g == o).mean() > 0.98
  assert nqr.normal(J.PRNGKey(0), (1,)).cpu().numpy()[0] == np.float32(-0.20584226)


def test_golden_rng_fixture(cuda):
  from noneq_opt_b200 import random as nqr
  g = json.load(open(os.path.join(GOLD, 'rng.json')))
  key = np.array(g['key'], np.uint32)
  np.testing.assert_array_equal(nqr.keys_to_numpy(nqr.split(key, 7)), np.array(g['split7'], np.uint32))
  np.testing.assert_array_equal(nqr.bits(key, (9,)).cpu().numpy().view(np.uint32), np.array(g['bits9'], np.uint32))
  np.testing.assert_array_equal(nqr.uniform(key, (9,)).cpu().numpy().view(np.uint32),
                                np.array(g['uniform9_bits'], np.uint32))


def test_shard_keys_and_device_resident_key_entry_points(cuda):
  """nq_langevin_keys (a rank's rows of the global split) and nq_random_split_dev (key read from device
  memory: the training loop's `key, a, b = split(key, 3)` without a host round trip)."""
  import ctypes as C
  import torch
  from noneq_opt_b200 import _lib, random as nqr
  lib = _lib.lib()
  key = J.PRNGKey(77)
  full = J.split(key, 5000)
  out = torch.empty((300, 2), dtype=torch.int32, device='cuda')
  _lib.check(lib.nq_langevin_keys(_lib.key_arg(key), 5000, 1200, 300, _lib.ptr(out), _lib.stream_arg()))
  np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), full[1200:1500])
  key_dev = nqr.as_key_tensor(key, torch.device('cuda')).reshape(2).clone()
  out3 = torch.empty((3, 2), dtype=torch.int32, device='cuda')
  _lib.check(lib.nq_random_split_dev(_lib.ptr(key_dev), 3, 0, 3, _lib.ptr(out3), _lib.stream_arg()))
  np.testing.assert_array_equal(out3.cpu().numpy().view(np.uint32), J.split(key, 3))
  _lib.check(lib.nq_random_split_dev(_lib.ptr(out3[2]), 5000, 4990, 10, _lib.ptr(out[:10]), _lib.stream_arg()))
  np.testing.assert_array_equal(out[:10].cpu().numpy().view(np.uint32), J.split(J.split(key, 3)[2], 5000)[4990:])
  assert lib.nq_random_split_dev(_lib.ptr(key_dev), 3, 0, 1, _lib.ptr(key_dev), _lib.stream_arg()) == -1   # aliasing
