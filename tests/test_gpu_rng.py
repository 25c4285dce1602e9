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
