"""Pins the oracle's RNG restatement (oracle/jaxlike.py) to published known answers."""
import json
import os

import numpy as np

from oracle import jaxlike as J

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_threefry_random123_kats():
  # Random123 kat_vectors: threefry2x32 20 rounds
  assert [int(v) for v in J.threefry2x32(0, 0, 0, 0)] == [0x6b200159, 0x99ba4efe]
  assert [int(v) for v in J.threefry2x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)] == [0x1cb996fc, 0xbb002be7]
  assert [int(v) for v in J.threefry2x32(0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3)] == [0xc4923a9c, 0x483df7a0]


def test_jax_documented_values():
  # jax docs / test-suite values under the original (non-partitionable) threefry layout
  np.testing.assert_array_equal(J.split(J.PRNGKey(0)), [[4146024105, 967050713], [2718843009, 1272950319]])
  np.testing.assert_array_equal(J.random_bits(J.PRNGKey(1701), (3,)), [56197195, 4200222568, 961309823])
  assert J.uniform(J.PRNGKey(0)) == np.float32(0.41845703)
  assert J.normal(J.PRNGKey(0), (1,))[0] == np.float32(-0.20584226)


def test_prngkey_layout():
  np.testing.assert_array_equal(J.PRNGKey(0), [0, 0])
  np.testing.assert_array_equal(J.PRNGKey((5 << 32) + 7), [5, 7])


def test_odd_length_padding_and_pairing():
  key = J.PRNGKey(3)
  bits5 = J.random_bits(key, (5,))
  # odd length: counts padded to 6 -> pairs (0,3),(1,4),(2,pad 0); output truncated to 5
  y0, y1 = J.threefry2x32(key[0], key[1], np.array([0, 1, 2], np.uint32), np.array([3, 4, 0], np.uint32))
  np.testing.assert_array_equal(bits5, np.concatenate([y0, y1])[:5])
  # a single draw is y0 of block (0, 0)
  assert J.random_bits(key, (1,))[0] == J.threefry2x32(key[0], key[1], 0, 0)[0]


def test_split_is_two_blocks():
  key = J.PRNGKey(9)
  a0, a1 = J.threefry2x32(key[0], key[1], 0, 2)
  b0, b1 = J.threefry2x32(key[0], key[1], 1, 3)
  np.testing.assert_array_equal(J.split(key), [[a0, b0], [a1, b1]])


def test_uniform_range_and_integer_compare():
  key = J.PRNGKey(11)
  bits = J.random_bits(key, (20000,))
  u = J.uniform(key, (20000,))
  assert u.min() >= 0 and u.max() < 1
  for p in (np.float32(0.3), np.float32(0.999999), np.float32(1e-7), np.float32(0.5)):
    thr = int(np.ceil(np.float64(p) * 2 ** 23))
    np.testing.assert_array_equal(u < p, (bits >> 9) < thr)


def test_erfinv_matches_scipy_to_polynomial_accuracy():
  import scipy.special as sps
  u = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
  got = J.erf_inv32(u).astype(np.float64)
  ref = sps.erfinv(u.astype(np.float64))
  assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-5   # Giles: ~6e-6 rel


def test_normal_moments():
  x = J.normal(J.PRNGKey(5), (200000,)).astype(np.float64)
  assert abs(x.mean()) < 0.01 and abs(x.std() - 1) < 0.01


def test_golden_rng_fixture():
  g = json.load(open(os.path.join(GOLD, 'rng.json')))
  key = np.array(g['key'], np.uint32)
  np.testing.assert_array_equal(J.split(key, 7), np.array(g['split7'], np.uint32))
  np.testing.assert_array_equal(J.random_bits(key, (9,)), np.array(g['bits9'], np.uint32))
  np.testing.assert_array_equal(J.uniform(key, (9,)).view(np.uint32), np.array(g['uniform9_bits'], np.uint32))
  np.testing.assert_array_equal(J.normal(key, (9,)).view(np.uint32), np.array(g['normal9_bits'], np.uint32))
