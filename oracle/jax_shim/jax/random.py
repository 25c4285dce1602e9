"""jax.random on top of oracle/jaxlike.py (threefry2x32, original counter layout; KAT-pinned)."""
import numpy as np

from oracle import jaxlike as J
from ._core import asarray


def PRNGKey(seed):
  return J.PRNGKey(seed)


def split(key, num=2):
  return J.split(np.asarray(key, np.uint32), num)


def _shape(shape):
  return tuple(int(s) for s in (shape if hasattr(shape, '__iter__') else (shape,)))


def normal(key, shape=(), dtype=None):
  return asarray(J.normal(np.asarray(key, np.uint32), _shape(shape)))


def uniform(key, shape=(), dtype=None, minval=0., maxval=1.):
  return asarray(J.uniform(np.asarray(key, np.uint32), _shape(shape), minval, maxval))


def bits(key, shape=(), dtype=None):
  return J.random_bits(np.asarray(key, np.uint32), _shape(shape))
