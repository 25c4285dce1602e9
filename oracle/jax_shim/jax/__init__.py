"""Minimal stand-in for the `jax` package (see ../README.md).  TEST INFRASTRUCTURE ONLY."""
from . import _core
from ._core import (jit, vmap, grad, value_and_grad, checkpoint, tree_map, tree_flatten, tree_unflatten,  # noqa: F401
                    tree_leaves)
from . import numpy, lax, random, tree_util, example_libraries, experimental  # noqa: F401,E402

tree_multimap = tree_map

soft_pmap = vmap


jvp = _core.jvp


class ops:   # `from jax import ops` at the top of simulate.py; nothing on the path calls it
  pass
