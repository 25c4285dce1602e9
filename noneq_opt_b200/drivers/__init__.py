"""Command-line drivers with the positional arguments and pickle outputs of the reference's
experiment scripts, running on the B200 kernels:

  run_barrier_crossing_opt  <- experiments/barrier_crossing/run_barrier_crossing_opt/run_barrier_crossing_opt.py
  run_barrier_crossing_fwd  <- experiments/barrier_crossing/run_barrier_crossing_fwd/run_barrier_crossing_fwd.py
  run_ising_opt             <- run_ising_opt/run_ising_opt.py
  run_ising_fwd             <- experiments/ising/run_ising_fwd.py

    python -m noneq_opt_b200.drivers.run_ising_opt 32 0 51 32 256 5 1

Like the reference they seed from the wall clock; set NQ_SEED=<int> for reproducible runs.  Arrays
are pickled as numpy arrays (the reference pickles jax arrays, which unpickle to the same shapes).
"""
import os
import pickle
import time

import numpy as np
import torch


def wall_clock_key():
  from .. import random as nqrandom
  return nqrandom.PRNGKey(int(os.environ.get('NQ_SEED', int(time.time()))))


def to_numpy(x):
  from .. import tree
  return tree.tree_map(lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a), x)


def dump(obj, path):
  os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
  with open(path, 'wb') as f:
    pickle.dump(obj, f)
