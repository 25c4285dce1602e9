"""Brownian simulation utilities -- drop-in for `noneq_opt/simulate.py`, which duplicates the
integrator of barrier_crossing.py (`BrownianState` :29-31, `brownian` :33-99,
`run_brownian_stationary_trap` :101-115) and adds the stiffness samplers (:117-130)."""
from __future__ import annotations

from . import random as nqrandom
from .barrier_crossing import (BrownianState, brownian, run_brownian_stationary_trap)  # noqa: F401


def single_brownian_stiffness(energy_fn, init_position, r0_init, shift_fn, sim_length, dt=1e-6, kT=1e-5,
                              mass=1.0, gamma=1.0):
  def _single_brownian_eqm(seed):
    return run_brownian_stationary_trap(energy_fn, init_position, r0_init, shift_fn, seed, sim_length, dt, kT,
                                        gamma)
  return _single_brownian_eqm


def mapped_brownian_stiffness(batch_size, energy_fn, init_position, r0_init, shift_fn, sim_length, dt, kT, mass,
                              gamma):
  """Full trajectories `[batch_size, sim_length, 1, 1]` (unlike mapped_brownian_eqm, the reference
  honours dt/kT/gamma here, simulate.py:124)."""
  single = single_brownian_stiffness(energy_fn, init_position, r0_init, shift_fn, sim_length, dt=dt, kT=kT,
                                     mass=mass, gamma=gamma)

  def _mapped_brownian_eqm(seed):
    return single(nqrandom.split(seed, batch_size))
  return _mapped_brownian_eqm
