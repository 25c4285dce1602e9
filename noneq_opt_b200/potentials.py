"""Energy closures of the reference (`noneq_opt/potentials.py`) as kernel descriptors.

`V_spring` potentials.py:47-52, `V_biomolecule` :27-35, `V_biomolecule_shifted` :37-45 (the
latter two are duplicated at barrier_crossing.py:176-194).  The reference hands arbitrary
Python closures to jax, which traces and differentiates them; a CUDA kernel cannot, so each
factory returns an `Energy` object that (i) is still callable as `E(position, r0=...)` for
host-side use (plots, sanity checks) and (ii) carries the `kind` / parameters the sm_100a
integrator is specialised on.  Unknown callables are rejected by the simulators with
NotImplementedError -- there is no CPU fallback.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import _lib


@dataclasses.dataclass(frozen=True)
class Energy:
  kind: int
  k_s: float
  kappa_l: float = 0.0
  kappa_r: float = 0.0
  x_m: float = 0.0
  delta_E: float = 0.0
  beta: float = 1.0

  def __call__(self, position, r0=0.0, **unused_kwargs):
    """Host evaluation in binary64 of the closure body (N=1, dim=1: position[0][0])."""
    x = np.asarray(position, np.float64)[0][0]
    r0 = np.asarray(r0, np.float64)
    es = self.k_s / 2 * (x - r0) ** 2
    if self.kind == _lib.NQ_POT_SPRING:
      return es
    if self.kind == _lib.NQ_POT_BIOMOLECULE:
      left, right = x + self.x_m, x - self.x_m
    else:
      left, right = x, x - 2. * self.x_m
    em = -(1. / self.beta) * np.log(
        np.exp(-0.5 * self.beta * self.kappa_l * left ** 2) +
        np.exp(-(0.5 * self.beta * self.kappa_r * right ** 2 + self.beta * self.delta_E)))
    return em + es

  def fill(self, desc):
    desc.potential = self.kind
    desc.k_s, desc.kappa_l, desc.kappa_r = float(self.k_s), float(self.kappa_l), float(self.kappa_r)
    desc.x_m, desc.delta_E, desc.beta = float(self.x_m), float(self.delta_E), float(self.beta)


def V_spring(k_s):
  return Energy(_lib.NQ_POT_SPRING, k_s)


def V_biomolecule(kappa_l, kappa_r, x_m, delta_E, k_s, beta):
  return Energy(_lib.NQ_POT_BIOMOLECULE, k_s, kappa_l, kappa_r, x_m, delta_E, beta)


def V_biomolecule_shifted(kappa_l, kappa_r, x_m, delta_E, k_s, beta):
  return Energy(_lib.NQ_POT_BIOMOLECULE_SHIFTED, k_s, kappa_l, kappa_r, x_m, delta_E, beta)


def as_energy(energy_fn) -> Energy:
  if isinstance(energy_fn, Energy):
    return energy_fn
  raise NotImplementedError(
      'energy_fn must come from noneq_opt_b200.potentials (V_spring, V_biomolecule, '
      'V_biomolecule_shifted): arbitrary Python closures cannot be compiled into the sm_100a '
      'integrator and there is no CPU fallback')
