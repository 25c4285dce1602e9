"""The two jax_md.space factories the reference's drivers use
(`displacement_fn, shift_fn = space.free()`, run_barrier_crossing_opt.py:102), as kernel
descriptors: `free()` -> R + dR, `periodic(box)` -> mod(R + dR, box)  (SURVEY Appendix A.8).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import _lib


@dataclasses.dataclass(frozen=True)
class Shift:
  kind: int
  box: float = 0.0

  def __call__(self, R, dR, **unused_kwargs):
    out = np.asarray(R) + np.asarray(dR)
    return np.mod(out, self.box) if self.kind == _lib.NQ_SHIFT_PERIODIC else out


def _displacement_free(Ra, Rb, **unused_kwargs):
  return np.asarray(Ra) - np.asarray(Rb)


def free():
  return _displacement_free, Shift(_lib.NQ_SHIFT_FREE)


def periodic(box):
  box = float(box)

  def displacement(Ra, Rb, **unused_kwargs):
    d = np.asarray(Ra) - np.asarray(Rb)
    return np.mod(d + box * 0.5, box) - box * 0.5

  return displacement, Shift(_lib.NQ_SHIFT_PERIODIC, box)


def as_shift(shift) -> Shift:
  if isinstance(shift, Shift):
    return shift
  raise NotImplementedError(
      'shift must come from noneq_opt_b200.space.free() / periodic(box); arbitrary shift '
      'callables cannot be compiled into the sm_100a integrator (no CPU fallback)')
