"""noneq_opt_b200 -- B200-native (sm_100a) implementation of the ensemble hot path of
mc2engel/noneq_opt: the batched overdamped-Langevin integrator with work / log-prob accumulation
and the checkerboard Glauber Ising sweep, with their gradient estimators, behind the reference's
own Python signatures.  See DESIGN.md; the C ABI is include/noneq_b200.h."""
from . import (barrier_crossing, control_flow, gradients, ising, optimizers, parameterization, potentials,  # noqa: F401
               random, simulate, space, tree)

__all__ = ['barrier_crossing', 'control_flow', 'gradients', 'ising', 'optimizers', 'parameterization',
           'potentials', 'random', 'simulate', 'space', 'tree']
