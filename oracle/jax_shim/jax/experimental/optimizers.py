from ..example_libraries.optimizers import *  # noqa: F401,F403  (the pre-0.2.25 import path the reference's tests use)
from ..example_libraries.optimizers import Optimizer, OptimizerState, adam, clip_grads, exponential_decay, l2_norm  # noqa: F401
