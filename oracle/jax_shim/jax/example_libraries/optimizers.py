"""Only the names the reference touches at import / call time; the optimisers themselves are exercised
through noneq_opt_b200.optimizers, not through this shim."""
import collections

Optimizer = collections.namedtuple('Optimizer', 'init_fn update_fn params_fn')
OptimizerState = collections.namedtuple('OptimizerState', 'packed_state tree_def subtree_defs')


def clip_grads(grad_tree, max_norm):
  raise NotImplementedError


def adam(*a, **k):
  raise NotImplementedError


def exponential_decay(*a, **k):
  raise NotImplementedError
