from ._core import (register_pytree_node_class, tree_flatten, tree_unflatten, tree_leaves, tree_map)  # noqa: F401
