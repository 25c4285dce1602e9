from . import optimizers  # noqa: F401
