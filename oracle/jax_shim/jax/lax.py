from ._core import scan, stop_gradient  # noqa: F401
