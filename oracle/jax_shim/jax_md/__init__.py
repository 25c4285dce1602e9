"""The handful of jax_md helpers the reference's integrator calls (space, quantity, interpolate, util)."""
from . import space, quantity, interpolate, util, energy, simulate, minimize, smap  # noqa: F401
