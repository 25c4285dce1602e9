import jax

Energy, Force = 'Energy', 'Force'


def force(energy_fn):
  return lambda R, *a, **kw: -jax.grad(energy_fn)(R, *a, **kw)


def canonicalize_force(energy_or_force_fn):
  """jax_md decides at the first call whether it was given an energy (scalar output) or a force."""
  state = {}

  def force_fn(R, **kw):
    if 'fn' not in state:
      out = energy_or_force_fn(R, **kw)
      state['fn'] = force(energy_or_force_fn) if getattr(out, 'shape', ()) == () else energy_or_force_fn
    return state['fn'](R, **kw)
  return force_fn


def canonicalize_mass(mass):
  return mass
