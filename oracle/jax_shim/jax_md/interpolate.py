def canonicalize(scalar_or_schedule):
  if callable(scalar_or_schedule):
    return scalar_or_schedule
  return lambda t: scalar_or_schedule
