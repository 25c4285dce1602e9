import jax.numpy as jnp


def free():
  def displacement(Ra, Rb, **unused):
    return Ra - Rb

  def shift(R, dR, **unused):
    return R + dR
  return displacement, shift


def periodic(side):
  def displacement(Ra, Rb, **unused):
    d = Ra - Rb
    return jnp.array(((d + side * 0.5) % side) - side * 0.5)

  def shift(R, dR, **unused):
    return jnp.array((R + dR) % side)
  return displacement, shift
