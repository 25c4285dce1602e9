"""tfd.Normal as the reference uses it (sample with a seed, log_prob)."""
import numpy as np
import jax.numpy as jnp
import jax.random
import torch


class _Normal:
  def __init__(self, loc, scale):
    self.loc, self.scale = jnp.array(loc), jnp.array(scale)

  def sample(self, seed=None):
    shape = torch.broadcast_shapes(self.loc.shape, self.scale.shape)
    # tfp draws normal(shape=[n=1] + batch_shape) and drops the leading axis: same flat stream
    sampled = jax.random.normal(seed, (1,) + tuple(shape))[0]
    return sampled * self.scale + self.loc

  def log_prob(self, x):
    log_unnormalized = -0.5 * jnp.square(x / self.scale - self.loc / self.scale)
    log_normalization = np.float32(0.5 * np.log(2. * np.pi)) + jnp.log(self.scale)
    return log_unnormalized - log_normalization


class distributions:
  Normal = _Normal
