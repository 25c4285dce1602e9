"""distrax.Bernoulli as the reference uses it (logits= or probs=; sample with a seed; log_prob)."""
import numpy as np
import jax.numpy as jnp
import jax.random
import torch


class Bernoulli:
  def __init__(self, logits=None, probs=None, dtype=int):
    self._logits = None if logits is None else jnp.array(logits)
    self._probs = None if probs is None else jnp.array(probs, np.float32)

  @property
  def probs(self):
    return self._probs if self._logits is None else jnp.array(torch.sigmoid(self._logits))

  def sample(self, seed=None, sample_shape=()):
    probs = self.probs
    sample_shape = tuple(sample_shape) if hasattr(sample_shape, '__iter__') else (sample_shape,)
    n = int(np.prod(sample_shape)) if sample_shape else 1
    uniform = jax.random.uniform(seed, (n,) + tuple(probs.shape))
    draws = jnp.array((uniform < probs).to(torch.int32))
    return draws.reshape(sample_shape + tuple(probs.shape))

  def log_prob(self, value):
    value = jnp.array(value)
    if self._logits is not None:       # -softplus(logits), -softplus(-logits) with softplus = logaddexp(x, 0)
      zero = torch.zeros_like(self._logits)
      lp0, lp1 = -torch.logaddexp(self._logits, zero), -torch.logaddexp(-self._logits, zero)
    else:
      lp0, lp1 = torch.log1p(-self._probs), torch.log(self._probs)
    v = value.to(lp0.dtype)
    return jnp.array(torch.where(v == 1, torch.zeros_like(lp0), lp0 * (1 - v)) +
                     torch.where(v == 0, torch.zeros_like(lp1), lp1 * v))
