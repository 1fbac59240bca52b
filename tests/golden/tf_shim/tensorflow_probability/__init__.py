"""Stand-in for the two tensorflow_probability distributions the reference's particle-Gibbs sampler uses
(`base_model.py:29-75`): Normal(loc, scale).sample() and Categorical(logits).sample(n).  Draws come from the tensorflow
shim's generator and are logged (tf.NOISE_LOG / tf.UNIFORM_LOG) so that a test can replay them; Categorical uses the
inverse CDF of softmax(logits) at a uniform draw (TF's own sampler cannot be reproduced without TF: only the
distribution is specified).  Test infrastructure only."""
import types

import torch

import tensorflow as tf


class _Normal:
    def __init__(self, loc, scale=1.0):
        self.loc, self.scale = tf._t(loc), scale

    def sample(self, n=None):
        shape = tuple(self.loc.shape) if n is None else (int(n),) + tuple(self.loc.shape)
        return self.loc + self.scale * tf.random.normal(shape, dtype=self.loc.dtype)


class _Categorical:
    def __init__(self, logits):
        self.logits = tf._t(logits).detach()

    def sample(self, n=None):
        k = 1 if n is None else int(n)
        u = torch.rand((k,), dtype=torch.float64, generator=tf._RNG)
        tf.UNIFORM_LOG.append(u.clone())
        cdf = torch.cumsum(torch.softmax(self.logits, dim=0), dim=0)
        idx = torch.clamp(torch.searchsorted(cdf, u), max=self.logits.shape[0] - 1)
        return idx[0] if n is None else idx


distributions = types.SimpleNamespace(Normal=_Normal, Categorical=_Categorical)
