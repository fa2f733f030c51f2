"""Uniform prior with Gaussian tails (differentiable everywhere) -- host-side, O(p).

Mirrors priors/mollified_uniform.py:24-93: log-density = N(0, tail_sigma).log_prob(distance outside
[a, b]) - log(1 + (b - a) / (sqrt(2 pi) tail_sigma)); samples are plain Uniform[a, b).
"""
import math
from numbers import Number

import torch
from torch.distributions import Normal, Uniform, constraints
from torch.distributions.utils import broadcast_all

from .._compat import Prior


class MollifiedUniformPrior(Prior, torch.distributions.Distribution):
    arg_constraints = {"a": constraints.real, "b": constraints.real, "tail_sigma": constraints.positive}
    support = constraints.real
    has_rsample = True

    def __init__(self, a, b, tail_sigma=0.1):
        self.a, self.b, self.tail_sigma = broadcast_all(
            *[torch.as_tensor(v, dtype=torch.float64) for v in (a, b, tail_sigma)])
        batch_shape = torch.Size() if isinstance(a, Number) or isinstance(b, Number) else self.a.size()
        torch.distributions.Distribution.__init__(self, batch_shape, validate_args=False)

    @property
    def mean(self):
        return (self.a + self.b) / 2

    @property
    def _half_range(self):
        return (self.b - self.a) / 2

    @property
    def _log_normalization_constant(self):
        return -torch.log(1 + (self.b - self.a) / (math.sqrt(2 * math.pi) * self.tail_sigma))

    def log_prob(self, X):
        outside = ((X - self.mean).abs() - self._half_range).clamp(min=0)
        tails = Normal(loc=torch.zeros_like(self.a), scale=self.tail_sigma)
        return tails.log_prob(outside) + self._log_normalization_constant

    def rsample(self, sample_shape=torch.Size([])):
        # drawn in float32 like the reference (its a / b are float32 tensors, priors/mollified_uniform.py:59,88):
        # torch.rand consumes the generator differently per dtype, and restart points must be reproducible
        return Uniform(self.a.float(), self.b.float()).rsample(sample_shape).to(self.a)

    def expand(self, expand_shape, _instance=None):
        shape = torch.Size(expand_shape)
        return MollifiedUniformPrior(self.a.expand(shape), self.b.expand(shape), self.tail_sigma.expand(shape))
