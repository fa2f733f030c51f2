"""Log-half-horseshoe prior on the raw (log) noise variance -- host-side, O(p).

Mirrors priors/horseshoe.py:52-79 of the reference: density of the half-horseshoe in the original
scale (the spearmint approximation log(1 + 3 (scale/x)^2)) plus the log-Jacobian of x = lb + e^raw;
samples are log(HalfNormal(HalfCauchy(1) * scale)) clamped below at lb.  ``expand`` keeps only the
scale (so expanded priors fall back to the default lb=1e-6), as the reference does (:77-79) --
that is what ``_sample_from_prior`` draws restart points from.
"""
from numbers import Number

import torch
from torch.distributions import HalfCauchy, HalfNormal, constraints
from torch.distributions.utils import broadcast_all

from .._compat import Prior


class LogHalfHorseshoePrior(Prior, torch.distributions.Distribution):
    arg_constraints = {"scale": constraints.positive, "lb": constraints.positive}
    support = constraints.real
    has_rsample = True

    def __init__(self, scale, lb=1e-6, validate_args=None):
        self.scale, self.lb = broadcast_all(*[torch.as_tensor(v, dtype=torch.float64) for v in (scale, lb)])
        batch_shape = torch.Size() if isinstance(scale, Number) else self.scale.size()
        torch.distributions.Distribution.__init__(self, batch_shape, validate_args=validate_args)

    def transform(self, x):
        return self.lb + torch.exp(x)

    def log_prob(self, X):
        ratio = self.scale / self.transform(X)
        return torch.log(torch.log(1 + 3 * ratio ** 2)) + X

    def rsample(self, sample_shape=torch.Size([])):
        shrink = HalfCauchy(1).rsample(self.scale.shape).to(self.lb)
        draw = HalfNormal(shrink * self.scale).rsample(sample_shape).to(self.lb)
        floor = self.lb[0] if len(self.lb.shape) > 0 and len(self.lb) > 1 else self.lb
        draw[draw < floor] = floor
        return draw.log()

    def expand(self, expand_shape, _instance=None):
        return LogHalfHorseshoePrior(self.scale.expand(torch.Size(expand_shape)))
