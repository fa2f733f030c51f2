"""gpytorch.likelihoods (<= 1.9): Gaussian likelihoods; ``likelihood(f_dist)`` is the marginal (covariance + noise)."""
import torch

from ..distributions import MultivariateNormal
from ..module import Module
from . import noise_models
from .noise_models import HomoskedasticNoise, _HomoskedasticNoiseBase  # noqa: F401


class _Likelihood(Module):
    def __init__(self, max_plate_nesting=1):
        super().__init__()
        self.max_plate_nesting = max_plate_nesting

    def marginal(self, function_dist, *args, **kwargs):
        raise NotImplementedError

    def __call__(self, input, *args, **kwargs):
        # Conditional
        if torch.is_tensor(input):
            return super().__call__(input, *args, **kwargs)
        # Marginal
        elif isinstance(input, MultivariateNormal):
            return self.marginal(input, *args, **kwargs)
        else:
            raise RuntimeError("Likelihoods expects a MultivariateNormal input to make marginal predictions, or a "
                               "torch.Tensor for conditional predictions. Got a {}".format(input.__class__.__name__))


Likelihood = _Likelihood


class _GaussianLikelihoodBase(Likelihood):
    def __init__(self, noise_covar, **kwargs):
        super().__init__()
        param_transform = kwargs.get("param_transform")
        if param_transform is not None:
            import warnings
            warnings.warn("The 'param_transform' argument is now deprecated.", DeprecationWarning)
        self.noise_covar = noise_covar

    def _shaped_noise_covar(self, base_shape, *params, **kwargs):
        return self.noise_covar(*params, shape=base_shape, **kwargs)

    def forward(self, function_samples, *params, **kwargs):
        noise = self._shaped_noise_covar(function_samples.shape, *params, **kwargs).diag()
        return torch.distributions.Normal(function_samples, noise.sqrt())

    def marginal(self, function_dist, *params, **kwargs):
        mean, covar = function_dist.mean, function_dist.lazy_covariance_matrix
        noise_covar = self._shaped_noise_covar(mean.shape, *params, **kwargs)
        full_covar = covar + noise_covar
        return function_dist.__class__(mean, full_covar)


class GaussianLikelihood(_GaussianLikelihoodBase):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        noise_covar = HomoskedasticNoise(noise_prior=noise_prior, noise_constraint=noise_constraint,
                                         batch_shape=batch_shape)
        super().__init__(noise_covar=noise_covar)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.initialize(noise=value)

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    @raw_noise.setter
    def raw_noise(self, value):
        self.noise_covar.initialize(raw_noise=value)


__all__ = ["Likelihood", "_GaussianLikelihoodBase", "GaussianLikelihood", "noise_models"]
