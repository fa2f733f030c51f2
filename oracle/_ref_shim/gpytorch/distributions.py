"""gpytorch.distributions.MultivariateNormal (gpytorch/distributions/multivariate_normal.py semantics): a torch
MultivariateNormal whose covariance may be a lazy tensor; ``log_prob`` with fast computations off goes through
``scale_tril = psd_safe_cholesky(covariance)``; ``variance`` is clamped below at
``settings.min_variance``."""
import math
import warnings

import torch
from torch.distributions import MultivariateNormal as TMultivariateNormal
from torch.distributions.utils import _standard_normal, lazy_property

from . import settings
from .lazy import LazyTensor, NonLazyTensor, delazify, lazify, psd_safe_cholesky
from .utils.warnings import NumericalWarning


class Distribution(torch.distributions.Distribution):
    pass


class MultivariateNormal(TMultivariateNormal, Distribution):
    def __init__(self, mean, covariance_matrix, validate_args=False):
        self._islazy = isinstance(mean, LazyTensor) or isinstance(covariance_matrix, LazyTensor)
        if self._islazy:
            if validate_args:
                ms = mean.size(-1)
                cs1 = covariance_matrix.size(-1)
                cs2 = covariance_matrix.size(-2)
                if not (ms == cs1 and ms == cs2):
                    raise ValueError(f"Wrong shapes in {self._repr_sizes(mean, covariance_matrix)}")
            self.loc = mean
            self._covar = covariance_matrix
            self.__unbroadcasted_scale_tril = None
            self._validate_args = validate_args
            batch_shape = torch.broadcast_shapes(self.loc.shape[:-1], covariance_matrix.shape[:-2])
            event_shape = self.loc.shape[-1:]
            # TODO: Integrate argument validation for LazyTensors into torch.distribution validation logic
            super(TMultivariateNormal, self).__init__(batch_shape, event_shape, validate_args=False)
        else:
            super().__init__(loc=mean, covariance_matrix=covariance_matrix, validate_args=validate_args)

    @property
    def _unbroadcasted_scale_tril(self):
        if self.islazy and self.__unbroadcasted_scale_tril is None:
            # cache root decomposition
            ust = delazify(self.lazy_covariance_matrix.cholesky())
            self.__unbroadcasted_scale_tril = ust
        return self.__unbroadcasted_scale_tril

    @_unbroadcasted_scale_tril.setter
    def _unbroadcasted_scale_tril(self, ust):
        if self.islazy:
            raise NotImplementedError("Cannot set _unbroadcasted_scale_tril for lazy MVN distributions")
        self.__unbroadcasted_scale_tril = ust

    @property
    def islazy(self):
        return self._islazy

    @property
    def mean(self):
        return self.loc

    @lazy_property
    def covariance_matrix(self):
        if self.islazy:
            return self._covar.evaluate()
        return super().covariance_matrix

    @property
    def lazy_covariance_matrix(self):
        if self.islazy:
            return self._covar
        return lazify(super().covariance_matrix)

    def add_jitter(self, noise=1e-4):
        return self.__class__(self.mean, self.lazy_covariance_matrix.add_jitter(noise))

    def confidence_region(self):
        std2 = self.stddev.mul_(2)
        mean = self.mean
        return mean.sub(std2), mean.add(std2)

    @property
    def stddev(self):
        return self.variance.sqrt()

    @property
    def variance(self):
        if self.islazy:
            # overwrite this since torch MVN uses unbroadcasted_scale_tril for this
            diag = self.lazy_covariance_matrix.diag()
            diag = diag.view(diag.shape[:-1] + self._event_shape)
            variance = diag.expand(self._batch_shape + self._event_shape)
        else:
            variance = super().variance
        # Check to make sure that variance isn't lower than minimum allowed value (default 1e-6).
        # This ensures that all variances are positive
        min_variance = settings.min_variance.value(variance.dtype)
        if variance.lt(min_variance).any():
            warnings.warn(
                f"Negative variance values detected. "
                "This is likely due to numerical instabilities. "
                f"Rounding negative variances up to {min_variance}.",
                NumericalWarning,
            )
            variance = variance.clamp_min(min_variance)
        return variance

    def log_prob(self, value):
        if settings.fast_computations.log_prob.off():
            return super().log_prob(value)
        # the stochastic (CG / Lanczos) path of gpytorch is not restated: exact Cholesky in both modes
        if self._validate_args:
            self._validate_sample(value)
        mean, covar = self.loc, self.lazy_covariance_matrix
        diff = value - mean
        inv_quad, logdet = covar.inv_quad_logdet(inv_quad_rhs=diff.unsqueeze(-1), logdet=True)
        res = -0.5 * sum([inv_quad, logdet, diff.size(-1) * math.log(2 * math.pi)])
        return res

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        covar = self.lazy_covariance_matrix
        L = psd_safe_cholesky(covar.evaluate())
        if base_samples is None:
            shape = torch.Size(sample_shape) + self.loc.shape
            base_samples = _standard_normal(shape, dtype=self.loc.dtype, device=self.loc.device)
        return self.loc + (L @ base_samples.unsqueeze(-1)).squeeze(-1)

    def __add__(self, other):
        if isinstance(other, MultivariateNormal):
            return self.__class__(self.mean + other.mean, self.lazy_covariance_matrix + other.lazy_covariance_matrix)
        if isinstance(other, (int, float)):
            return self.__class__(self.mean + other, self.lazy_covariance_matrix)
        raise RuntimeError("Unsupported type {} for addition w/ MultivariateNormal".format(type(other)))

    def __mul__(self, other):
        if not (isinstance(other, int) or isinstance(other, float)):
            raise RuntimeError("Can only multiply by scalars")
        if other == 1:
            return self
        return self.__class__(self.mean * other, self.lazy_covariance_matrix * (other ** 2))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return self.__class__(self.mean[idx], NonLazyTensor(self.covariance_matrix[idx + (slice(None),)][..., idx[-1]]))
