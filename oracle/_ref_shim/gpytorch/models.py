"""gpytorch.models.ExactGP with the exact (Cholesky) DefaultPredictionStrategy (gpytorch/models/exact_gp.py,
exact_prediction_strategies.py semantics):

* train mode: the inputs must equal ``train_inputs``; returns ``forward(*train_inputs)``;
* eval mode: ``forward`` is called on ``cat([train_inputs, inputs], dim=-2)`` -- so everything ``forward`` does
  per call (GP+ re-ranks categorical columns with ``setlevels``) sees train AND test rows; the joint is split into
  the train / test blocks, ``mean_cache = K_y^-1 (y - m)`` with the likelihood applied to the train block, and
  ``mu* = m* + K*^T mean_cache``, ``Sigma* = K** - K*^T K_y^-1 K*`` through the Cholesky factor.
"""
import warnings
from copy import deepcopy

import torch

from . import settings
from .distributions import MultivariateNormal
from .lazy import NonLazyTensor, delazify, psd_safe_cholesky
from .likelihoods import _GaussianLikelihoodBase
from .module import Module
from .utils.warnings import GPInputWarning


class GP(Module):
    pass


class ExactGP(GP):
    def __init__(self, train_inputs, train_targets, likelihood):
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        if train_inputs is not None and not all(torch.is_tensor(train_input) for train_input in train_inputs):
            raise RuntimeError("Train inputs must be a tensor, or a list/tuple of tensors")
        if not isinstance(likelihood, _GaussianLikelihoodBase):
            raise RuntimeError("ExactGP can only handle Gaussian likelihoods")

        super(ExactGP, self).__init__()
        if train_inputs is not None:
            self.train_inputs = tuple(tri.unsqueeze(-1) if tri.ndimension() == 1 else tri for tri in train_inputs)
            self.train_targets = train_targets
        else:
            self.train_inputs = None
            self.train_targets = None
        self.likelihood = likelihood
        self.prediction_strategy = None

    @property
    def train_targets(self):
        return self._train_targets

    @train_targets.setter
    def train_targets(self, value):
        object.__setattr__(self, "_train_targets", value)

    def _apply(self, fn):
        if self.train_inputs is not None:
            self.train_inputs = tuple(fn(train_input) for train_input in self.train_inputs)
            self.train_targets = fn(self.train_targets)
        return super(ExactGP, self)._apply(fn)

    def local_load_samples(self, samples_dict, memo, prefix):
        pass

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            if torch.is_tensor(inputs):
                inputs = (inputs,)
            inputs = tuple(input_.unsqueeze(-1) if input_.ndimension() == 1 else input_ for input_ in inputs)
            self.train_inputs = inputs
        if targets is not None:
            self.train_targets = targets
        self.prediction_strategy = None

    def train(self, mode=True):
        if mode:
            self.prediction_strategy = None
        return super(ExactGP, self).train(mode)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        self.prediction_strategy = None
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def __call__(self, *args, **kwargs):
        train_inputs = list(self.train_inputs) if self.train_inputs is not None else []
        inputs = [i.unsqueeze(-1) if i.ndimension() == 1 else i for i in args]

        # Training mode: optimizing
        if self.training:
            if self.train_inputs is None:
                raise RuntimeError("train_inputs, train_targets cannot be None in training mode. "
                                   "Call .eval() for prior predictions, or call .set_train_data() to add training data.")
            if settings.debug.on():
                if not all(torch.equal(train_input, input) for train_input, input in zip(train_inputs, inputs)):
                    raise RuntimeError("You must train on the training inputs!")
            res = super().__call__(*inputs, **kwargs)
            return res

        # Prior mode
        elif settings.prior_mode.on() or self.train_inputs is None or self.train_targets is None:
            full_inputs = args
            full_output = super(ExactGP, self).__call__(*full_inputs, **kwargs)
            return full_output

        # Posterior mode
        else:
            if settings.debug.on():
                if all(torch.equal(train_input, input) for train_input, input in zip(train_inputs, inputs)):
                    warnings.warn("The input matches the stored training data. Did you forget to call model.train()?",
                                  GPInputWarning)

            # Get the terms that only depend on training data
            if self.prediction_strategy is None:
                train_output = super().__call__(*train_inputs, **kwargs)
                # Create the prediction strategy for
                self.prediction_strategy = DefaultPredictionStrategy(
                    train_inputs=train_inputs, train_prior_dist=train_output, train_labels=self.train_targets,
                    likelihood=self.likelihood)

            # Concatenate the input to the training input
            full_inputs = []
            batch_shape = train_inputs[0].shape[:-2]
            for train_input, input in zip(train_inputs, inputs):
                # Make sure the batch shapes agree for training/test data
                if batch_shape != train_input.shape[:-2]:
                    batch_shape = torch.broadcast_shapes(batch_shape, train_input.shape[:-2])
                    train_input = train_input.expand(*batch_shape, *train_input.shape[-2:])
                if batch_shape != input.shape[:-2]:
                    batch_shape = torch.broadcast_shapes(batch_shape, input.shape[:-2])
                    train_input = train_input.expand(*batch_shape, *train_input.shape[-2:])
                    input = input.expand(*batch_shape, *input.shape[-2:])
                full_inputs.append(torch.cat([train_input, input], dim=-2))

            # Get the joint distribution for training/test data
            full_output = super(ExactGP, self).__call__(*full_inputs, **kwargs)
            if settings.debug().on():
                if not isinstance(full_output, MultivariateNormal):
                    raise RuntimeError("ExactGP.forward must return a MultivariateNormal")
            full_mean, full_covar = full_output.loc, full_output.lazy_covariance_matrix

            # Determine the shape of the joint distribution
            batch_shape = full_output.batch_shape
            joint_shape = full_output.event_shape
            tasks_shape = joint_shape[1:]  # For multitask learning
            test_shape = torch.Size([joint_shape[0] - self.prediction_strategy.train_shape[0], *tasks_shape])

            # Make the prediction
            predictive_mean, predictive_covar = self.prediction_strategy.exact_prediction(full_mean, full_covar)

            # Reshape predictive mean to match the appropriate event shape
            predictive_mean = predictive_mean.view(*batch_shape, *test_shape).contiguous()
            return full_output.__class__(predictive_mean, predictive_covar)


class DefaultPredictionStrategy(object):
    def __init__(self, train_inputs, train_prior_dist, train_labels, likelihood, root=None, inv_root=None):
        # Get training shape
        self._train_shape = train_prior_dist.event_shape

        # Flatten the training labels
        train_labels = train_labels.reshape(*train_labels.shape[: -len(self.train_shape)], self._train_shape.numel())

        self.train_inputs = train_inputs
        self.train_prior_dist = train_prior_dist
        self.train_labels = train_labels
        self.likelihood = likelihood
        self._last_test_train_covar = None
        mvn = self.likelihood(train_prior_dist, train_inputs)
        self.lik_train_train_covar = mvn.lazy_covariance_matrix
        self._mean_cache = None
        self._chol = None

    @property
    def num_train(self):
        return self._train_shape.numel()

    @property
    def train_shape(self):
        return self._train_shape

    def _cholesky(self):
        if self._chol is None:
            chol = psd_safe_cholesky(delazify(self.lik_train_train_covar))
            self._chol = chol.detach() if settings.detach_test_caches.on() else chol
        return self._chol

    @property
    def mean_cache(self):
        if self._mean_cache is None:
            mvn = self.likelihood(self.train_prior_dist, self.train_inputs)
            train_mean = mvn.loc
            train_labels_offset = (self.train_labels - train_mean).unsqueeze(-1)
            mean_cache = torch.cholesky_solve(train_labels_offset, psd_safe_cholesky(delazify(mvn.lazy_covariance_matrix)))
            mean_cache = mean_cache.squeeze(-1)
            if settings.detach_test_caches.on():
                mean_cache = mean_cache.detach()
            self._mean_cache = mean_cache
        return self._mean_cache

    def exact_prediction(self, joint_mean, joint_covar):
        # Find the components of the distribution that contain test data
        test_mean = joint_mean[..., self.num_train:]
        joint = delazify(joint_covar)
        test_test_covar = joint[..., self.num_train:, self.num_train:]
        test_train_covar = joint[..., self.num_train:, : self.num_train]
        return (self.exact_predictive_mean(test_mean, test_train_covar),
                self.exact_predictive_covar(test_test_covar, test_train_covar))

    def exact_predictive_mean(self, test_mean, test_train_covar):
        res = (test_train_covar @ self.mean_cache.unsqueeze(-1)).squeeze(-1)
        res = res + test_mean
        return res

    def exact_predictive_covar(self, test_test_covar, test_train_covar):
        # exact path (fast_pred_var off; Cholesky solves): K** - K*^T K_y^-1 K*
        L = self._cholesky()
        train_test_covar = test_train_covar.transpose(-1, -2)
        covar_correction_rhs = torch.cholesky_solve(train_test_covar, L)
        return NonLazyTensor(test_test_covar + test_train_covar @ covar_correction_rhs.mul(-1))
