"""gpytorch.kernels (<= 1.9, pre-linear_operator): Kernel base with ``active_dims`` / ARD lengthscale /
``covar_dist``, RBFKernel, MaternKernel, ScaleKernel, ProductKernel, AdditiveKernel, LinearKernel.

Semantics restated from gpytorch/kernels/kernel.py:
* ``Kernel.__call__`` slices ``active_dims`` with ``index_select(-1, ...)``, promotes 1-D inputs, and returns a
  lazily evaluated kernel tensor whose ``.evaluate()`` calls ``forward``; ``.to(dtype=...)`` on that object casts
  the inputs AND the kernel module (in place), which is how a default-dtype GP+ model ends up with a float32 K.
* ``Distance._sq_dist``: subtract the column mean of x1 from both inputs, one matmul of width d+2, diagonal zeroed
  only when ``x1 is x2``-equal and neither requires grad, ``clamp_min_(0)``; ``_dist`` = ``clamp_min_(1e-30).sqrt_()``.
* ``RBFKernel.forward``: inputs divided by the lengthscale, ``exp(-0.5 * sq_dist)``.
* ``MaternKernel.forward``: inputs centred by the mean of x1 and divided by the lengthscale, distance,
  ``exp(-sqrt(2 nu) d) * poly(d)``.
* ``ScaleKernel.forward``: ``base * outputscale``; ``ProductKernel.forward``: product of the sub-kernel matrices.
"""
import math
import warnings
from abc import abstractmethod
from copy import deepcopy

import torch
from torch.nn import ModuleList

from . import settings
from .constraints import Positive
from .lazy import LazyTensor, NonLazyTensor, delazify, lazify
from .module import Module


def default_postprocess_script(x):
    return x


class Distance(torch.nn.Module):
    def __init__(self, postprocess_script=default_postprocess_script):
        super().__init__()
        self._postprocess = postprocess_script

    def _sq_dist(self, x1, x2, postprocess, x1_eq_x2=False):
        adjustment = x1.mean(-2, keepdim=True)
        x1 = x1 - adjustment
        x2 = x2 - adjustment  # x1 and x2 should be identical in all dims except -2 at this point

        # Compute squared distance matrix using quadratic expansion
        x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
        x1_pad = torch.ones_like(x1_norm)
        if x1_eq_x2 and not x1.requires_grad and not x2.requires_grad:
            x2_norm, x2_pad = x1_norm, x1_pad
        else:
            x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
            x2_pad = torch.ones_like(x2_norm)
        x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
        x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
        res = x1_.matmul(x2_.transpose(-2, -1))

        if x1_eq_x2 and not x1.requires_grad and not x2.requires_grad:
            res.diagonal(dim1=-2, dim2=-1).fill_(0)

        # Zero out negative values
        res.clamp_min_(0)
        return self._postprocess(res) if postprocess else res

    def _dist(self, x1, x2, postprocess, x1_eq_x2=False):
        # Need to set postprocess to false here, otherwise the distance is postprocessed twice
        res = self._sq_dist(x1, x2, postprocess=False, x1_eq_x2=x1_eq_x2)
        res = res.clamp_min_(1e-30).sqrt_()
        return self._postprocess(res) if postprocess else res


class LazyEvaluatedKernelTensor(LazyTensor):
    """Result of ``Kernel.__call__``: evaluates ``kernel.forward(x1, x2)`` on demand."""

    def __init__(self, x1, x2, kernel, last_dim_is_batch=False, **params):
        self.x1 = x1
        self.x2 = x2
        self.kernel = kernel
        self.last_dim_is_batch = last_dim_is_batch
        self.params = params
        self._cached = None

    @property
    def tensor(self):
        return self.evaluate()

    @property
    def shape(self):
        return torch.Size([*self.x1.shape[:-2], self.x1.shape[-2], self.x2.shape[-2]])

    def size(self, *a):
        return self.shape[a[0]] if a else self.shape

    def dim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return self.kernel.dtype

    @property
    def device(self):
        return self.x1.device

    def evaluate_kernel(self):
        if self._cached is None:
            # the inputs were already restricted to active_dims by Kernel.__call__
            with settings.lazily_evaluate_kernels(False):
                temp_active_dims = self.kernel.active_dims
                self.kernel.active_dims = None
                res = self.kernel(self.x1, self.x2, diag=False, last_dim_is_batch=self.last_dim_is_batch,
                                  **self.params)
                self.kernel.active_dims = temp_active_dims
            self._cached = lazify(res)
        return self._cached

    def evaluate(self):
        return self.evaluate_kernel().evaluate()

    def diag(self):
        return self.evaluate().diagonal(dim1=-2, dim2=-1)

    def to(self, *args, **kwargs):
        """LazyTensor.to casts every tensor argument and every Module argument -- the kernel module is cast IN
        PLACE (nn.Module.to), which is what turns a float64 request into ... whatever dtype is asked for."""
        x1 = self.x1.to(*args, **kwargs)
        x2 = self.x2.to(*args, **kwargs)
        kernel = self.kernel.to(*args, **kwargs)
        return LazyEvaluatedKernelTensor(x1, x2, kernel, self.last_dim_is_batch, **self.params)

    def __getitem__(self, idx):
        return NonLazyTensor(self.evaluate()[idx])


class Kernel(Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super(Kernel, self).__init__()
        self._batch_shape = batch_shape
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.eps = eps

        param_transform = kwargs.get("param_transform")
        if lengthscale_constraint is None:
            lengthscale_constraint = Positive()
        if param_transform is not None:
            warnings.warn("The 'param_transform' argument is now deprecated.", DeprecationWarning)

        if self.has_lengthscale:
            lengthscale_num_dims = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter(
                name="raw_lengthscale",
                parameter=torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, lengthscale_num_dims)),
            )
            if lengthscale_prior is not None:
                self.register_prior("lengthscale_prior", lengthscale_prior, lambda m: m.lengthscale,
                                    lambda m, v: m._set_lengthscale(v))
            self.register_constraint("raw_lengthscale", lengthscale_constraint)

        self.distance_module = None
        # TODO: Remove this on next official PyTorch release.
        self.__pdist_supports_batch = True

    @abstractmethod
    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        raise NotImplementedError()

    @property
    def batch_shape(self):
        kernels = list(self.sub_kernels())
        if len(kernels):
            return torch.broadcast_shapes(self._batch_shape, *[k.batch_shape for k in kernels])
        return self._batch_shape

    @batch_shape.setter
    def batch_shape(self, val):
        self._batch_shape = val

    @property
    def dtype(self):
        if self.has_lengthscale:
            return self.lengthscale.dtype
        for param in self.parameters():
            return param.dtype
        return torch.get_default_dtype()

    @property
    def is_stationary(self):
        return self.has_lengthscale

    @property
    def lengthscale(self):
        if self.has_lengthscale:
            return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)
        return None

    @lengthscale.setter
    def lengthscale(self, value):
        self._set_lengthscale(value)

    def _set_lengthscale(self, value):
        if not self.has_lengthscale:
            raise RuntimeError("Kernel has no lengthscale.")
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_lengthscale)
        self.initialize(raw_lengthscale=self.raw_lengthscale_constraint.inverse_transform(value))

    def local_load_samples(self, samples_dict, memo, prefix):
        pass

    def covar_dist(self, x1, x2, diag=False, last_dim_is_batch=False, square_dist=False,
                   dist_postprocess_func=default_postprocess_script, postprocess=True, **params):
        if last_dim_is_batch:
            x1 = x1.transpose(-1, -2).unsqueeze(-1)
            x2 = x2.transpose(-1, -2).unsqueeze(-1)

        x1_eq_x2 = torch.equal(x1, x2)

        # torch scripts expect tensors
        postprocess = torch.tensor(postprocess)

        res = None

        # Cache the Distance object or else JIT will recompile every time
        if not self.distance_module or self.distance_module._postprocess != dist_postprocess_func:
            self.distance_module = Distance(dist_postprocess_func)

        if diag:
            # Special case the diagonal because we can return all zeros most of the time.
            if x1_eq_x2:
                res = torch.zeros(*x1.shape[:-2], x1.shape[-2], dtype=x1.dtype, device=x1.device)
                if postprocess:
                    res = dist_postprocess_func(res)
                return res
            else:
                res = torch.norm(x1 - x2, p=2, dim=-1)
                if square_dist:
                    res = res.pow(2)
            if postprocess:
                res = dist_postprocess_func(res)
            return res

        elif square_dist:
            res = self.distance_module._sq_dist(x1, x2, postprocess, x1_eq_x2)
        else:
            res = self.distance_module._dist(x1, x2, postprocess, x1_eq_x2)

        return res

    def named_sub_kernels(self):
        for name, module in self.named_modules():
            if module is not self and isinstance(module, Kernel):
                yield name, module

    def sub_kernels(self):
        for _, kernel in self.named_sub_kernels():
            yield kernel

    def num_outputs_per_input(self, x1, x2):
        return 1

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_, x2_ = x1, x2

        # Select the active dimensions
        if self.active_dims is not None:
            x1_ = x1_.index_select(-1, self.active_dims)
            if x2_ is not None:
                x2_ = x2_.index_select(-1, self.active_dims)

        # Give x1_ and x2_ a last dimension, if necessary
        if x1_.ndimension() == 1:
            x1_ = x1_.unsqueeze(1)
        if x2_ is not None:
            if x2_.ndimension() == 1:
                x2_ = x2_.unsqueeze(1)
            if not x1_.size(-1) == x2_.size(-1):
                raise RuntimeError("x1_ and x2_ must have the same number of dimensions!")

        if x2_ is None:
            x2_ = x1_

        # Check that ard_num_dims matches the supplied number of dimensions
        if settings.debug.on():
            if self.ard_num_dims is not None and self.ard_num_dims != x1_.size(-1):
                raise RuntimeError(
                    "Expected the input to have {} dimensionality "
                    "(based on the ard_num_dims argument). Got {}.".format(self.ard_num_dims, x1_.size(-1))
                )

        if diag:
            res = super(Kernel, self).__call__(x1_, x2_, diag=True, last_dim_is_batch=last_dim_is_batch, **params)
            # Did this Kernel eat the diag option?
            # If it does not return a LazyEvaluatedKernelTensor, we can call diag on the output
            if not isinstance(res, LazyEvaluatedKernelTensor):
                if res.dim() == x1_.dim() and res.shape[-2:] == torch.Size((x1_.size(-2), x2_.size(-2))):
                    res = res.diag()
            return res

        else:
            if settings.lazily_evaluate_kernels.on():
                res = LazyEvaluatedKernelTensor(x1_, x2_, kernel=self, last_dim_is_batch=last_dim_is_batch, **params)
            else:
                res = lazify(super(Kernel, self).__call__(x1_, x2_, last_dim_is_batch=last_dim_is_batch, **params))
            return res

    def __add__(self, other):
        kernels = []
        kernels += self.kernels if isinstance(self, AdditiveKernel) else [self]
        kernels += other.kernels if isinstance(other, AdditiveKernel) else [other]
        return AdditiveKernel(*kernels)

    def __mul__(self, other):
        kernels = []
        kernels += self.kernels if isinstance(self, ProductKernel) else [self]
        kernels += other.kernels if isinstance(other, ProductKernel) else [other]
        return ProductKernel(*kernels)

    def __getstate__(self):
        self.distance_module = None
        return self.__dict__


class AdditiveKernel(Kernel):
    @property
    def is_stationary(self):
        return all(k.is_stationary for k in self.kernels)

    def __init__(self, *kernels):
        super(AdditiveKernel, self).__init__()
        self.kernels = ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        res = 0 if diag else None
        for kern in self.kernels:
            next_term = kern(x1, x2, diag=diag, **params)
            if not diag:
                next_term = lazify(next_term).evaluate()
                res = next_term if res is None else res + next_term
            else:
                res = res + next_term
        return res


class ProductKernel(Kernel):
    @property
    def is_stationary(self):
        return all(k.is_stationary for k in self.kernels)

    def __init__(self, *kernels):
        super(ProductKernel, self).__init__()
        self.kernels = ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        x1_eq_x2 = torch.equal(x1, x2)

        if not x1_eq_x2:
            # If x1 != x2, then we can't make a MulLazyTensor because the kernel won't necessarily be square/symmetric
            res = delazify(self.kernels[0](x1, x2, diag=diag, **params))
        else:
            res = self.kernels[0](x1, x2, diag=diag, **params)
            if not diag:
                res = lazify(res).evaluate()

        for kern in self.kernels[1:]:
            next_term = kern(x1, x2, diag=diag, **params)
            if not x1_eq_x2:
                # Again delazify if x1 != x2
                res = res * delazify(next_term)
            else:
                if not diag:
                    res = res * lazify(next_term).evaluate()
                else:
                    res = res * next_term
        return res


def postprocess_rbf(dist_mat):
    return dist_mat.div_(-2).exp_()


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        # (RBFCovariance, the custom-autograd fast path for the non-ARD / no-grad case, evaluates the same formula)
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        return self.covar_dist(x1_, x2_, square_dist=True, diag=diag, dist_postprocess_func=postprocess_rbf,
                               postprocess=True, **params)


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu=2.5, **kwargs):
        if nu not in {0.5, 1.5, 2.5}:
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super(MaternKernel, self).__init__(**kwargs)
        self.nu = nu

    def forward(self, x1, x2, diag=False, **params):
        mean = x1.reshape(-1, x1.size(-1)).mean(0)[(None,) * (x1.dim() - 1)]

        x1_ = (x1 - mean).div(self.lengthscale)
        x2_ = (x2 - mean).div(self.lengthscale)
        distance = self.covar_dist(x1_, x2_, diag=diag, **params)
        exp_component = torch.exp(-math.sqrt(self.nu * 2) * distance)

        if self.nu == 0.5:
            constant_component = 1
        elif self.nu == 1.5:
            constant_component = (math.sqrt(3) * distance).add(1)
        elif self.nu == 2.5:
            constant_component = (math.sqrt(5) * distance).add(1).add(5.0 / 3.0 * distance ** 2)
        return constant_component * exp_component


class ScaleKernel(Kernel):
    @property
    def is_stationary(self):
        return self.base_kernel.is_stationary

    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super(ScaleKernel, self).__init__(**kwargs)
        if outputscale_constraint is None:
            outputscale_constraint = Positive()

        self.base_kernel = base_kernel
        outputscale = torch.zeros(*self.batch_shape) if len(self.batch_shape) else torch.tensor(0.0)
        self.register_parameter(name="raw_outputscale", parameter=torch.nn.Parameter(outputscale))
        if outputscale_prior is not None:
            self.register_prior("outputscale_prior", outputscale_prior, lambda m: m.outputscale,
                                lambda m, v: m._set_outputscale(v))

        self.register_constraint("raw_outputscale", outputscale_constraint)

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        self._set_outputscale(value)

    def _set_outputscale(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_outputscale)
        self.initialize(raw_outputscale=self.raw_outputscale_constraint.inverse_transform(value))

    def forward(self, x1, x2, last_dim_is_batch=False, diag=False, **params):
        orig_output = self.base_kernel.forward(x1, x2, diag=diag, last_dim_is_batch=last_dim_is_batch, **params)
        outputscales = self.outputscale
        if last_dim_is_batch:
            outputscales = outputscales.unsqueeze(-1)
        if diag:
            outputscales = outputscales.unsqueeze(-1)
            return delazify(orig_output) * outputscales
        else:
            outputscales = outputscales.view(*outputscales.shape, 1, 1)
            return delazify(orig_output) * outputscales if torch.is_tensor(orig_output) else \
                orig_output.evaluate() * outputscales

    def num_outputs_per_input(self, x1, x2):
        return self.base_kernel.num_outputs_per_input(x1, x2)


class LinearKernel(Kernel):
    def __init__(self, num_dimensions=None, offset_prior=None, variance_prior=None, variance_constraint=None,
                 **kwargs):
        super(LinearKernel, self).__init__(**kwargs)
        if variance_constraint is None:
            variance_constraint = Positive()
        self.register_parameter(name="raw_variance",
                                parameter=torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, 1)))
        self.register_constraint("raw_variance", variance_constraint)

    @property
    def variance(self):
        return self.raw_variance_constraint.transform(self.raw_variance)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        x1_ = x1 * self.variance.sqrt()
        x2_ = x2 * self.variance.sqrt()
        return x1_ @ x2_.transpose(-2, -1)
