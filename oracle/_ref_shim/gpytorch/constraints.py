"""gpytorch.constraints: Interval / GreaterThan / Positive / LessThan (gpytorch/constraints/constraints.py semantics:
bounds are float32 buffers, ``value = transform(raw) [* (ub - lb)] + lb``)."""
import math

import torch
from torch import sigmoid
from torch.nn import Module
from torch.nn.functional import softplus


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


def inv_sigmoid(x):
    return torch.log(x) - torch.log(1 - x)


TRANSFORM_REGISTRY = {torch.exp: torch.log, torch.nn.functional.softplus: inv_softplus, torch.sigmoid: inv_sigmoid}


def _get_inv_param_transform(param_transform, inv_param_transform=None):
    if inv_param_transform is None:
        inv_param_transform = TRANSFORM_REGISTRY.get(param_transform, None)
        if inv_param_transform is None:
            raise RuntimeError("Must specify inv_param_transform for custom param_transforms")
    return inv_param_transform


class Interval(Module):
    def __init__(self, lower_bound, upper_bound, transform=sigmoid, inv_transform=inv_sigmoid, initial_value=None):
        lower_bound = torch.as_tensor(lower_bound).float()
        upper_bound = torch.as_tensor(upper_bound).float()
        if torch.any(torch.ge(lower_bound, upper_bound)):
            raise RuntimeError("Got parameter bounds with empty intervals.")
        super().__init__()
        self.register_buffer("lower_bound", lower_bound)
        self.register_buffer("upper_bound", upper_bound)
        self._transform = transform
        self._inv_transform = inv_transform
        self._initial_value = initial_value
        if transform is not None and inv_transform is None:
            self._inv_transform = _get_inv_param_transform(transform)

    @property
    def enforced(self):
        return self._transform is not None

    def check(self, tensor):
        return bool(torch.all(tensor <= self.upper_bound) and torch.all(tensor >= self.lower_bound))

    def check_raw(self, tensor):
        t = self.transform(torch.as_tensor(tensor))
        return bool(torch.all(t <= self.upper_bound) and torch.all(t >= self.lower_bound))

    def intersect(self, other):
        if self.transform != other.transform:
            raise RuntimeError("Cant intersect Interval constraints with conflicting transforms!")
        return Interval(torch.max(self.lower_bound, other.lower_bound), torch.min(self.upper_bound, other.upper_bound))

    def transform(self, tensor):
        if not self.enforced:
            return tensor
        return (self._transform(tensor) * (self.upper_bound - self.lower_bound)) + self.lower_bound

    def inverse_transform(self, transformed_tensor):
        if not self.enforced:
            return transformed_tensor
        return self._inv_transform((transformed_tensor - self.lower_bound) / (self.upper_bound - self.lower_bound))

    @property
    def initial_value(self):
        return self._initial_value


class GreaterThan(Interval):
    def __init__(self, lower_bound, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(lower_bound=lower_bound, upper_bound=math.inf, transform=transform,
                         inv_transform=inv_transform, initial_value=initial_value)

    def transform(self, tensor):
        return self._transform(tensor) + self.lower_bound if self.enforced else tensor

    def inverse_transform(self, transformed_tensor):
        return self._inv_transform(transformed_tensor - self.lower_bound) if self.enforced else transformed_tensor


class Positive(GreaterThan):
    def __init__(self, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(lower_bound=0.0, transform=transform, inv_transform=inv_transform,
                         initial_value=initial_value)

    def transform(self, tensor):
        return self._transform(tensor) if self.enforced else tensor

    def inverse_transform(self, transformed_tensor):
        return self._inv_transform(transformed_tensor) if self.enforced else transformed_tensor


class LessThan(Interval):
    def __init__(self, upper_bound, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(lower_bound=-math.inf, upper_bound=upper_bound, transform=transform,
                         inv_transform=inv_transform, initial_value=initial_value)

    def transform(self, tensor):
        return -self._transform(-tensor) + self.upper_bound if self.enforced else tensor

    def inverse_transform(self, transformed_tensor):
        return -self._inv_transform(-(transformed_tensor - self.upper_bound)) if self.enforced else transformed_tensor
