"""Dense-torch re-statement of the slice of gpytorch (<= 1.9) the GP+ reference touches -- TEST INFRASTRUCTURE.
See oracle/_ref_shim/__init__.py.  Written from the published gpytorch semantics; not
gpytorch itself, not copied from it."""
from . import (constraints, distributions, functions, kernels, lazy, likelihoods, means, metrics, mlls, models, priors,
               settings, utils)
from .lazy import delazify, lazify
from .mlls import ExactMarginalLogLikelihood
from .module import Module

__version__ = "1.8.1+gpplus_b200.shim"

__all__ = ["Module", "ExactMarginalLogLikelihood", "constraints", "distributions", "functions", "kernels", "lazy",
           "likelihoods", "means", "metrics", "mlls", "models", "priors", "settings", "utils", "lazify", "delazify"]
