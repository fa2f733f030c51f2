"""Kernel zoo of the GP+ hot path (kernels/__init__.py:1-6 of the reference).

Every class is a parameter descriptor; the arithmetic is the fused sm_100a covariance kernel
(csrc/cov.cuh), selected by ``family``.
"""
from .._compat import Kernel, MaternKernel, ProductKernel, RBFKernel, ScaleKernel
from .matern import Matern32Kernel, Matern52Kernel
from .Rough_RBF import Rough_RBF
from .wighted_RBF import wighted_RBF
from .wighted_RBF_Z import wighted_RBF_Z

__all__ = ["Kernel", "ScaleKernel", "RBFKernel", "MaternKernel", "ProductKernel", "Matern32Kernel",
           "Matern52Kernel", "Rough_RBF", "wighted_RBF", "wighted_RBF_Z"]
