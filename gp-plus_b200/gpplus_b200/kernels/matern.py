"""Matern kernels with nu fixed (kernels/matern.py:4-8)."""
from .._compat import MaternKernel


class Matern32Kernel(MaternKernel):
    def __init__(self, **kwargs):
        kwargs.pop("nu", None)
        super().__init__(nu=1.5, **kwargs)


class Matern52Kernel(MaternKernel):
    def __init__(self, **kwargs):
        kwargs.pop("nu", None)
        super().__init__(nu=2.5, **kwargs)
