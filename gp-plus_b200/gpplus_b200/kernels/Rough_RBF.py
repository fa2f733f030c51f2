"""Rough RBF: k = exp(-sum_d l_d dx_d^2), the "lengthscale" acting as a precision
(kernels/Rough_RBF.py:27-32: x is multiplied by sqrt(lengthscale) and the squared distance is
exponentiated without the 1/2)."""
from .._compat import FAMILY_EXPSQ, Kernel


class Rough_RBF(Kernel):
    has_lengthscale = True
    family = FAMILY_EXPSQ

    def distance_weights(self):
        return self.lengthscale.reshape(-1)
