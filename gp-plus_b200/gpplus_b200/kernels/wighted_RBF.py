"""Experimental weighting kernel kept importable with its literal semantics.

The reference's ``wighted_RBF.forward`` (kernels/wighted_RBF.py:31-41) builds the Rough-RBF matrix
and then discards it: it returns ``1 + 0*AA + 0*diag(x1.flatten())`` -- an all-ones matrix, valid
only for 1-D inputs.  It is not used by ``GP_Plus`` (models/gp_plus.py:150 whitelist) and is outside
the accelerated hot path (note 1), so no device kernel exists for it.
"""
import torch

from .._compat import FAMILY_EXPSQ, Kernel


class wighted_RBF(Kernel):
    has_lengthscale = True
    family = FAMILY_EXPSQ

    def distance_weights(self):
        return self.lengthscale.reshape(-1)

    def forward(self, x1, x2=None, **params):
        x2 = x1 if x2 is None else x2
        return torch.ones(x1.shape[-2], x2.shape[-2], dtype=x1.dtype, device=x1.device)
