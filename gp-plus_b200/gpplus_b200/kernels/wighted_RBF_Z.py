"""Second experimental weighting kernel, literal semantics only (kernels/wighted_RBF_Z.py:54-78):
``(1/l) * weight`` where ``weight`` is ones with the rows and columns of source-0 points (last input
column == 0) divided by ``1/l``.  Unused by ``GP_Plus``; no device kernel (note 1).
"""
import torch

from .._compat import FAMILY_EXPSQ, Kernel


class wighted_RBF_Z(Kernel):
    has_lengthscale = True
    family = FAMILY_EXPSQ

    def distance_weights(self):
        return self.lengthscale.reshape(-1)

    def forward(self, x1, x2=None, **params):
        x2 = x1 if x2 is None else x2
        inv_l = 1.0 / self.lengthscale.reshape(-1)[0]
        weight = torch.ones(x1.shape[-2], x2.shape[-2], dtype=x1.dtype, device=x1.device)
        weight[x1[..., -1] == 0, :] = weight[x1[..., -1] == 0, :] / inv_l
        weight[:, x2[..., -1] == 0] = weight[:, x2[..., -1] == 0] / inv_l
        return inv_l * weight
