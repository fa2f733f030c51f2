"""Constraint transform of the kernel outputscale: sigma_f^2 = softplus(raw) (reference: utils/transforms.py,
used by models/gpregression.py:108-111).

``softplus`` is a plain function object on purpose: the closed-form host objective (optim/_fast_objective.py)
recognises the outputscale constraint by identity with it.
"""
import torch
import torch.nn.functional as F


def softplus(raw: torch.Tensor) -> torch.Tensor:
    """log(1 + e^raw) with torch's default switch to the identity above 20 (torch.nn.Softplus defaults)."""
    return F.softplus(raw, beta=1.0, threshold=20.0)


def inv_softplus(value: torch.Tensor) -> torch.Tensor:
    """raw such that softplus(raw) = value, i.e. log(e^value - 1) evaluated as value + log(1 - e^-value)."""
    one_minus_exp = -torch.expm1(-value)
    return value + torch.log(one_minus_exp)
