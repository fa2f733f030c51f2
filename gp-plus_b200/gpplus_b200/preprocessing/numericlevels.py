"""Ordinal encoding of categorical columns (preprocessing/numericlevels.py:5-53): each listed column
is replaced by the rank of its value among the column's sorted unique values.  Called once per data
set and, in eval mode, on the categorical columns of prediction inputs (models/gp_plus.py:1081-1082).
"""
import numpy as np
import torch

from ..utils.data_type_check import data_type_check


def setlevels(X, qual_index=None, return_label=False):
    if qual_index == []:
        return X
    was_numpy = isinstance(X, np.ndarray)
    if not was_numpy:
        X = data_type_check(X)
    if isinstance(X, torch.Tensor):
        arr = X.detach().cpu().clone().numpy()
    elif was_numpy:
        arr = X.copy()
    else:
        raise TypeError("X must be a PyTorch tensor or a NumPy array.")
    labels = []
    if arr.ndim > 1:
        cols = list(range(arr.shape[-1])) if qual_index is None else qual_index
        for j in cols:
            uniq, inv = np.unique(arr[..., j], return_inverse=True)
            labels.append(uniq.tolist())
            arr[..., j] = inv.reshape(arr[..., j].shape)
    else:
        uniq, inv = np.unique(arr, return_inverse=True)
        arr = inv.astype(arr.dtype)
    if arr.dtype == object:
        arr = arr.astype(float)
    out = torch.from_numpy(arr)
    return (out, labels) if return_label else out
