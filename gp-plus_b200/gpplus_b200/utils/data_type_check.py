"""Inputs of the public API may be tensors, numpy arrays or nested sequences; everything downstream works on
torch tensors.  The notices printed for converted inputs are the reference's (utils/data_type_check.py)."""
import numpy as np
import torch

_NOTICE = "Warning: Data type was %s. GP+ made it a torch tensor to be able to continue."


def data_type_check(data):
    if torch.is_tensor(data):
        return data
    is_array = isinstance(data, np.ndarray)
    print(_NOTICE % ("numpy.ndarray" if is_array else type(data)))
    return torch.from_numpy(data) if is_array else torch.tensor(data)
