"""Standardise quantitative columns with NaN-aware mean / population std
(preprocessing/normalizeX.py:8-72)."""
import torch


def compute_mean_std(tensor):
    means = torch.nanmean(tensor, dim=0)
    present = ~torch.isnan(tensor)
    dev = tensor - means
    dev[~present] = 0
    count = present.sum(dim=0)
    empty = count == 0
    count[empty] = 1
    stds = torch.sqrt(torch.sum(dev ** 2, dim=0) / count)
    stds[empty] = 0
    return means, stds


def standard(Xtrain, qual_index, Xtest=None):
    if not isinstance(Xtrain, torch.Tensor):
        Xtrain = torch.tensor(Xtrain)
    if Xtest is not None and not isinstance(Xtest, torch.Tensor):
        Xtest = torch.tensor(Xtest)
    quant = [c for c in range(Xtrain.shape[1]) if c not in qual_index.keys()]
    if len(quant) == 0:
        return Xtrain
    block = Xtrain[..., quant]
    mean, std = compute_mean_std(block)
    if torch.isnan(block).any():
        print("Warning: There are NaN values in the data. Mean and standard deviation were calculated "
              "excluding these values.")
    Xtrain[..., quant] = (block - mean) / std
    if Xtest is None:
        return Xtrain, mean, std
    Xtest[..., quant] = (Xtest[..., quant] - mean) / std
    return Xtrain, Xtest, mean, std
