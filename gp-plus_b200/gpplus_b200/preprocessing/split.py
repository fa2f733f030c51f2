"""Train/test split followed by standardisation of the quantitative columns
(preprocessing/split.py:7-47)."""
import torch
from sklearn.model_selection import train_test_split

from .normalizeX import standard
from .numericlevels import setlevels


def train_test_split_normalizeX(X, y, test_size=None, shuffle=True, stratify=None, qual_dict={}, random_state=1,
                                return_mean_std=False, set_levels=False):
    if set_levels:
        X = setlevels(X, qual_index=list(qual_dict.keys()))
    Xtrain, Xtest, ytrain, ytest = train_test_split(X, y, test_size=test_size, shuffle=shuffle,
                                                    random_state=random_state, stratify=stratify)
    Xtrain, Xtest, mean_train, std_train = standard(Xtrain=Xtrain, qual_index=qual_dict, Xtest=Xtest)
    parts = [p if isinstance(p, torch.Tensor) else torch.tensor(p) for p in (Xtrain, Xtest, ytrain, ytest)]
    if return_mean_std:
        return (*parts, mean_train, std_train)
    return tuple(parts)
