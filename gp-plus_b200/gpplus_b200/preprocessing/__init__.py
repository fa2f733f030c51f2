from .numericlevels import setlevels
from .normalizeX import compute_mean_std, standard
from .split import train_test_split_normalizeX

__all__ = ["setlevels", "standard", "compute_mean_std", "train_test_split_normalizeX"]
