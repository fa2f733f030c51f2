"""Seed python / numpy / torch RNGs (utils/set_seed.py:6-15): restart starting points are drawn from
the global torch RNG, so this fixes the multi-start list."""
import random

import numpy as np
import torch


def set_seed(seed):
    random.seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
