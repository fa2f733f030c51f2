"""One seed for every RNG the fit touches.  Restart starting points are drawn from torch's global generator
(optim/mll_scipy._sample_from_prior), data generators use numpy's, so both must be fixed for a reproducible
multi-start list (reference: utils/set_seed.py)."""
import random

import numpy as np
import torch


def set_seed(seed):
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
