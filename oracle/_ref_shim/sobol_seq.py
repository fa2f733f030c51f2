"""Stub of the ``sobol_seq`` package (models/gp_plus.py:20): scipy's unscrambled Sobol sequence, first point skipped."""
import numpy as np


def i4_sobol_generate(dim_num, n, skip=1):
    from scipy.stats.qmc import Sobol
    return np.asarray(Sobol(d=dim_num, scramble=False).random(n + skip)[skip:])
