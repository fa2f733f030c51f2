"""Analytical workload generators used by the example configs (test_functions/analytical.py:6-165 of
the reference): the wing-weight and borehole functions on Sobol designs, and the mixed-variable
borehole whose categorical columns take a few random levels inside the bounds.

These are data generators, not part of the accelerated path; they exist so the benchmark and tests
can rebuild the reference's example workloads (same bounds, same Sobol design, same numpy RNG use).
"""
import numpy as np
from scipy.stats.qmc import Sobol, scale

from ..preprocessing import setlevels

WING_BOUNDS = ([150, 220, 6, -10, 16, 0.5, 0.08, 2.5, 1700, 0.025], [200, 300, 10, 10, 45, 1, 0.18, 6, 2500, 0.08])
BOREHOLE_BOUNDS = ([0.05, 100, 63070, 990, 63.1, 700, 1120, 9855], [0.15, 50000, 115600, 1110, 116, 820, 1680, 12045])


def _sobol_design(n, bounds, seed):
    lo, hi = bounds
    pts = Sobol(d=len(lo), seed=seed).random(2 ** (np.log2(n) + 1).astype(int))[:n, :]
    return scale(pts, l_bounds=lo, u_bounds=hi)


def _finish(X, y, generated, noise_std, shuffle):
    if shuffle:
        pick = np.random.randint(0, len(y), size=len(y))  # sampling WITH replacement, as the reference does
        X, y = X[pick, ...], y[pick]
    if noise_std > 0.0:
        y = y + np.random.randn(*y.shape) * noise_std
    return (X, y) if generated else y


def wing_weight(X, sw_exp=0.758, wp_mode="sw"):
    Sw, Wfw, A = X[..., 0], X[..., 1], X[..., 2]
    gam = X[..., 3] * (np.pi / 180.0)
    q, lam, tc, Nz, Wdg, Wp = X[..., 4], X[..., 5], X[..., 6], X[..., 7], X[..., 8], X[..., 9]
    core = 0.036 * Sw ** sw_exp * Wfw ** 0.0035 * (A / (np.cos(gam)) ** 2) ** 0.6 * q ** 0.006 * lam ** 0.04 \
        * ((100 * tc) / (np.cos(gam))) ** (-0.3) * (Nz * Wdg) ** 0.49
    paint = {"sw": Sw * Wp, "one": 1 * Wp, "zero": 0 * Wp}[wp_mode]
    return core + paint


def wing(n=100, X=None, noise_std=0.0, random_state=None, shuffle=True):
    if random_state is not None:
        np.random.seed(random_state)
    generated = X is None
    if generated:
        X = _sobol_design(n, WING_BOUNDS, random_state)
    X = np.asarray(X)
    return _finish(X, wing_weight(X), generated, noise_std, shuffle)


def borehole_flow(X):
    rw, r, Tu, Hu, Tl, Hl, L, Kw = [X[..., i] for i in range(8)]
    lg = np.log(r / rw)
    return 2 * np.pi * Tu * (Hu - Hl) / (lg * (1 + 2 * L * Tu / (lg * rw ** 2 * Kw) + Tu / Tl))


def borehole(n=100, X=None, noise_std=0.0, random_state=None, shuffle=True):
    if random_state is not None:
        np.random.seed(random_state)
    generated = X is None
    if generated:
        X = _sobol_design(n, BOREHOLE_BOUNDS, random_state)
    X = np.asarray(X)
    return _finish(X, borehole_flow(X), generated, noise_std, shuffle)


def borehole_mixed_variables(n=100, X=None, qual_dict={0: 5, 6: 3}, noise_std=0.0, random_state=None, shuffle=True):
    generated = X is None
    if generated:
        X = _sobol_design(n, BOREHOLE_BOUNDS, random_state)
        lo, hi = BOREHOLE_BOUNDS
        for col, n_levels in qual_dict.items():
            levels = np.random.uniform(lo[col], hi[col], size=n_levels)
            X[..., col] = np.random.choice(levels, size=len(X), replace=True)
    X = np.asarray(X)
    y = borehole_flow(X)
    if shuffle:
        pick = np.random.randint(0, len(y), size=len(y))
        X, y = X[pick, ...], y[pick]
    X = setlevels(X, qual_index=list(qual_dict.keys()))
    if noise_std > 0.0:
        y = y + np.random.randn(*y.shape) * noise_std
    return (X, y) if generated else y


def sine_1D(n=100, X=None, noise_std=0.0, frequency=1.0, absolute_value_flag=False, random_state=None, shuffle=True):
    """y = sin(2 pi f x) on [-1, 1] (test_functions/analytical.py:224-255 of the reference, Example 05)."""
    if random_state is not None:
        np.random.seed(random_state)
    generated = X is None
    if generated:
        X = _sobol_design(n, ([-1.0], [1.0]), random_state)
    X = np.asarray(X)
    y = np.sin(2 * np.pi * frequency * X[:, 0])
    if absolute_value_flag:
        y = np.abs(y)
    return _finish(X, y, generated, noise_std, shuffle)
