"""Multi-fidelity workload generators (test_functions/multi_fidelity.py:7-226 of the reference):
four wing-weight fidelities and the five-source borehole used by the MFBO example.  The reference
draws the borehole design with ``pyDOE.lhs`` (not installed here); the Latin hypercube is restated
with numpy (one stratified uniform draw per dimension, independently permuted).
"""
import numpy as np
import torch

from .analytical import WING_BOUNDS, _sobol_design, wing_weight

_WING_VARIANTS = {0: (0.758, "sw"), 1: (0.758, "one"), 2: (0.8, "one"), 3: (0.9, "zero")}


def wing(n=100, X=None, fidelity=0, noise_std=0.0, random_state=None, shuffle=True):
    if random_state is not None:
        np.random.seed(random_state)
    if fidelity not in _WING_VARIANTS:
        raise ValueError("only 4 fidelities of 0,1,2,3 have been implemented ")
    generated = X is None
    if generated:
        X = _sobol_design(n, WING_BOUNDS, random_state)
    X = np.asarray(X)
    y = wing_weight(X, *_WING_VARIANTS[fidelity])
    if noise_std > 0.0:
        y = y + np.random.randn(*y.shape) * noise_std
    return (X, y) if generated else y


def multi_fidelity_wing(X=None, n={"0": 50, "1": 100, "2": 100, "3": 100},
                        noise_std={"0": 0.0, "1": 0.0, "2": 0.0, "3": 0.0}, random_state=None, shuffle=True):
    if X is None:
        xs, ys = [], []
        for level, num in n.items():
            if level not in ("0", "1", "2", "3") or num <= 0:
                raise ValueError("Wrong label, should be h, l1, l2 or l3")
            Xl, yl = wing(n=num, fidelity=int(level), noise_std=noise_std[level], random_state=random_state)
            xs.append(np.hstack([Xl, np.full((num, 1), float(level))]))
            ys.append(yl)
        return np.vstack(xs), np.hstack(ys)
    X = torch.as_tensor(np.asarray(X))
    ys = []
    for f in n.keys():
        rows = torch.nonzero(X[..., -1] == int(f)).reshape(-1)
        ys.append(wing(X=X[rows, 0:-1].numpy(), fidelity=int(f), noise_std=noise_std[f]))
    return torch.tensor(np.hstack(ys))


def multi_fidelity_wing_value(input):
    out = []
    for row in input:
        if float(row[-1]) not in (0.0, 1.0, 2.0, 3.0):
            raise ValueError("Wrong label, should be 0, 1, 2 or 3")
        out.append(wing(X=np.asarray(row)))
    return torch.tensor(np.hstack(out))


def _lhs(dim, samples):
    cut = np.linspace(0.0, 1.0, samples + 1)
    u = np.random.rand(samples, dim)
    pts = cut[:samples, None] + u * (cut[1:, None] - cut[:samples, None])
    for j in range(dim):
        pts[:, j] = pts[np.random.permutation(samples), j]
    return pts


def _bh(x, hu=1.0, hl=1.0, lfac=2.0, tfac=1.0, rfac=1.0):
    Tu, Hu, Hl, r, rw, Tl, L, Kw = [x[:, i] for i in range(8)]
    lg = np.log(r / rw)
    return (2 * np.pi * Tu * (hu * Hu - hl * Hl)) / (np.log(rfac * r / rw) * (1 + (lfac * L * Tu) / (lg * rw ** 2 * Kw)
                                                                                  + tfac * (Tu / Tl)))


_BH_SOURCES = [dict(), dict(hl=0.8, lfac=1.0), dict(lfac=8.0, tfac=0.75), dict(hu=1.09, lfac=3.0, rfac=4.0),
               dict(hu=1.05, lfac=3.0, rfac=2.0)]
BH_MIN = (100, 990, 700, 100, .05, 10, 1000, 6000)
BH_MAX = (1000, 1110, 820, 10000, .15, 500, 2000, 12000)


def Borehole_MF_BO(init_data, x, var=(0, 0, 0, 0, 0)):
    """init_data=True: x maps source -> number of initial samples, returns (x_train [n,9], y_train [n,1]);
    otherwise evaluates the sources named by the last column of x (the high-fidelity source is noisy)."""
    if init_data:
        counts = tuple(x.values())
        span = np.array(BH_MAX, dtype=float) - np.array(BH_MIN, dtype=float)
        xs, ys = [], []
        for src, cnt in enumerate(counts[:5]):
            pts = _lhs(8, cnt) * span + np.array(BH_MIN, dtype=float)
            yv = _bh(pts, **_BH_SOURCES[src])
            if src == 0:
                yv = yv + np.random.randn(*yv.shape) * 4
            yv = yv.reshape(cnt, 1)
            yv = yv + np.sqrt(var[src]) * np.random.standard_normal(size=yv.shape)
            xs.append(np.hstack([pts, np.full((cnt, 1), float(src))]))
            ys.append(yv)
        return np.vstack(xs), np.vstack(ys)
    X = np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=float)
    out = []
    for row in X:
        src = int(row[-1])
        if src not in range(5):
            raise ValueError("Wrong label, should be h, l1, l2 or l3")
        yv = _bh(row[None, :-1], **_BH_SOURCES[src])
        if src == 0:
            yv = yv + np.random.randn(*yv.shape) * 2
        out.append(yv)
    return torch.tensor(np.hstack(out))
