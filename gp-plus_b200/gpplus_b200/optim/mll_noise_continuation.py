"""Noise-continuation fitting (optim/mll_noise_continuation.py:28-244 of the reference): the noise
variance is frozen and walked down a ladder; every rung is a full multi-start ``fit_model_scipy``
warm-started from the distinct optima of the previous rung; the ladder is then refined once around
the best rung.  A pure caller of the engine-backed ``fit_model_scipy``.
"""
import math
from copy import deepcopy
from typing import Dict, Tuple

import numpy as np
import torch
from scipy.spatial import distance_matrix

from .mll_scipy import fit_model_scipy


def loocv_rrmse(model) -> float:
    """Leave-one-out RMSE in scaled units: rms of alpha_i / (K_y^-1)_ii (mll_noise_continuation.py:28-42);
    alpha and K_y^-1 come from the engine's factorisation of the current hyper-parameters."""
    model.eval()
    with torch.no_grad():
        eng = model._get_engine()
        eng.mll_grad(model._hyper_numpy(), want_grad=True)  # leaves alpha and K_y^-1 on the device
        alpha = eng.fetch("alpha")
        kinv_diag = eng.fetch("Kinv_diag")  # n doubles (a strided device copy), not the N x N matrix
    return float(np.sqrt(np.mean((alpha / kinv_diag) ** 2)))


def _distinct_optima(reslist):
    starts = []
    for res in reslist:
        if isinstance(res, Exception):
            continue
        if len(starts) > 0:
            d = distance_matrix(res.x.reshape(1, -1), np.vstack(starts)).ravel()
            if np.any(d < 1e-2 * res.x.shape[0]):
                continue
        starts.append(res.x)
    return starts


def fit_model_continuation(model, add_prior: bool = True, num_restarts: int = 32, criterion: str = "NLL",
                           initial_noise_var: float = 1, red_factor: float = math.sqrt(10), options: Dict = {},
                           n_jobs: int = -1, accuracy=1e-2, method="L-BFGS-B", constraint=False,
                           regularization_parameter=[0, 0], bounds=False) -> Tuple[float, Dict]:
    if criterion.upper() not in ["NLL", "LOOCV"]:
        raise AttributeError("criterion must be one of NLL or LOOCV")
    if red_factor < 2:
        raise RuntimeError("Reduction factor for noise variance needs to be greater then 2")
    if model.likelihood.raw_noise.requires_grad:
        model.likelihood.raw_noise.requires_grad_(False)

    theta0_list = None
    history, states, best = None, None, None
    first = True
    while True:
        start_noise = initial_noise_var
        if first:
            noises = [start_noise / (10 ** i) for i in range(10)]
            first = False
        else:
            hist_n = history["noise_history"]
            if 1 <= best < len(hist_n) - 1:
                noises = np.linspace(float(hist_n[best - 1]), float(hist_n[best + 1]), 10)
                initial_noise_var = hist_n[best - 1]
                model.load_state_dict(states[best - 1])
            else:
                model.load_state_dict(states[best])
                print(f"Negative log likelihood={history['nll_history'][best]}")
                return history["nll_history"][best], history

        noise_list, nll_list, loocv_list, reslist_list = [], [], [], []
        states = {}
        for i in range(len(noises)):
            model.train()
            model.likelihood.initialize(**{"noise": noises[i]})
            states[i] = deepcopy(model.state_dict())
            reslist, nll = fit_model_scipy(model, add_prior, num_restarts=num_restarts, theta0_list=theta0_list,
                                           options=options, n_jobs=n_jobs, method=method, constraint=constraint,
                                           regularization_parameter=regularization_parameter, bounds=bounds)
            if all(isinstance(res, (RuntimeError, TypeError)) for res in reslist):
                break  # every restart failed numerically at this noise level
            noise_list.append(model.likelihood.noise.data)
            nll_list.append(nll)
            loocv_list.append("NLL")
            reslist_list.append(reslist)
            theta0_list = _distinct_optima(reslist)
            try:
                model.likelihood.initialize(**{"noise": noise_list[-1] / red_factor})
            except Exception:
                try:
                    model.likelihood.initialize(**{"noise": noise_list[-1] / red_factor + 1e-10})
                except Exception:
                    break

        history = {"noise_history": noise_list, "nll_history": nll_list, "loocv_history": loocv_list,
                   "optimization_history": reslist_list}
        best = int(np.argmin(history["nll_history"]))
        print("Finished for loop")
        print(history["nll_history"])
        if np.abs(float(start_noise) - float(history["noise_history"][best])) < accuracy:
            model.load_state_dict(states[best])
            break
    return history["nll_history"][best], history
