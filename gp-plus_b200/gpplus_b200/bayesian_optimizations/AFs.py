"""Acquisition functions of the multi-fidelity BO loop (reference: bayesian_optimizations/AFs.py).

Two families, same names, argument order and sign conventions as the reference:

* ``AF_LF`` / ``AF_HF`` / ``AF_EI`` (AFs.py:1-99) -- objectives for ``scipy.optimize.minimize`` over ONE raw
  design point ``samples = [x_1..x_d, source]``: the point is standardised with ``xmean`` / ``xstd``, the model
  predicts with ``include_noise=True`` and the NEGATIVE cost-scaled utility is returned.
* ``AF_LF_Engineering`` / ``AF_HF_Engineering`` (AFs.py:102-159) -- vectorised utilities of a candidate
  table given predictive mean / std tensors (positive sign, to be arg-maxed).

With ``u = (mean - best_f - sign(best_f) * si) / sigma`` (negated when minimising):
    HF: sigma * u          LF: sigma * pdf(u)          EI: sigma * (pdf(u) + u * cdf(u)),   all divided by cost.

The table branch of ``BO`` does not call these per element; it uses the fused predict + acquisition + arg-max
kernel of the engine (``gpp_acq_argmax``), which evaluates the same formulas on the GPU.  The functions here
are the host-side definitions (used by the scipy branch, by user code and by the parity tests).
"""
from __future__ import annotations

import math

import numpy as np
import torch

_INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)


def _utility(kind: str, mean: torch.Tensor, sigma: torch.Tensor, best_f, maximize: bool, si: float) -> torch.Tensor:
    u = (mean - best_f - float(np.sign(best_f)) * si) / sigma
    if not maximize:
        u = -u
    if kind == "HF":
        return sigma * u
    pdf = torch.exp(-0.5 * u * u) * _INV_SQRT_2PI
    if kind == "LF":
        return sigma * pdf
    cdf = 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0)))
    return sigma * (pdf + u * cdf)


def _costs(cost_fun, source_column: torch.Tensor, shape) -> torch.Tensor:
    vals = [cost_fun(s) for s in source_column.clone().detach()]
    return torch.tensor(vals, dtype=torch.float64).view(shape)


def _flatten(mean: torch.Tensor, std: torch.Tensor):
    """mean[...,m,1] -> [...,m] exactly like the reference's view_shape logic (AFs.py:13-16)."""
    view_shape = mean.shape[:-2] if mean.shape[-2] == 1 else mean.shape[:-1]
    return mean.view(view_shape), std.view(view_shape)


def _point_objective(kind, samples, best_f, model, xmean, xstd, cost_fun, maximize, si):
    samples = np.asarray(samples, dtype=np.float64)
    row = np.concatenate([((samples[0:-1] - xmean) / xstd).reshape(1, -1), samples[-1].reshape(-1, 1)], axis=-1)
    x = torch.tensor(row.reshape(1, -1))
    with torch.no_grad():
        mean, std = model.predict(x, return_std=True, include_noise=True)
    mean, sigma = _flatten(mean.reshape(-1, 1), std)
    cost = _costs(cost_fun, x[:, -1], mean.shape)
    return -1 * (_utility(kind, mean, sigma, best_f, maximize, si) / cost)


def AF_LF(samples, best_f, model, xmean, xstd, cost_fun, maximize=False, si=0.0):
    return _point_objective("LF", samples, best_f, model, xmean, xstd, cost_fun, maximize, si)


def AF_HF(samples, best_f, model, xmean, xstd, cost_fun, maximize=False, si=0.0, data_gen_func=None):
    return _point_objective("HF", samples, best_f, model, xmean, xstd, cost_fun, maximize, si)


def AF_EI(samples, best_f, model, xmean, xstd, cost_fun, maximize=False, si=0.0):
    return _point_objective("EI", samples, best_f, model, xmean, xstd, cost_fun, maximize, si)


def AF_LF_Engineering(best_f, mean, std, x_val, cost_fun, maximize=True, si=0.0, cost=None):
    mean, sigma = _flatten(mean, std)
    cost = _costs(cost_fun, x_val[:, -1], mean.shape)
    return _utility("LF", mean, sigma, best_f, maximize, si) / cost


def AF_HF_Engineering(best_f, mean, std, x_val, cost_fun, maximize=True, si=0.0):
    mean, sigma = _flatten(mean, std)
    cost = _costs(cost_fun, x_val[:, -1], mean.shape)
    return _utility("HF", mean, sigma, best_f, maximize, si) / cost
