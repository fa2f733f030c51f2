"""CPU oracle for the GP+ exact-GP hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package (``gp-plus_b200/gpplus_b200``) never does.

PARITY UNPINNED: the reference (Bostanabad-Research-Group/GP-Plus) ships no tests, golden vectors
or fixtures, and cannot be imported in this image (its arithmetic lives in gpytorch / botorch, which
are neither vendored under /root/reference nor installed; no version is pinned anywhere in the
reference -- best guess gpytorch 1.8.1 / botorch 0.6.6).  This file
restates, in float64 torch on the CPU, the algorithm the reference reaches through those
libraries, following the reference's own call sites:

  * kernel tree            models/gp_plus.py:219-303, models/gpregression.py:108-111,
                           kernels/Rough_RBF.py:18-32, kernels/matern.py:4-8
  * latent map             models/gp_plus.py:1027-1095 (zeta / one-hot), :1227-1265, :1456-1461
  * forward                models/gp_plus.py:386-484
  * means                  models/gp_plus.py:488-544
  * noise model            likelihoods_noise/multifidelity.py:63-136, models/gpregression.py:58-66
  * marginal likelihood    optim/mll_scipy.py:37-60 (Cholesky log_prob, fast_computations off)
  * prediction             models/gpregression.py:122-149
  * acquisition            bayesian_optimizations/AFs.py:102-159
and the published gpytorch semantics summarised in SURVEY.md Appendix A (covar_dist quadratic
expansion with mean-centring, clamp at 0, MaternKernel / RBFKernel formulas, psd_safe_cholesky
jitter ladder, exact prediction with variance clamped at min_variance).

It is pinned against closed forms derivable by hand from that code (tests/test_oracle.py: N=1 and N=2
likelihoods, kernel limits, finite differences) and, as an independent implementation of the same exact-GP
mathematics, against scikit-learn's GaussianProcessRegressor (log marginal likelihood, its gradient and the
predictive mean / variance for the RBF, Matern-3/2 and Matern-5/2 families).  Neither is the reference
itself, hence "unpinned".

Everything here works on the *natural* parameterisation that crosses the C ABI
(include/gpplus_b200.h): distance weights w, latent table Z, outputscale, noise variances, mean
constants.  ``oracle/gpplus_oracle.py`` adds the raw-parameter / prior layer on top.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

KERNEL_EXPSQ = 0
KERNEL_MATERN32 = 1
KERNEL_MATERN52 = 2

JITTER_LADDER = (0.0, 1e-8, 1e-7, 1e-6)  # psd_safe_cholesky, double precision


class NotPSDError(RuntimeError):
    pass


class NanError(RuntimeError):
    pass


def _t(x, dtype=torch.float64):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def sq_dist(x1: torch.Tensor, x2: torch.Tensor, mode: str = "expansion", centre: Optional[torch.Tensor] = None):
    """Squared euclidean distances between rows of x1 [n1,d] and x2 [n2,d].

    mode="expansion": gpytorch ``Kernel.covar_dist`` / ``Distance._sq_dist``: subtract the
    column mean of x1 from both, one matmul of width d+2, clamp at zero.
    mode="direct": sum of squared differences (the rounding-free definition)."""
    if x1.shape[-1] == 0:
        return torch.zeros(x1.shape[0], x2.shape[0], dtype=x1.dtype)
    if mode == "direct":
        d = x1[:, None, :] - x2[None, :, :]
        return (d * d).sum(-1)
    adj = x1.mean(-2, keepdim=True) if centre is None else centre
    x1 = x1 - adj
    x2 = x2 - adj
    x1n = x1.pow(2).sum(-1, keepdim=True)
    x2n = x2.pow(2).sum(-1, keepdim=True)
    a = torch.cat([-2.0 * x1, x1n, torch.ones_like(x1n)], dim=-1)
    b = torch.cat([x2, torch.ones_like(x2n), x2n], dim=-1)
    return (a @ b.transpose(-1, -2)).clamp_min(0.0)


def quant_corr(s: torch.Tensor, kind: int) -> torch.Tensor:
    """f(s), s = sum_d w_d (x_id - x_jd)^2 (already scaled)."""
    if kind == KERNEL_EXPSQ:  # Rough_RBF.py:27-32 / gpytorch RBFKernel
        return torch.exp(-s)
    r = s.clamp_min(1e-30).sqrt()  # gpytorch covar_dist non-squared distance
    if kind == KERNEL_MATERN32:
        c = math.sqrt(3.0)
        return (1.0 + c * r) * torch.exp(-c * r)
    if kind == KERNEL_MATERN52:
        c = math.sqrt(5.0)
        return ((c * r + 1.0) + 5.0 / 3.0 * r * r) * torch.exp(-c * r)
    raise ValueError("unknown kernel kind %r" % (kind,))


def covariance(xq1, lvl1, xq2, lvl2, w, z, sf2, kind, mode="expansion", centre=None) -> torch.Tensor:
    """sigma_f^2 * exp(-0.5 |z_i - z_j|^2) * f(sum_d w_d dx_d^2)   (gp_plus.py:219-303, gpregression.py:108-111)."""
    sw = w.sqrt() if w.numel() else w
    s = sq_dist(xq1 * sw, xq2 * sw, mode, None if centre is None else centre * sw)
    k = quant_corr(s, kind)
    if z is not None and z.numel() > 0:
        if z.dim() == 3:
            # multi-pass ensemble (gp_plus.py:387-399, 474-482): Sigma = (1/k) sum_p (K_p + m m^T) - m m^T with one
            # latent table per pass; the quantitative factor and the mean are common to the passes
            lat = sum(torch.exp(-0.5 * sq_dist(zp[lvl1], zp[lvl2], mode)) for zp in z) / z.shape[0]
        else:
            lat = torch.exp(-0.5 * sq_dist(z[lvl1], z[lvl2], mode))  # RBFKernel, lengthscale 1 (gp_plus.py:223-226)
        k = k * lat
    return sf2 * k


def _problem_tensors(problem: Dict):
    n = int(problem["n"])
    dq = int(problem["dq"])
    xq = _t(problem["xq"]).reshape(n, dq)
    y = _t(problem["y"]).reshape(n)
    lvl = None if problem.get("level_idx") is None else torch.as_tensor(np.asarray(problem["level_idx"]), dtype=torch.long)
    nidx = torch.zeros(n, dtype=torch.long) if problem.get("noise_idx") is None else torch.as_tensor(
        np.asarray(problem["noise_idx"]), dtype=torch.long)
    midx = torch.zeros(n, dtype=torch.long) if problem.get("mean_idx") is None else torch.as_tensor(
        np.asarray(problem["mean_idx"]), dtype=torch.long)
    return n, dq, xq, y, lvl, nidx, midx


def _mean_vector(midx: torch.Tensor, beta: Optional[torch.Tensor], n_mean: int) -> torch.Tensor:
    """ConstantMean / ZeroMean / multiple_constant gather (gp_plus.py:509-544); index -1 = zero mean."""
    if n_mean == 0 or beta is None:
        return torch.zeros(midx.shape[0], dtype=torch.float64)
    safe = midx.clamp_min(0)
    m = beta[safe]
    return torch.where(midx >= 0, m, torch.zeros_like(m))


def _hyper_tensors(problem: Dict, hyper: Dict, requires_grad: bool):
    dq, dz = int(problem["dq"]), int(problem.get("dz", 0))
    w = _t(hyper["w"]).reshape(dq).clone().requires_grad_(requires_grad)
    z = None
    if dz > 0:
        n_pass = int(problem.get("n_pass", 1) or 1)
        shape = (int(problem["n_combo"]), dz) if n_pass <= 1 else (n_pass, int(problem["n_combo"]), dz)
        z = _t(hyper["z"]).reshape(shape).clone().requires_grad_(requires_grad)
    sf2 = _t(float(hyper["sigma_f2"])).clone().requires_grad_(requires_grad)
    noise = _t(hyper["noise"]).reshape(-1).clone().requires_grad_(requires_grad)
    n_mean = int(problem.get("n_mean", 0))
    beta = None
    if n_mean > 0:
        beta = _t(hyper["beta"]).reshape(n_mean).clone().requires_grad_(requires_grad)
    return w, z, sf2, noise, beta


def psd_safe_cholesky(K: torch.Tensor):
    """gpytorch.utils.cholesky.psd_safe_cholesky: returns (L, jitter_used)."""
    L, info = torch.linalg.cholesky_ex(K)
    if int(info) == 0:
        return L, 0.0
    if torch.isnan(K).any():
        raise NanError("cholesky_cpu: NaN in the covariance matrix")
    for jit in JITTER_LADDER[1:]:
        Kj = K + jit * torch.eye(K.shape[0], dtype=K.dtype)
        L, info = torch.linalg.cholesky_ex(Kj)
        if int(info) == 0:
            return L, jit
    raise NotPSDError("Matrix not positive definite after repeatedly adding jitter up to 1e-06.")


def mll(problem: Dict, hyper: Dict, want_grad: bool = True, mode: str = "expansion", return_mats: bool = False) -> Dict:
    """Negative log marginal likelihood (data term only, NOT divided by n: mll_scipy.py:38-39) and its
    gradient w.r.t. the natural parameters through autograd (mll_scipy.py:123)."""
    n, dq, xq, y, lvl, nidx, midx = _problem_tensors(problem)
    kind = int(problem["kernel"])
    n_mean = int(problem.get("n_mean", 0))
    w, z, sf2, noise, beta = _hyper_tensors(problem, hyper, want_grad)
    K = covariance(xq, lvl, xq, lvl, w, z, sf2, kind, mode)
    Ky = K + torch.diag(noise[nidx])
    m = _mean_vector(midx, beta, n_mean)
    L, jit = psd_safe_cholesky(Ky.detach())
    if want_grad:
        # differentiate through the factorisation like torch autograd does in the reference
        L = torch.linalg.cholesky(Ky + jit * torch.eye(n, dtype=torch.float64))
    r = (y - m).unsqueeze(-1)
    v = torch.linalg.solve_triangular(L, r, upper=False)
    quad = (v * v).sum()
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
    nll = 0.5 * (quad + logdet + n * math.log(2.0 * math.pi))
    out = {"nll": float(nll.detach()), "quad": float(quad.detach()), "logdet": float(logdet.detach()), "jitter": jit}
    if want_grad:
        params = [w, sf2, noise] + ([z] if z is not None else []) + ([beta] if beta is not None else [])
        grads = torch.autograd.grad(nll, params, allow_unused=True)
        gi = iter([torch.zeros_like(p_) if g_ is None else g_ for g_, p_ in zip(grads, params)])
        out["d_w"] = next(gi).numpy().copy()
        out["d_sigma_f2"] = float(next(gi))
        out["d_noise"] = next(gi).numpy().copy()
        if z is not None:
            out["d_z"] = next(gi).numpy().copy()
        if beta is not None:
            out["d_beta"] = next(gi).numpy().copy()
    if return_mats:
        with torch.no_grad():
            Ld = L.detach()
            alpha = torch.cholesky_solve(r.detach(), Ld).squeeze(-1)
            out["K"] = K.detach().numpy()
            out["L"] = Ld.numpy()
            out["alpha"] = alpha.numpy()
            out["Kinv"] = torch.cholesky_inverse(Ld).numpy()
    return out


def predict(problem: Dict, hyper: Dict, cand: Dict, include_noise: bool = False, min_var: float = 1e-10,
            mode: str = "expansion"):
    """Exact predictive mean / variance in SCALED y units (gpregression.py:122-149):
    mu = m* + K*^T alpha, var = clamp(k** - |L^-1 k*|^2 (+ noise of the candidate's source), min_var)."""
    n, dq, xq, y, lvl, nidx, midx = _problem_tensors(problem)
    kind = int(problem["kernel"])
    n_mean = int(problem.get("n_mean", 0))
    with torch.no_grad():
        w, z, sf2, noise, beta = _hyper_tensors(problem, hyper, False)
        m_c = int(cand["m"])
        xc = _t(cand["xq"]).reshape(m_c, dq)
        lc = None if cand.get("level_idx") is None else torch.as_tensor(np.asarray(cand["level_idx"]), dtype=torch.long)
        centre = xq.mean(0, keepdim=True)
        K = covariance(xq, lvl, xq, lvl, w, z, sf2, kind, mode, centre)
        Ky = K + torch.diag(noise[nidx])
        L, _ = psd_safe_cholesky(Ky)
        r = (y - _mean_vector(midx, beta, n_mean)).unsqueeze(-1)
        alpha = torch.cholesky_solve(r, L)
        Ks = covariance(xc, lc, xq, lvl, w, z, sf2, kind, mode, centre)  # [m, n]
        cm = torch.zeros(m_c, dtype=torch.long) if cand.get("mean_idx") is None else torch.as_tensor(
            np.asarray(cand["mean_idx"]), dtype=torch.long)
        mu = _mean_vector(cm, beta, n_mean) + (Ks @ alpha).squeeze(-1)
        V = torch.linalg.solve_triangular(L, Ks.T, upper=False)  # [n, m]
        var = sf2 - (V * V).sum(0)
        if include_noise:
            cn = torch.zeros(m_c, dtype=torch.long) if cand.get("noise_idx") is None else torch.as_tensor(
                np.asarray(cand["noise_idx"]), dtype=torch.long)
            var = var + noise[cn]
        var = var.clamp_min(min_var)
    return mu.numpy(), var.numpy()


def acquisition(mean, std, kind: int, best_f: float, cost, maximize: bool = True, si: float = 0.0):
    """AF_HF_Engineering (kind 0), AF_LF_Engineering (1) and the EI formula of AF_EI (2): AFs.py:67-159."""
    mean = _t(mean)
    sigma = _t(std)
    cost = _t(cost)
    u = (mean - best_f - float(np.sign(best_f)) * si) / sigma
    if not maximize:
        u = -u
    normal = torch.distributions.Normal(torch.zeros_like(u), torch.ones_like(u))
    if kind == 0:
        ei = sigma * u
    elif kind == 1:
        ei = sigma * torch.exp(normal.log_prob(u))
    else:
        ei = sigma * (torch.exp(normal.log_prob(u)) + u * normal.cdf(u))
    return (ei / cost).numpy()
