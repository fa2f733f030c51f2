"""Model-level CPU oracle: GP+'s raw-parameter objective restated literally -- TEST INFRASTRUCTURE.

PARITY UNPINNED (see oracle/gp_oracle.py): there are no reference tests or fixtures, and the reference
cannot be imported here (gpytorch / botorch missing).  This file follows the reference's own code path
for one evaluation of ``MLLObjective.fun`` (optim/mll_scipy.py:112-127):

  theta --(float cast, :97)--> raw parameters in ``named_parameters`` order (:70-79)
        --> GP_Plus.forward (models/gp_plus.py:386-484): per-row ``perm_dict[str(row)]`` one-hot lookup
            (:1077-1095), latent map z = zeta A^T (:1456-1461), x_new = [z, x_quant] (:437), constant /
            multiple-constant mean with a per-row loop (:509-544), ScaleKernel(RBF(z; l=1) * k_quant(x))
            (:219-303, gpregression.py:108-111) evaluated with gpytorch's centred quadratic expansion
            (SURVEY A.3), Sigma = (K + m m^T) - m m^T in float64 (:474-481)
        --> likelihood: + diag(lb + exp(raw_noise[source])) (gpregression.py:59, multifidelity.py:105-136)
        --> MultivariateNormal.log_prob via Cholesky (mll_scipy.py:38-39)
        --> + log-priors (mll_scipy.py:40-43; horseshoe.py:63-66, NormalPrior, LogNormalPrior on the
            CONSTRAINED outputscale, MollifiedUniformPrior)
        --> backward (:123)

It shares no code with the product package: it takes a plain ``spec`` dict.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List

import numpy as np
import torch

from . import gp_oracle as O


def _one_hot_table(levels: List[int]):
    """zeta_matrix (gp_plus.py:1027-1073): all level combinations in itertools.product order, their
    concatenated one-hot rows, and the str(row) -> index dictionary."""
    perm = list(itertools.product(*[range(l) for l in levels]))
    rows = []
    for combo in perm:
        r = []
        for lv, nl in zip(combo, levels):
            oh = [0.0] * nl
            oh[lv] = 1.0
            r.extend(oh)
        rows.append(r)
    lookup = {str(list(c)): i for i, c in enumerate(perm)}
    return torch.tensor(rows, dtype=torch.float64), lookup


def theta_layout(spec: Dict):
    """[(name, shape)] in the reference's ``named_parameters`` order for trainable parameters."""
    out = []
    qual = spec.get("qual_dict", {})
    prob = spec.get("embedding_type", "deterministic") == "probabilistic"
    if len(qual) > 0 and not prob:
        cols = list(qual.keys())
        out.append(("latent" + str(cols), (2, sum(qual.values()))))
    n_noise = list(qual.values())[-1] if spec.get("multiple_noise", False) else 1
    if not spec.get("fix_noise", False):
        out.append(("likelihood.noise_covar.raw_noise", (n_noise,)))
    out.append(("covar_module.raw_outputscale", ()))
    d = spec["X"].shape[1]
    nq = d - len(qual)
    if nq > 0:
        name = "covar_module.base_kernel.kernels.1.raw_lengthscale" if len(qual) > 0 \
            else "covar_module.base_kernel.raw_lengthscale"
        out.append((name, (1, nq)))
    if len(qual) > 0 and prob:
        # Variational_Encoder registered as the sub-module A_matrix after the kernel (gp_plus.py:347-349, 1373-1385)
        out.append(("A_matrix.fci.fciweight", (5, sum(qual.values()))))
        out.append(("A_matrix.fci.fcibias", (5,)))
    m_gp = spec.get("m_gp", "single_constant")
    if m_gp == "single_constant":
        out.append(("mean_module.constant", (1,)))
    elif m_gp == "multiple_constant":
        n_src = int(np.max(spec["X"][:, -1]))
        first = 0 if spec.get("m_gp_ref", "zero") == "constant" else 1
        for s in range(first, n_src + 1):
            out.append(("mean_module_%d.constant" % s, (1,)))
    return out


def neg_log_posterior(spec: Dict, theta: np.ndarray, add_prior: bool = True, theta_dtype=torch.float32,
                      want_grad: bool = True, mode: str = "expansion"):
    """(nll, grad) exactly as ``MLLObjective.fun`` returns them."""
    X = torch.as_tensor(np.asarray(spec["X"]), dtype=torch.float64)
    y_raw = torch.as_tensor(np.asarray(spec["y"]), dtype=torch.float64).reshape(-1)
    qual = spec.get("qual_dict", {})
    kname = spec.get("kernel", "Rough_RBF")
    lb = float(spec.get("lb_noise", 1e-8))
    n = X.shape[0]
    y = (y_raw - y_raw.min()) / (y_raw.max() - y_raw.min())  # gpregression.py:67-69

    params = {}
    i = 0
    th = torch.as_tensor(np.asarray(theta, dtype=np.float64))
    for name, shape in theta_layout(spec):
        k = int(np.prod(shape)) if len(shape) else 1
        v = th[i:i + k].to(theta_dtype).to(torch.float64).reshape(shape if len(shape) else ())
        params[name] = v.clone().requires_grad_(want_grad)
        i += k
    assert i == th.numel(), "theta has %d entries, layout needs %d" % (th.numel(), i)

    qcols = list(qual.keys())
    quant_cols = [c for c in range(X.shape[1]) if c not in qcols]
    prob = spec.get("embedding_type", "deterministic") == "probabilistic"
    parts = []
    if len(qual) > 0:
        zeta, lookup = _one_hot_table(list(qual.values()))
        idx = [lookup[str([int(v) for v in row])] for row in X[:, qcols].tolist()]  # per-row dict lookup
        if prob:
            # gp_plus.py:388-437: generator re-seeded at every forward; per pass epsilon ~ N(0,1) [u, 2] for the u
            # unique one-hot rows of the inputs (torch.unique order); encoder output (L22, L21, L11, mu2, mu1);
            # positions copied into a float32 buffer (x_raw) -> float32 rounding with a pass-through gradient
            rows = zeta[idx]
            uniq, inverse = torch.unique(rows, dim=0, return_inverse=True)
            gen = torch.Generator().manual_seed(int(spec.get("seed_number", 1)))
            W, b = params["A_matrix.fci.fciweight"], params["A_matrix.fci.fcibias"]
            for _ in range(int(spec.get("num_pass_train", 1))):
                eps = torch.normal(mean=0.0, std=1.0, size=[uniq.shape[0], 2], generator=gen).to(torch.float64)
                o = uniq @ W.T + b
                x1 = o[:, 4:5] + torch.abs(o[:, 2:3]) * eps[:, 0:1]
                x2 = o[:, 3:4] + o[:, 1:2] * eps[:, 0:1] + torch.abs(o[:, 0:1]) * eps[:, 1:2]
                pos = torch.cat([x1, x2], 1)
                pos = pos + (pos.float().double() - pos).detach()
                parts.append(pos[inverse])
        else:
            A = params["latent" + str(qcols)]
            parts.append(zeta[idx] @ A.T)  # Linear_MAP, no bias
    x_quant = X[:, quant_cols]

    # means
    m_gp = spec.get("m_gp", "single_constant")
    if m_gp == "single_constant":
        mean = params["mean_module.constant"].expand(n)
    elif m_gp == "single_zero":
        mean = torch.zeros(n, dtype=torch.float64)
    else:
        vals = []
        for r in range(n):  # per-row module call in the reference (gp_plus.py:532-534)
            s = int(X[r, -1])
            key = "mean_module_%d.constant" % s
            vals.append(params[key].reshape(()) if key in params else torch.zeros((), dtype=torch.float64))
        mean = torch.stack(vals)

    # kernel tree
    k = torch.ones(n, n, dtype=torch.float64)
    if len(qual) > 0:
        # one latent table per forward pass; the passes are averaged (gp_plus.py:474-482; the quantitative factor,
        # the output scale and the mean are the same in every pass)
        k = k * sum(torch.exp(-0.5 * O.sq_dist(z, z, mode)) for z in parts) / len(parts)  # RBFKernel, lengthscale 1
    if len(quant_cols) > 0:
        raw = params["covar_module.base_kernel.kernels.1.raw_lengthscale" if len(qual) > 0
                     else "covar_module.base_kernel.raw_lengthscale"].reshape(-1)
        if kname == "RBFKernel":
            ls = torch.exp(raw)
        else:
            ls = 2.0 ** (-0.5) * torch.pow(10.0, -raw / 2)  # gp_plus.py:252
        xs = x_quant / ls
        sq = O.sq_dist(xs, xs, mode)
        if kname in ("Rough_RBF", "RBFKernel"):
            k = k * torch.exp(-0.5 * sq)
        elif kname == "Matern32Kernel":
            k = k * O.quant_corr(sq, O.KERNEL_MATERN32)
        elif kname == "Matern52Kernel":
            k = k * O.quant_corr(sq, O.KERNEL_MATERN52)
        else:
            raise ValueError(kname)
    sf2 = torch.nn.functional.softplus(params["covar_module.raw_outputscale"])
    K = sf2 * k
    Sigma = (K + torch.outer(mean, mean)) - torch.outer(mean, mean)  # gp_plus.py:474-481

    if spec.get("fix_noise", False):
        noise = torch.full((1,), float(spec.get("fix_noise_val", 1e-5)), dtype=torch.float64)
        raw_noise = None
    else:
        raw_noise = params["likelihood.noise_covar.raw_noise"]
        noise = lb + torch.exp(raw_noise)
    if spec.get("multiple_noise", False):
        src = X[:, -1].long()
        diag = noise[src]
    else:
        diag = noise.expand(n)
    Ky = Sigma + torch.diag(diag)
    _, jit = O.psd_safe_cholesky(Ky.detach())
    L = torch.linalg.cholesky(Ky + jit * torch.eye(n, dtype=torch.float64))
    r = (y - mean).unsqueeze(-1)
    v = torch.linalg.solve_triangular(L, r, upper=False)
    logp = -0.5 * ((v * v).sum() + n * math.log(2 * math.pi)) - torch.log(torch.diagonal(L)).sum()

    if add_prior:
        def normal(loc, scale):  # float64 prior constants (python floats would give float32 buffers)
            return torch.distributions.Normal(torch.tensor(loc, dtype=torch.float64),
                                              torch.tensor(scale, dtype=torch.float64))

        if len(qual) > 0 and prob:
            logp = logp + normal(0.0, 0.2).log_prob(params["A_matrix.fci.fciweight"]).sum()
            logp = logp + normal(0.0, 0.05).log_prob(params["A_matrix.fci.fcibias"]).sum()
        elif len(qual) > 0:
            logp = logp + normal(0.0, 1.0).log_prob(params["latent" + str(qcols)]).sum()
        # marginal_log_likelihood adds EVERY registered prior (mll_scipy.py:40-43 has no requires_grad test): with
        # fix_noise the horseshoe term of the frozen raw noise is a constant offset of the objective
        scale = 0.01
        rn = raw_noise if raw_noise is not None else torch.log(noise - lb)
        logp = logp + (torch.log(torch.log(1 + 3 * (scale / (lb + torch.exp(rn))) ** 2)) + rn).sum()
        logp = logp + torch.distributions.LogNormal(torch.tensor(1e-6, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64)).log_prob(sf2).sum()
        if len(quant_cols) > 0:
            if kname == "RBFKernel":
                a, b, ts = math.log(0.1), math.log(10), 0.1
                tail = ((raw - (a + b) / 2).abs() - (b - a) / 2).clamp(min=0)
                logp = logp + (normal(0.0, ts).log_prob(tail) - math.log(1 + (b - a) / (math.sqrt(2 * math.pi) * ts))).sum()
            else:
                logp = logp + normal(-3.0, 3.0).log_prob(raw).sum()
        for name, p in params.items():
            if name.startswith("mean_module"):
                logp = logp + normal(0.0, 1.0).log_prob(p).sum()
    if spec.get("interval_score", False):
        # mll_scipy.py:57-59: interval score of the PRIOR at the training inputs (training-mode forward): mean and
        # variance diag(Sigma), clamped at gpytorch's min_variance, against the scaled targets
        var = torch.diagonal(Sigma).clamp_min(1e-10)
        up, lo_ = mean + 1.96 * var.sqrt(), mean - 1.96 * var.sqrt()
        sc = up - lo_
        sc = sc + (y > up).to(torch.int64) * 2 / 0.05 * (y - up)
        sc = sc + (y < lo_).to(torch.int64) * 2 / 0.05 * (lo_ - y)
        logp = logp - 0.08 * torch.abs(logp) * sc.mean()
    obj = -logp
    if not want_grad:
        return float(obj)
    grads = torch.autograd.grad(obj, list(params.values()), allow_unused=True)
    g = np.concatenate([(torch.zeros_like(p) if gi is None else gi).reshape(-1).numpy()
                        for gi, p in zip(grads, params.values())])
    return float(obj.detach()), g.astype(np.float64)
