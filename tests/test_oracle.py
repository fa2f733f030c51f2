"""Pins the CPU oracle: closed forms derivable by hand from the reference code, finite differences,
consistency between its two distance modes and its two layers, and the committed golden fixtures."""
import glob
import math
import os

import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle import gpplus_oracle as GO
from problems import make_candidates, make_hyper, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))


def test_n1_closed_form():
    # one point: K_y = sf2 + noise, nll = 0.5*((y-b)^2/(sf2+noise) + log(sf2+noise) + log 2pi)
    p = {"n": 1, "dq": 2, "dz": 0, "n_combo": 0, "n_noise": 1, "n_mean": 1, "kernel": O.KERNEL_MATERN52,
         "xq": np.array([[0.3, -1.2]]), "y": np.array([0.7]), "level_idx": None, "noise_idx": None, "mean_idx": None}
    h = {"w": np.array([0.5, 2.0]), "z": None, "sigma_f2": 0.9, "noise": np.array([0.05]), "beta": np.array([0.2])}
    r = O.mll(p, h)
    s = 0.95
    assert r["nll"] == pytest.approx(0.5 * (0.25 / s + math.log(s) + math.log(2 * math.pi)), rel=1e-14)
    assert r["d_beta"][0] == pytest.approx(-(0.5) / s, rel=1e-13)
    assert r["d_noise"][0] == pytest.approx(0.5 * (1 / s - 0.25 / s ** 2), rel=1e-13)
    assert np.allclose(r["d_w"], 0.0)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_n2_closed_form(kind):
    x = np.array([[0.0], [0.8]])
    w, sf2, nz = 1.7, 0.6, 0.01
    s = w * 0.64
    rr = math.sqrt(s)
    f = {0: math.exp(-s), 1: (1 + math.sqrt(3) * rr) * math.exp(-math.sqrt(3) * rr),
         2: (1 + math.sqrt(5) * rr + 5 * s / 3) * math.exp(-math.sqrt(5) * rr)}[kind]
    a, b = sf2 + nz, sf2 * f
    y = np.array([0.1, 0.9])
    det = a * a - b * b
    quad = (a * y[0] ** 2 - 2 * b * y[0] * y[1] + a * y[1] ** 2) / det
    want = 0.5 * (quad + math.log(det) + 2 * math.log(2 * math.pi))
    p = {"n": 2, "dq": 1, "dz": 0, "n_combo": 0, "n_noise": 1, "n_mean": 0, "kernel": kind, "xq": x, "y": y,
         "level_idx": None, "noise_idx": None, "mean_idx": None}
    h = {"w": np.array([w]), "z": None, "sigma_f2": sf2, "noise": np.array([nz]), "beta": None}
    for mode in ("expansion", "direct"):
        assert O.mll(p, h, want_grad=False, mode=mode)["nll"] == pytest.approx(want, rel=1e-12)


def test_constant_kernel_limit():
    # all weights -> 0: K = sf2 * ones
    p = make_problem(20, 3, 0, seed=2)
    h = make_hyper(p)
    h["w"] = np.full(3, 1e-300)
    K = O.mll(p, h, want_grad=False, return_mats=True)["K"]
    assert np.allclose(K, h["sigma_f2"], rtol=0, atol=1e-14)


def test_rough_rbf_is_rbf_with_rough_constraint():
    # exp(-sum 10^omega dx^2) == exp(-0.5 sum (dx/l)^2) with l = 2^-1/2 10^(-omega/2)   (gp_plus.py:252)
    import torch
    om = np.array([0.3, -1.1])
    ls = 2.0 ** -0.5 * 10.0 ** (-om / 2)
    x = torch.tensor(np.random.default_rng(0).standard_normal((7, 2)))
    a = torch.exp(-O.sq_dist(x * torch.tensor(10.0 ** om).sqrt(), x * torch.tensor(10.0 ** om).sqrt(), "direct"))
    b = torch.exp(-0.5 * O.sq_dist(x / torch.tensor(ls), x / torch.tensor(ls), "direct"))
    assert torch.allclose(a, b, rtol=1e-13, atol=0)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_gradients_match_finite_differences(kind):
    p = make_problem(40, 3, kind, dz=2, n_combo=4, n_noise=2, n_mean=3, seed=5, zero_mean_group=True)
    h = make_hyper(p, noise=1e-2)
    r = O.mll(p, h)
    eps = 1e-6

    def fd(key, idx):
        hp, hm = {k: (np.array(v, dtype=float, copy=True) if v is not None else None) for k, v in h.items()}, None
        hm = {k: (np.array(v, dtype=float, copy=True) if v is not None else None) for k, v in h.items()}
        if key == "sigma_f2":
            hp[key] = float(h[key]) + eps
            hm[key] = float(h[key]) - eps
        else:
            hp[key][idx] += eps
            hm[key][idx] -= eps
        return (O.mll(p, hp, want_grad=False)["nll"] - O.mll(p, hm, want_grad=False)["nll"]) / (2 * eps)

    assert r["d_sigma_f2"] == pytest.approx(fd("sigma_f2", None), rel=1e-6, abs=1e-7)
    for d in range(3):
        assert r["d_w"][d] == pytest.approx(fd("w", d), rel=1e-6, abs=1e-7)
    assert r["d_z"][1, 0] == pytest.approx(fd("z", (1, 0)), rel=1e-6, abs=1e-7)
    assert r["d_noise"][1] == pytest.approx(fd("noise", 1), rel=1e-5, abs=1e-6)
    assert r["d_beta"][0] == pytest.approx(fd("beta", 0), rel=1e-6, abs=1e-7)


def test_expansion_and_direct_distances_agree():
    p = make_problem(150, 8, 2, dz=2, n_combo=6, seed=7)
    h = make_hyper(p)
    a, b = O.mll(p, h, mode="expansion"), O.mll(p, h, mode="direct")
    assert a["nll"] == pytest.approx(b["nll"], rel=1e-11)
    assert np.allclose(a["d_w"], b["d_w"], rtol=1e-8, atol=1e-9)


def test_predict_interpolates_with_small_noise():
    p = make_problem(30, 2, 0, seed=8, n_mean=0)
    h = make_hyper(p, noise=1e-10)
    h["noise"] = np.array([1e-10])
    c = {"m": 30, "xq": p["xq"], "level_idx": None, "noise_idx": None, "mean_idx": None}
    mu, var = O.predict(p, h, c, include_noise=False)
    assert np.allclose(mu, p["y"], atol=1e-5)
    assert np.all(var >= 1e-10) and np.all(var < 1e-5)


def test_jitter_ladder_and_failures():
    import torch
    K = torch.ones(4, 4, dtype=torch.float64)  # singular: needs jitter
    L, jit = O.psd_safe_cholesky(K - 1e-9 * torch.eye(4, dtype=torch.float64))
    assert jit in (1e-8, 1e-7, 1e-6)
    with pytest.raises(O.NotPSDError):
        O.psd_safe_cholesky(-torch.eye(3, dtype=torch.float64))
    bad = torch.eye(3, dtype=torch.float64)
    bad[1, 1] = float("nan")
    with pytest.raises(O.NanError):
        O.psd_safe_cholesky(bad)


def test_acquisition_formulas():
    mean, std = np.array([1.0, 2.0]), np.array([0.5, 0.25])
    u = (mean - 1.5) / std
    pdf = np.exp(-0.5 * u ** 2) / math.sqrt(2 * math.pi)
    cdf = np.array([0.5 * (1 + math.erf(v / math.sqrt(2))) for v in u])
    cost = np.array([10.0, 1000.0])
    assert np.allclose(O.acquisition(mean, std, 0, 1.5, cost), std * u / cost)
    assert np.allclose(O.acquisition(mean, std, 1, 1.5, cost), std * pdf / cost)
    assert np.allclose(O.acquisition(mean, std, 2, 1.5, cost), std * (pdf + u * cdf) / cost)
    assert np.allclose(O.acquisition(mean, std, 0, 1.5, cost, maximize=False), -std * u / cost)


def test_model_level_oracle_agrees_with_natural_level():
    """gpplus_oracle (raw theta, per-row lookups, priors) minus its prior terms == gp_oracle on the same
    natural parameters."""
    rng = np.random.default_rng(3)
    n = 60
    X = np.hstack([rng.integers(0, 3, (n, 1)).astype(float), rng.standard_normal((n, 3)),
                   rng.integers(0, 2, (n, 1)).astype(float)])
    y = np.sin(X[:, 1]) + 0.2 * X[:, 0] + 0.1 * X[:, 4] + 0.01 * rng.standard_normal(n)
    qual = {0: 3, 4: 2}
    spec = {"X": X, "y": y, "qual_dict": qual, "kernel": "Matern52Kernel", "multiple_noise": True,
            "m_gp": "multiple_constant"}
    layout = GO.theta_layout(spec)
    assert [nm for nm, _ in layout] == ["latent[0, 4]", "likelihood.noise_covar.raw_noise",
                                         "covar_module.raw_outputscale",
                                         "covar_module.base_kernel.kernels.1.raw_lengthscale",
                                         "mean_module_1.constant"]
    p_tot = sum(int(np.prod(s)) if len(s) else 1 for _, s in layout)
    theta = 0.3 * rng.standard_normal(p_tot)
    theta[10:12] = [-4.0, -5.0]
    f = GO.neg_log_posterior(spec, theta, add_prior=False, theta_dtype=__import__("torch").float64, want_grad=False)
    A = theta[:10].reshape(2, 5)
    zeta, lookup = GO._one_hot_table([3, 2])
    lvl = np.array([lookup[str([int(a), int(b)])] for a, b in X[:, [0, 4]]])
    ys = (y - y.min()) / (y.max() - y.min())
    p = {"n": n, "dq": 3, "dz": 2, "n_combo": 6, "n_noise": 2, "n_mean": 1, "kernel": 2, "xq": X[:, 1:4], "y": ys,
         "level_idx": lvl, "noise_idx": X[:, 4].astype(int), "mean_idx": X[:, 4].astype(int) - 1}
    h = {"w": 2 * 10.0 ** theta[13:16], "z": zeta.numpy() @ A.T, "sigma_f2": math.log1p(math.exp(theta[12])),
         "noise": 1e-8 + np.exp(theta[10:12]), "beta": theta[16:17]}
    assert f == pytest.approx(O.mll(p, h, want_grad=False)["nll"], rel=1e-12)


def test_prior_terms_closed_form():
    # horseshoe: log log(1 + 3 (s/(lb+e^v))^2) + v ; lognormal on softplus(raw); N(-3,3) on omega; N(0,1) on beta
    import torch
    rng = np.random.default_rng(5)
    X = rng.standard_normal((12, 2))
    y = rng.standard_normal(12)
    spec = {"X": X, "y": y, "qual_dict": {}, "kernel": "Rough_RBF"}
    theta = np.array([-3.0, 0.4, -1.0, 0.5, 0.1])
    f1 = GO.neg_log_posterior(spec, theta, add_prior=True, theta_dtype=torch.float64, want_grad=False)
    f0 = GO.neg_log_posterior(spec, theta, add_prior=False, theta_dtype=torch.float64, want_grad=False)
    v, rho, om, beta = theta[0], theta[1], theta[2:4], theta[4]
    sf2 = math.log1p(math.exp(rho))
    lp = math.log(math.log(1 + 3 * (0.01 / (1e-8 + math.exp(v))) ** 2)) + v
    lp += -math.log(sf2) - 0.5 * math.log(2 * math.pi) - 0.5 * (math.log(sf2) - 1e-6) ** 2
    lp += sum(-math.log(3.0) - 0.5 * math.log(2 * math.pi) - 0.5 * ((o + 3) / 3) ** 2 for o in om)
    lp += -0.5 * math.log(2 * math.pi) - 0.5 * beta ** 2
    assert (f0 - f1) == pytest.approx(lp, rel=1e-10)


def test_golden_fixtures_reproduce():
    files = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(f).startswith("ref_"))
    assert len(files) >= 6
    for f in files:
        g = np.load(f)
        p = {k[2:]: g[k] for k in g.files if k.startswith("p_")}
        for k in ("n", "dq", "dz", "n_combo", "n_noise", "n_mean", "kernel"):
            p[k] = int(p[k])
        for k in ("level_idx", "noise_idx", "mean_idx"):
            p.setdefault(k, None)
        h = {k[2:]: g[k] for k in g.files if k.startswith("h_")}
        h["sigma_f2"] = float(h["sigma_f2"])
        h.setdefault("z", None)
        h.setdefault("beta", None)
        r = O.mll(p, h, mode="direct")
        assert r["nll"] == pytest.approx(float(g["r_nll"]), rel=1e-11), f
        assert np.allclose(r["d_w"], g["r_d_w"], rtol=1e-7, atol=1e-9), f


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_oracle_against_scikit_learn_gpr(kind):
    """Independent pin: scikit-learn's GaussianProcessRegressor evaluates the same exact-GP log marginal likelihood
    (Cholesky based, Rasmussen & Williams alg. 2.1) and its gradient for ConstantKernel * {RBF, Matern} + WhiteKernel.
    Mapping: k = sf2 * f(sum_d w_d dx_d^2);  RBF: l_d = 1/sqrt(2 w_d);  Matern: l_d = 1/sqrt(w_d)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel
    rng = np.random.default_rng(11)
    n, dq = 70, 4
    X = rng.standard_normal((n, dq))
    y = np.sin(X[:, 0]) + 0.4 * X[:, 1] ** 2 + 0.1 * rng.standard_normal(n)
    w = np.array([0.3, 1.1, 0.05, 0.6])
    sf2, noise = 0.8, 0.02
    p = {"n": n, "dq": dq, "dz": 0, "n_combo": 0, "n_noise": 1, "n_mean": 0, "kernel": kind, "xq": X, "y": y,
         "level_idx": None, "noise_idx": None, "mean_idx": None}
    h = {"w": w, "z": None, "sigma_f2": sf2, "noise": np.array([noise]), "beta": None}
    ref = O.mll(p, h, want_grad=True)
    if kind == 0:
        ls = 1.0 / np.sqrt(2.0 * w)
        base = RBF(length_scale=ls)
    else:
        ls = 1.0 / np.sqrt(w)
        base = Matern(length_scale=ls, nu=1.5 if kind == 1 else 2.5)
    kernel = ConstantKernel(sf2) * base + WhiteKernel(noise)
    gpr = GaussianProcessRegressor(kernel=kernel, optimizer=None, alpha=0.0, normalize_y=False).fit(X, y)
    lml, grad = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    assert abs(-lml - ref["nll"]) <= 1e-9 * abs(ref["nll"])
    # sklearn's theta = log(sf2), log(l_1..l_d), log(noise); d nll/d log sf2 = sf2 * d_sf2; d/d log l_d = -2 w_d d_w
    g_ref = np.concatenate([[sf2 * ref["d_sigma_f2"]], -2.0 * w * ref["d_w"], [noise * ref["d_noise"][0]]])
    assert np.max(np.abs(-grad - g_ref)) <= 1e-7 * max(1.0, np.max(np.abs(g_ref)))
    # predictions
    Xs = rng.standard_normal((9, dq))
    mu, sd = gpr.predict(Xs, return_std=True)
    mu_o, var_o = O.predict(p, h, {"m": 9, "xq": Xs, "level_idx": None}, include_noise=True)
    assert np.max(np.abs(mu - mu_o)) <= 1e-8
    assert np.max(np.abs(sd ** 2 - var_o)) <= 1e-8
