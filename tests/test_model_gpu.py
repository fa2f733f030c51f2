"""The drop-in Python API on the GPU: MLLObjective.fun, fit_model_scipy and predict of GP_Plus against the
model-level CPU oracle, on the reference's example workloads (BASELINE.json configs 1-3)."""
import math

import numpy as np
import pytest
import torch
from scipy.optimize import minimize

pytestmark = pytest.mark.gpu


def _c1():
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.preprocessing import train_test_split_normalizeX
    from gpplus_b200.test_functions import borehole
    from gpplus_b200.utils import set_seed
    set_seed(1245)
    X, y = borehole(n=2000, random_state=12345)
    Xtr, Xte, ytr, yte = train_test_split_normalizeX(X, y, test_size=0.9)
    m = GP_Plus(Xtr, ytr, dtype=torch.float64)
    return m, {"X": Xtr.numpy(), "y": ytr.numpy(), "qual_dict": {}, "kernel": "Rough_RBF"}, Xte, yte


def _c2(kernel="Rough_RBF"):
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.preprocessing import train_test_split_normalizeX
    from gpplus_b200.test_functions import borehole_mixed_variables
    from gpplus_b200.utils import set_seed
    set_seed(4)
    qd = {0: 5, 5: 5}
    X, y = borehole_mixed_variables(n=2000, qual_dict=qd, random_state=4)
    Xtr, Xte, ytr, yte = train_test_split_normalizeX(X, y, test_size=0.75, qual_dict=qd)
    m = GP_Plus(Xtr, ytr, qual_dict=qd, dtype=torch.float64, quant_correlation_class=kernel)
    return m, {"X": Xtr.numpy(), "y": ytr.numpy(), "qual_dict": qd, "kernel": kernel}, Xte, yte


def _c3():
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.preprocessing import train_test_split_normalizeX
    from gpplus_b200.test_functions import multi_fidelity_wing
    from gpplus_b200.utils import set_seed
    set_seed(4)
    X, y = multi_fidelity_wing(n={"0": 150, "1": 300, "2": 300, "3": 300}, random_state=4)
    qd = {10: 4}
    Xtr, Xte, ytr, yte = train_test_split_normalizeX(X, y, test_size=0.05, qual_dict=qd, stratify=X[:, -1])
    m = GP_Plus(Xtr, ytr, qual_dict=qd, multiple_noise=True, m_gp="multiple_constant", dtype=torch.float64)
    spec = {"X": Xtr.numpy(), "y": ytr.numpy(), "qual_dict": qd, "kernel": "Rough_RBF", "multiple_noise": True,
            "m_gp": "multiple_constant"}
    return m, spec, Xte, yte


@pytest.mark.parametrize("maker", [_c1, _c2, lambda: _c2("Matern52Kernel"), lambda: _c2("Matern32Kernel"),
                                   lambda: _c2("RBFKernel"), _c3])
def test_objective_matches_oracle(maker):
    from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior
    from oracle import gpplus_oracle as GO
    m, spec, _, _ = maker()
    obj = MLLObjective(m, True, [0, 0])
    torch.manual_seed(0)
    thetas = [obj.pack_parameters()] + [_sample_from_prior(m) for _ in range(2)]
    for th in thetas:
        th = np.clip(th, -6, 4)
        f_ref, g_ref = GO.neg_log_posterior(spec, th)
        f, g = obj.fun(th)
        assert abs(f - f_ref) <= 1e-9 * abs(f_ref)
        assert np.max(np.abs(g - g_ref)) <= 1e-7 * max(1.0, np.max(np.abs(g_ref)))
        assert obj.fun(th, False) == f
    # after fun() the model holds theta cast through float32 (side-effect contract, mll_scipy.py:115-117)
    assert np.array_equal(obj.pack_parameters(), th.astype(np.float32).astype(np.float64))


def test_fit_matches_cpu_oracle_fit_from_same_starts():
    from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior, fit_model_scipy
    from oracle import gpplus_oracle as GO
    m, spec, Xte, yte = _c1()
    torch.manual_seed(3)
    starts = [_sample_from_prior(m) for _ in range(4)]
    out, best = fit_model_scipy(m, theta0_list=[s.copy() for s in starts], bounds=True)
    assert len(out) == 4 and best == min(r.fun for r in out)
    opts = {"ftol": 1e-6, "gtol": 1e-5, "maxfun": 5000, "maxiter": 2000}
    from scipy.optimize import Bounds
    from gpplus_b200.optim.mll_scipy import get_bounds
    lo, hi = get_bounds(MLLObjective(m, True, [0, 0]), starts[0])
    cpu = [minimize(lambda t: GO.neg_log_posterior(spec, t), s, jac=True, method="L-BFGS-B", bounds=Bounds(lo, hi),
                    options=opts) for s in starts]
    # same optimiser, same starts, objective equal to ~1e-12: the trajectories coincide to optimiser tolerance
    for a, b in zip(out, cpu):
        assert abs(a.fun - b.fun) <= 1e-4 * max(1.0, abs(b.fun))
    # best theta is loaded into the model; predictions are sane on held-out points
    mean, std = m.predict(Xte[:500], return_std=True)
    rrmse = float(torch.sqrt(torch.mean((mean - yte[:500]) ** 2)) / torch.std(yte[:500]))
    assert rrmse < 0.2 and bool(torch.all(std > 0))


def test_predict_matches_oracle_mixed_multifidelity():
    from oracle import gp_oracle as O
    m, spec, Xte, yte = _c3()
    with torch.no_grad():
        m.covar_module.base_kernel.kernels[1].raw_lengthscale.fill_(-0.7)
        m.likelihood.noise_covar.raw_noise.copy_(torch.tensor([-6.0, -5.0, -4.0, -7.0], dtype=torch.float64))
        getattr(m, "latent[10]").copy_(torch.tensor([[0.0, 0.4, -0.3, 0.8], [0.0, -0.5, 0.6, 0.2]],
                                                      dtype=torch.float64))
        m.mean_module_2.constant.fill_(0.1)
    Xq = Xte[:40]
    mean, std = m.predict(Xq, return_std=True, include_noise=True)
    Xtr = m.train_inputs[0]
    w, z, sf2, noise, beta = [t.detach().numpy() for t in m._natural()]
    # prediction inputs contain all four sources, so the eval-mode re-levelling is the identity here
    assert sorted(set(Xq[:, -1].tolist())) == [0.0, 1.0, 2.0, 3.0]
    p = {"n": Xtr.shape[0], "dq": 10, "dz": 2, "n_combo": 4, "n_noise": 4, "n_mean": 3, "kernel": 0,
         "xq": Xtr[:, :10].numpy(), "y": m.train_targets.numpy(), "level_idx": Xtr[:, -1].numpy().astype(int),
         "noise_idx": Xtr[:, -1].numpy().astype(int), "mean_idx": Xtr[:, -1].numpy().astype(int) - 1}
    h = {"w": w, "z": z, "sigma_f2": float(sf2), "noise": noise, "beta": beta}
    c = {"m": 40, "xq": Xq[:, :10].numpy(), "level_idx": Xq[:, -1].numpy().astype(int),
         "noise_idx": Xq[:, -1].numpy().astype(int), "mean_idx": Xq[:, -1].numpy().astype(int) - 1}
    mu, var = O.predict(p, h, c, include_noise=True)
    y_min, y_std = float(m.y_min), float(m.y_std)
    assert np.max(np.abs(mean.numpy() - (y_min + y_std * mu))) < 1e-8 * max(1.0, abs(y_min) + y_std)
    assert np.max(np.abs(std.numpy() - np.sqrt(var) * y_std)) < 1e-8 * y_std
    mean2 = m.predict(Xq, return_std=False)
    assert torch.equal(mean, mean2)


def test_full_fit_config2_runs_and_improves():
    from gpplus_b200.optim.mll_scipy import MLLObjective
    m, spec, Xte, yte = _c2()
    obj = MLLObjective(m, True, [0, 0])
    f0 = obj.fun(obj.pack_parameters(), False)
    torch.manual_seed(0)
    m.fit(num_restarts=8, bounds=True)
    f1 = obj.fun(obj.pack_parameters(), False)
    assert f1 < f0
    mean = m.predict(Xte[:300], return_std=False)
    assert torch.isfinite(mean).all()


def test_adam_and_continuation_paths_drive_the_same_engine():
    from gpplus_b200.optim import fit_model_continuation, fit_model_torch
    m, spec, _, _ = _c1()
    f, hist = fit_model_torch(m, num_iter=15, num_restarts=0)
    assert np.isfinite(f) and hist[0][-1] < hist[0][0]
    m2, *_ = _c1()
    torch.manual_seed(1)
    nll, history = fit_model_continuation(m2, num_restarts=1, bounds=True)
    assert np.isfinite(nll) and len(history["noise_history"]) >= 1
    assert not m2.likelihood.raw_noise.requires_grad


def test_fused_table_acquisition_matches_slice_by_slice_reference_sequence():
    """BO's candidate-table step (BO_GP_plus.py:183-194): per-source predict(include_noise=False) +
    AF_HF/LF_Engineering + argmax(cat(scores)) versus the single fused predict+acquisition+arg-max pass."""
    from gpplus_b200.bayesian_optimizations import AF_HF_Engineering, AF_LF_Engineering, acquisition_table_argmax
    m, spec, Xte, yte = _c3()
    torch.manual_seed(3)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.1 * torch.randn_like(p))
    rng = np.random.default_rng(11)
    M = 3000
    Xq = rng.standard_normal((M, 10))
    src = rng.integers(0, 4, (M, 1)).astype(float)
    table = np.hstack([Xq, src])
    costs = {"0": 1000.0, "1": 100.0, "2": 10.0, "3": 100.0}
    cost_fun = lambda s: costs[str(int(s))]  # noqa: E731
    ytr = m.y_min + m.y_std * m.train_targets
    best = [float(ytr[m.train_inputs[0][:, -1] == i].max()) for i in range(4)]
    for maximize in (True, False):
        scores = []
        for i in range(4):
            part = torch.tensor(table[table[:, -1] == i])
            mu, sd = m.predict(part, return_std=True, include_noise=False)
            af = AF_HF_Engineering if i == 0 else AF_LF_Engineering
            scores.append(af(best[i], mu.reshape(-1, 1), sd.reshape(-1, 1), part, cost_fun, maximize=maximize))
        ref_scores = torch.cat(scores, dim=0).numpy()
        ref_index = int(np.argmax(ref_scores))
        score, index, order, got = acquisition_table_argmax(m, table, best, [costs[str(i)] for i in range(4)],
                                                            maximize=maximize, return_scores=True)
        assert index == ref_index
        assert np.allclose(got, ref_scores, rtol=1e-9, atol=1e-14)
        assert abs(score - ref_scores[ref_index]) <= 1e-9 * abs(ref_scores[ref_index])
        assert np.array_equal(table[order][:, -1], np.sort(table[:, -1]))


@pytest.mark.parametrize("maker", [_c1, _c2, lambda: _c2("Matern52Kernel"), _c3])
def test_native_objective_equals_torch_path(maker):
    """gpp_objective (float32 cast, transforms, priors, chain rule inside the library) == MLLObjective.fun."""
    from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior
    m, spec, _, _ = maker()
    obj = MLLObjective(m, True, [0, 0])
    assert obj.enable_fast_path()
    torch.manual_seed(1)
    thetas = [obj.pack_parameters()] + [np.clip(_sample_from_prior(m), -6, 4) for _ in range(3)]
    for th in thetas:
        f_ref, g_ref = obj.fun(th)
        f, g = obj.fun_fast(th)
        assert abs(f - f_ref) <= 1e-10 * max(1.0, abs(f_ref))
        assert np.max(np.abs(g - g_ref)) <= 1e-9 * max(1.0, np.max(np.abs(g_ref)))
        assert isinstance(obj.fun_fast(th, False), float)


def test_evaluation_joint_nlpd_matches_dense_oracle():
    """GP_Plus.evaluation (gp_plus.py:889-932): joint NLPD via the two-likelihood identity against the dense
    full-covariance predictive computed on the CPU; MSE / MAE / RRMSE / IS against their definitions."""
    from oracle import gp_oracle as O
    m, spec, Xte, yte = _c3()
    torch.manual_seed(5)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
        m.likelihood.noise_covar.raw_noise.fill_(-4.0)
    Xte, yte = Xte[:40], yte[:40]
    got = m.evaluation(Xte, yte, return_metrics=True)
    # dense reference
    h = m._hyper_numpy()
    xtr = m.train_inputs[0].double()
    cols = m._quant_columns()
    w, z, sf2 = torch.tensor(h["w"]), torch.tensor(h["z"]), h["sigma_f2"]
    noise, beta = torch.tensor(h["noise"]), torch.tensor(h["beta"])
    ltr = torch.as_tensor(m._level_index(xtr, True), dtype=torch.long)
    lte = torch.as_tensor(m._level_index(Xte, False), dtype=torch.long)
    kind = m._quant_kernel().family
    centre = xtr[:, cols].mean(0, keepdim=True)
    Ktt = O.covariance(xtr[:, cols], ltr, xtr[:, cols], ltr, w, z, sf2, kind, centre=centre)
    Kst = O.covariance(Xte[:, cols], lte, xtr[:, cols], ltr, w, z, sf2, kind, centre=centre)
    Kss = O.covariance(Xte[:, cols], lte, Xte[:, cols], lte, w, z, sf2, kind, centre=centre)
    ntr = noise[torch.as_tensor(m._noise_index(xtr), dtype=torch.long)]
    nte = noise[torch.as_tensor(m._noise_index(Xte), dtype=torch.long)]
    mtr = O._mean_vector(torch.as_tensor(m._mean_index(xtr), dtype=torch.long), beta, len(beta))
    mte = O._mean_vector(torch.as_tensor(m._mean_index(Xte), dtype=torch.long), beta, len(beta))
    L = torch.linalg.cholesky(Ktt + torch.diag(ntr))
    alpha = torch.cholesky_solve((m.train_targets.double() - mtr).unsqueeze(-1), L).squeeze(-1)
    mu = mte + Kst @ alpha
    V = torch.linalg.solve_triangular(L, Kst.T, upper=False)
    Sig = Kss - V.T @ V + torch.diag(nte)
    y_sc = (yte.double() - m.y_min) / m.y_std
    dist = torch.distributions.MultivariateNormal(mu, covariance_matrix=Sig)
    nlpd = float(-dist.log_prob(y_sc) / y_sc.shape[0])
    assert abs(got["NLL"] - nlpd) <= 1e-7 * max(1.0, abs(nlpd))
    sd = torch.sqrt(torch.diagonal(Sig))
    mse = float(torch.mean((mu - y_sc) ** 2) * m.y_std ** 2)
    assert abs(got["MSE"] - mse) <= 1e-8 * mse
    assert abs(got["RRMSE"] - math.sqrt(mse / float(torch.var(yte.double())))) <= 1e-8
    lo, up = mu - 2 * sd, mu + 2 * sd
    isc = ((up - lo) + (y_sc > up) * 40.0 * (y_sc - up) + (y_sc < lo) * 40.0 * (lo - y_sc)).mean() * abs(float(m.y_std))
    assert abs(got["IS"] - float(isc)) <= 1e-7 * float(isc)


def test_sobol_indices_of_an_additive_function():
    """GP_Plus.Sobol (gp_plus.py:1148-1224) on y = 2 x0 + x1^2 + level effect: first-order and total indices
    coincide (no interactions) and sum to one."""
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.optim import fit_model_scipy
    rng = np.random.default_rng(0)
    n = 240
    X = np.hstack([rng.uniform(-1, 1, (n, 3)), rng.integers(0, 3, (n, 1)).astype(float)])
    y = 2.0 * X[:, 0] + X[:, 1] ** 2 + 0.5 * X[:, 3]
    m = GP_Plus(torch.tensor(X), torch.tensor(y), qual_dict={3: 3}, dtype=torch.float64)
    torch.manual_seed(0)
    fit_model_scipy(m, num_restarts=4, bounds=True)
    with pytest.warns(UserWarning):
        S, ST = m.Sobol(N=4096)
    assert S.shape == (1, 4) and ST.shape == (1, 4)
    assert abs(S.sum() - 1.0) < 0.05 and np.all(np.abs(S - ST) < 0.05)
    assert abs(S[0, 2]) < 0.02            # x2 is inert
    assert S[0, 0] > S[0, 1] > 0.0        # var(2 x0) = 4/3 > var(x1^2) = 4/45


@pytest.mark.parametrize("maker", [_c3, lambda: _c2("Matern52Kernel")])
def test_device_side_table_preparation_equals_host_preparation(maker):
    """prepare_candidate_table_on_device (ordering, gather and level ranking as torch device ops) must hand the
    engine exactly the arrays of the host preparation, including the per-slice setlevels quirk."""
    from gpplus_b200.bayesian_optimizations import (prepare_candidate_table, prepare_candidate_table_on_device,
                                                    score_prepared)
    m, spec, Xte, yte = maker()
    rng = np.random.default_rng(5)
    M = 5000
    xtr = m.train_inputs[0].double().numpy()
    table = xtr[rng.integers(0, xtr.shape[0], M)].copy()
    qcols = m._quant_columns()
    table[:, qcols] += 0.1 * rng.standard_normal((M, len(qcols)))
    n_src = 4
    if spec.get("m_gp") != "multiple_constant":   # mixed-variable model: last column is a 5-level categorical
        table[:, -1] = rng.integers(0, n_src, M)
    host = prepare_candidate_table(m, table, n_src)
    dev = prepare_candidate_table_on_device(m, table, n_src, 0)
    assert np.array_equal(host["order"], dev["order"]) and host["lo"] == dev["lo"] and host["count"] == dev["count"]
    assert np.array_equal(host["xq"], dev["xq"].cpu().numpy())
    assert np.array_equal(host["cost_idx"], dev["cost_idx"].cpu().numpy())
    for key in ("level_idx", "mean_idx"):
        assert (host[key] is None) == (dev[key] is None)
        if host[key] is not None:
            assert np.array_equal(host[key], dev[key].cpu().numpy())
    best, costs = [0.3, 0.2, 0.1, 0.0], [1000.0, 100.0, 10.0, 100.0]
    a = score_prepared(m, host, best, costs, maximize=False, return_scores=True)
    b = score_prepared(m, dev, best, costs, maximize=False, return_scores=True)
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_plain_gpr_with_string_kernels():
    """GPR (models/gpregression.py:38-175) built from a kernel name: likelihood against the natural-parameter oracle,
    prediction interpolates a smooth function."""
    from gpplus_b200.models import GPR
    from oracle import gp_oracle as O
    rng = np.random.default_rng(8)
    n = 150
    X = rng.uniform(-1, 1, (n, 3))
    y = np.sin(2 * X[:, 0]) + X[:, 1] ** 2 - 0.5 * X[:, 2]
    for name, kind, wfun in (("RBFKernel", 0, lambda ls: 0.5 / ls ** 2), ("Matern52Kernel", 2, lambda ls: 1.0 / ls ** 2)):
        m = GPR(torch.tensor(X), torch.tensor(y), name, noise_indices=[], lb_noise=1e-8)
        m = m.double() if hasattr(m, "double") else m
        with torch.no_grad():
            m.covar_module.base_kernel.raw_lengthscale.fill_(0.3)
            m.likelihood.noise_covar.raw_noise.fill_(-7.0)
        lm = float(m.log_marginal().detach())
        ls = float(np.exp(0.3))
        ys = (y - y.min()) / (y.max() - y.min())
        p = {"n": n, "dq": 3, "dz": 0, "n_combo": 0, "n_noise": 1, "n_mean": 0, "kernel": kind, "xq": X, "y": ys,
             "level_idx": None, "noise_idx": None, "mean_idx": None}
        h = {"w": np.full(3, wfun(ls)), "z": None, "sigma_f2": float(m.covar_module.outputscale),
             "noise": np.array([1e-8 + np.exp(-7.0)]), "beta": None}
        ref = O.mll(p, h, want_grad=False)
        assert abs(-lm - ref["nll"]) <= 1e-8 * abs(ref["nll"])
        Xs = rng.uniform(-1, 1, (50, 3))
        mu, sd = m.predict(torch.tensor(Xs), return_std=True)
        truth = np.sin(2 * Xs[:, 0]) + Xs[:, 1] ** 2 - 0.5 * Xs[:, 2]
        assert float(np.sqrt(np.mean((mu.numpy() - truth) ** 2))) < 0.15 and bool(torch.all(sd > 0))


@pytest.mark.parametrize("kwargs", [dict(fix_noise=True, fix_noise_val=1e-4), dict(m_gp="single_zero"),
                                    dict(fixed_length_scale=True, fixed_length_scale_val=torch.tensor([[0.5, 0.5, 0.5]])),
                                    dict(lb_noise=1e-6, quant_correlation_class="Matern32Kernel")])
def test_constructor_variants_fit_through_the_native_objective(kwargs):
    """Frozen noise / zero mean / frozen lengthscales / other noise floor: the layout handed to gpp_objective has
    frozen blocks (negative offsets); value and gradient must still equal the torch path, and a fit must run."""
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.optim.mll_scipy import MLLObjective, fit_model_scipy
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (120, 3))
    y = np.sin(2 * X[:, 0]) + X[:, 1] ** 2 - 0.5 * X[:, 2] + 0.01 * rng.standard_normal(120)
    m = GP_Plus(torch.tensor(X), torch.tensor(y), dtype=torch.float64, **kwargs)
    obj = MLLObjective(m, True, [0, 0])
    assert obj.enable_fast_path()
    th = obj.pack_parameters() + 0.1 * rng.standard_normal(obj.pack_parameters().shape[0])
    f_ref, g_ref = obj.fun(th)
    f, g = obj.fun_fast(th)
    assert abs(f - f_ref) <= 1e-10 * max(1.0, abs(f_ref)) and np.max(np.abs(g - g_ref)) <= 1e-9 * max(1.0, np.max(np.abs(g_ref)))
    torch.manual_seed(0)
    res, best = fit_model_scipy(m, num_restarts=7, bounds=True)
    assert np.isfinite(best) and len(res) == 8 and best == min(r.fun for r in res if not isinstance(r, Exception))
    mu, sd = m.predict(torch.tensor(X), return_std=True)
    assert bool(torch.all(torch.isfinite(mu))) and bool(torch.all(sd > 0))
    assert float(torch.sqrt(torch.mean((mu - torch.tensor(y)) ** 2))) < 0.35   # y spans about +-1.5


def test_loocv_rrmse_matches_the_oracle_inverse_diagonal():
    """loocv_rrmse (optim/mll_noise_continuation.py:28-42): rms of alpha_i / (K_y^-1)_ii, with alpha and diag(K_y^-1)
    fetched from the device (gpp_fetch which = 3 / 4) against the oracle's dense inverse."""
    from gpplus_b200.optim.mll_noise_continuation import loocv_rrmse
    from gpplus_b200.optim.mll_scipy import MLLObjective
    from oracle import gp_oracle as O
    for builder in (_c1, _c3):
        m, spec, _, _ = builder()
        obj = MLLObjective(m, True, [0, 0])
        obj._load((obj.pack_parameters() + 0.1).astype(np.float32).astype(np.float64))
        got = loocv_rrmse(m)
        eng = m._get_engine()
        x = m.train_inputs[0]
        cols = m._quant_columns()
        with torch.no_grad():
            w, zt, sf2, noise, beta = m._natural()
        n_mean, _ = m._mean_layout()
        prob = {"n": x.shape[0], "dq": len(cols), "dz": eng.dz, "n_combo": eng.n_combo, "n_noise": eng.n_noise,
                "n_mean": n_mean, "kernel": eng.kernel, "xq": x[:, cols].double().numpy(),
                "y": m.train_targets.double().numpy(), "level_idx": m._level_index(x, True),
                "noise_idx": m._noise_index(x), "mean_idx": m._mean_index(x)}
        hyp = {"w": w.double().numpy(), "z": zt.double().numpy() if eng.dz > 0 else None, "sigma_f2": float(sf2),
               "noise": noise.double().numpy(), "beta": beta.double().numpy() if n_mean > 0 else None}
        ref = O.mll(prob, hyp, want_grad=False, return_mats=True)
        want = float(np.sqrt(np.mean((ref["alpha"] / np.diag(ref["Kinv"])) ** 2)))
        assert abs(got - want) <= 1e-8 * abs(want), (got, want)
        np.testing.assert_allclose(eng.fetch("Kinv_diag"), np.diag(ref["Kinv"]), rtol=1e-8)
        m.release_engine()


def test_sobol_indices_match_the_same_design_on_oracle_predictions():
    """GP_Plus.Sobol (models/gp_plus.py:1148-1224): Saltelli estimators on engine predictions equal the same
    estimators evaluated on the CPU oracle's predictive mean over the same A / B / AB_i design."""
    from scipy.stats.qmc import Sobol as _Sobol
    from oracle import gp_oracle as O
    import warnings
    m, spec, _, _ = _c1()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S, ST = m.Sobol(N=512)
    eng = m._get_engine()
    x = m.train_inputs[0].double()
    p = x.shape[1]
    seq = torch.from_numpy(_Sobol(d=2 * p, scramble=False).random(513)[1:])
    mins, maxs = x.min(dim=0)[0], x.max(dim=0)[0]
    A = mins + (maxs - mins) * seq[:, p:]
    B = mins + (maxs - mins) * seq[:, :p]
    with torch.no_grad():
        w, zt, sf2, noise, beta = m._natural()
    prob = {"n": x.shape[0], "dq": p, "dz": 0, "n_combo": 0, "n_noise": 1, "n_mean": 1, "kernel": eng.kernel,
            "xq": x.numpy(), "y": m.train_targets.double().numpy(), "level_idx": None, "noise_idx": None,
            "mean_idx": None}
    hyp = {"w": w.double().numpy(), "z": None, "sigma_f2": float(sf2), "noise": noise.double().numpy(),
           "beta": beta.double().numpy()}

    def f(X):
        mu, _ = O.predict(prob, hyp, {"m": X.shape[0], "xq": X.numpy(), "level_idx": None, "noise_idx": None,
                                     "mean_idx": None})
        return (float(m.y_min) + float(m.y_std) * mu).reshape(-1, 1)

    FA, FB = f(A), f(B)
    S_ref, ST_ref = np.zeros(p), np.zeros(p)
    for i in range(p):
        ABi = A.clone()
        ABi[:, i] = B[:, i]
        Fi = f(ABi)
        S_ref[i] = np.sum(FB * (Fi - FA)) / 512
        ST_ref[i] = np.sum((FA - Fi) ** 2) / (2 * 512)
    varY = np.var(np.concatenate([FA, FB]))
    np.testing.assert_allclose(S.reshape(-1), S_ref / varY, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(ST.reshape(-1), ST_ref / varY, rtol=1e-6, atol=1e-10)
    m.release_engine()


def test_split_objective_equals_the_blocking_call_bitwise():
    """gpp_objective_enqueue + gpp_objective_collect on several handles in flight == gpp_objective, bit for bit."""
    from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior
    m, spec, _, _ = _c2()
    obj = MLLObjective(m, True, [0, 0])
    assert obj.enable_fast_path()
    layout = obj._fast.layout_spec()
    torch.manual_seed(5)
    thetas = [np.clip(_sample_from_prior(m), -5, 3) for _ in range(6)]
    ref = m._new_engine(0)
    ref.set_theta_layout(layout)
    want = [ref.objective(t, True) for t in thetas]
    engines = [m._new_engine(0) for _ in thetas]
    try:
        for e in engines:
            e.set_theta_layout(layout)
        for rep in range(2):
            for e, t in zip(engines, thetas):
                e.objective_enqueue(t, True)
            for e, (f, g) in zip(engines, want):
                grad = np.empty_like(g)
                val = e.objective_collect(grad)
                assert val == f and np.array_equal(grad, g)
        with pytest.raises(ValueError):
            engines[0].objective_collect(np.empty_like(want[0][1]))  # nothing enqueued
    finally:
        for e in engines + [ref]:
            e.close()
    m.release_engine()


def test_lockstep_driver_reproduces_the_threaded_fit_restart_by_restart(monkeypatch):
    """The single-thread lock-step driver (optim/_lockstep.py) and the one-thread-per-restart path run the same
    scipy L-BFGS-B state machines on the same objective: every restart ends at the same point after the same
    number of iterations and evaluations."""
    from gpplus_b200.optim.mll_scipy import _sample_from_prior, fit_model_scipy
    m, spec, Xte, yte = _c2()
    torch.manual_seed(11)
    starts = [_sample_from_prior(m) for _ in range(10)]
    monkeypatch.setenv("GPPLUS_LOCKSTEP", "0")
    out_t, best_t = fit_model_scipy(m, theta0_list=[s.copy() for s in starts], bounds=True)
    monkeypatch.setenv("GPPLUS_LOCKSTEP", "1")
    m2, *_ = _c2()
    out_l, best_l = fit_model_scipy(m2, theta0_list=[s.copy() for s in starts], bounds=True)
    assert best_l == best_t
    for a, b in zip(out_t, out_l):
        if isinstance(a, Exception) or isinstance(b, Exception):
            assert type(a) is type(b)
            continue
        assert a.fun == b.fun and a.nit == b.nit and a.nfev == b.nfev and a.status == b.status
        assert np.array_equal(a.x, b.x) and a.message == b.message
    m.release_engine()
    m2.release_engine()
