"""Synthetic workloads of bench.py, shared with tests/ and tests/golden/make_golden.py.

C4 = BASELINE.json configs[3]: synthetic exact GP, N=16384, D=10, Matern-5/2, FP64, one constant mean,
homoskedastic noise with lb_noise = 1e-8.  Evaluation points: theta_init (every raw parameter 0) and the 65
prior draws of a 64-restart fit with ``torch.manual_seed(0)`` (Normal(-3,3) on omega, log-half-horseshoe on the
raw noise, LogNormal on the output scale, Normal(0,1) on the mean) -- exactly the starts ``fit_model_scipy``
would generate (reference optim/mll_scipy.py:130-138, :281-285).
"""
import numpy as np

D = 10
N_HEADLINE = 16384
N_PRIOR_DRAWS = 65


def c4_workload(n):
    """Sobol(d=10, seed=0) scaled to the wing bounds, y = wing(X) + N(0, 0.5^2) with numpy seed 0; X standardised
    per column (preprocessing/normalizeX.py:31-36); y is min-max scaled by the model (gpregression.py:67-69)."""
    from scipy.stats.qmc import Sobol, scale
    from gpplus_b200.test_functions.analytical import WING_BOUNDS, wing_weight
    X = scale(Sobol(d=D, seed=0).random(n), l_bounds=WING_BOUNDS[0], u_bounds=WING_BOUNDS[1])
    rng = np.random.RandomState(0)
    y = wing_weight(X) + 0.5 * rng.randn(n)
    Xs = (X - X.mean(0)) / X.std(0)
    return Xs, y


def c4_model(n):
    """The GP_Plus model of the C4 workload (host object; no engine is created until it is evaluated)."""
    import torch
    from gpplus_b200.models import GP_Plus
    X, y = c4_workload(n)
    return GP_Plus(torch.from_numpy(X), torch.from_numpy(y), dtype=torch.float64,
                   quant_correlation_class="Matern52Kernel")


def c4_theta_points(model):
    """[theta_init, prior draw 1, ..., prior draw 65] in raw-parameter (theta) space, float64 arrays of length 13.
    The draws do not depend on the training data (only on the priors), so every N shares the same list."""
    import torch
    from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior
    obj = MLLObjective(model, True, [0, 0])
    base = obj.pack_parameters() * 0.0
    state = torch.random.get_rng_state()
    torch.manual_seed(0)
    draws = [_sample_from_prior(model).astype(np.float64) for _ in range(N_PRIOR_DRAWS)]
    torch.random.set_rng_state(state)
    return [base] + draws


def c4_natural(theta, lb_noise=1e-8):
    """raw [noise, outputscale, omega_1..10, mean] -> natural parameters of the C ABI (Matern-5/2): the float32
    cast of theta (mll_scipy.py:97), noise = lb + exp(raw), sigma_f^2 = softplus(raw), w_d = 2 * 10^omega_d
    (lengthscale 2^-1/2 10^(-omega/2), gp_plus.py:252), beta = raw."""
    th = np.asarray(theta, dtype=np.float32).astype(np.float64)
    sp = float(np.log1p(np.exp(th[1]))) if th[1] < 30 else float(th[1])
    return {"w": 2.0 * 10.0 ** th[2:12], "z": None, "sigma_f2": sp,
            "noise": np.array([lb_noise + np.exp(th[0])]), "beta": th[12:13].copy()}


def c4_oracle_problem(n):
    """The C4 workload as a problem dictionary of oracle/gp_oracle.py (y already min-max scaled)."""
    X, y = c4_workload(n)
    ys = (y - y.min()) / (y.max() - y.min())
    return {"n": n, "dq": D, "dz": 0, "n_combo": 0, "n_noise": 1, "n_mean": 1, "kernel": 2, "xq": X, "y": ys,
            "level_idx": None, "noise_idx": None, "mean_idx": None}
