"""Host-side mirror of the GP+ interface (no GPU needed): parameter / prior names and order, theta
packing with the float32 cast, priors, bounds, level indexing, restart sharding over 2 gloo ranks."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from gpplus_b200.models import GP_Plus
from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior, fit_model_scipy, get_bounds
from gpplus_b200.preprocessing import setlevels, standard, train_test_split_normalizeX
from gpplus_b200.priors import LogHalfHorseshoePrior, MollifiedUniformPrior
from gpplus_b200.test_functions import borehole, borehole_mixed_variables, multi_fidelity_wing, wing
from gpplus_b200.utils import inv_softplus, set_seed, softplus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mixed_model(dtype=torch.float64, **kw):
    set_seed(4)
    qual_dict = {0: 5, 5: 5}
    X, y = borehole_mixed_variables(n=400, qual_dict=qual_dict, random_state=4)
    Xtr, Xte, ytr, yte = train_test_split_normalizeX(X, y, test_size=0.75, qual_dict=qual_dict)
    return GP_Plus(Xtr, ytr, qual_dict=qual_dict, dtype=dtype, **kw), Xtr, ytr, Xte


def test_parameter_and_prior_order_mixed():
    m, Xtr, _, _ = _mixed_model()
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names == ["latent[0, 5]", "likelihood.noise_covar.raw_noise", "covar_module.raw_outputscale",
                     "covar_module.base_kernel.kernels.1.raw_lengthscale", "mean_module.constant"]
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert frozen == ["covar_module.base_kernel.kernels.0.raw_lengthscale"]
    pri = [n for n, *_ in m.named_priors()]
    assert pri == ["latent_prior_latent[0, 5]", "likelihood.noise_prior", "covar_module.outputscale_prior",
                   "covar_module.base_kernel.kernels.1.lengthscale_prior", "mean_module.mean_prior"]
    obj = MLLObjective(m, True, [0, 0])
    assert obj.pack_parameters().shape == (29,)  # 2*10 + 1 + 1 + 6 + 1
    assert float(m.covar_module.base_kernel.kernels[0].lengthscale) == pytest.approx(1.0)


def test_multifidelity_model_layout():
    set_seed(4)
    X, y = multi_fidelity_wing(n={"0": 20, "1": 30, "2": 30, "3": 30}, random_state=4)
    Xtr, Xte, ytr, yte = train_test_split_normalizeX(X, y, test_size=0.2, qual_dict={10: 4}, stratify=X[:, -1])
    m = GP_Plus(Xtr, ytr, qual_dict={10: 4}, multiple_noise=True, m_gp="multiple_constant", dtype=torch.float64)
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names == ["latent[10]", "likelihood.noise_covar.raw_noise", "covar_module.raw_outputscale",
                     "covar_module.base_kernel.kernels.1.raw_lengthscale", "mean_module_1.constant",
                     "mean_module_2.constant", "mean_module_3.constant"]
    assert m.likelihood.noise_covar.raw_noise.shape == (4,)
    assert MLLObjective(m, True, [0, 0]).pack_parameters().shape == (26,)  # 8 + 4 + 1 + 10 + 3
    mi = m._mean_index(Xtr)
    assert mi.min() == -1 and mi.max() == 2
    assert np.array_equal(m._noise_index(Xtr), Xtr[:, -1].numpy().astype(np.int32))


def test_borehole_model_layout_and_transforms():
    set_seed(1245)
    X, y = borehole(n=1000, random_state=12345)
    Xtr, Xte, ytr, yte = train_test_split_normalizeX(X, y, test_size=0.8)
    m = GP_Plus(Xtr, ytr, dtype=torch.float64)
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names == ["likelihood.noise_covar.raw_noise", "covar_module.raw_outputscale",
                     "covar_module.base_kernel.raw_lengthscale", "mean_module.constant"]
    with torch.no_grad():
        m.covar_module.base_kernel.raw_lengthscale.fill_(0.5)
        w, z, sf2, noise, beta = m._natural()
    assert torch.allclose(w, torch.full((8,), 10.0 ** 0.5, dtype=torch.float64))   # Rough_RBF: w = 10^omega
    assert float(sf2) == pytest.approx(math.log(2.0))                               # softplus(0)
    assert float(noise) == pytest.approx(1e-8 + 1.0)                                # lb + exp(0)
    m2 = GP_Plus(Xtr, ytr, dtype=torch.float64, quant_correlation_class="Matern52Kernel")
    with torch.no_grad():
        m2.covar_module.base_kernel.raw_lengthscale.fill_(0.5)
        assert torch.allclose(m2._natural()[0], torch.full((8,), 2 * 10.0 ** 0.5, dtype=torch.float64))
    m3 = GP_Plus(Xtr, ytr, dtype=torch.float64, quant_correlation_class="RBFKernel")
    with torch.no_grad():
        m3.covar_module.base_kernel.raw_lengthscale.fill_(0.5)
        assert torch.allclose(m3._natural()[0], torch.full((8,), 0.5 * math.exp(-1.0), dtype=torch.float64))
    assert type(next(p for n, _, p, _, _ in m3.named_priors() if "lengthscale" in n)).__name__ == \
        "MollifiedUniformPrior"


def test_unsupported_variants_raise_instead_of_silently_differing():
    X = torch.randn(20, 3, dtype=torch.float64)
    y = torch.randn(20, dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        GP_Plus(X, y, calibration_type="probabilistic")
    with pytest.raises(NotImplementedError):
        GP_Plus(X, y, calibration_id=[1])
    with pytest.raises(NotImplementedError):
        GP_Plus(X, y, m_gp="single_polynomial-d2")
    with pytest.raises(RuntimeError):
        GP_Plus(X, y, quant_correlation_class="Matern12Kernel")
    with pytest.raises(ValueError):
        GP_Plus(X, y, quant_correlation_class="Cubic")


def test_theta_is_cast_to_float32_like_the_reference():
    m, *_ = _mixed_model()
    obj = MLLObjective(m, True, [0, 0])
    x = obj.pack_parameters() + 1e-9 + 0.123456789
    obj._load(x)
    got = obj.pack_parameters()
    assert np.array_equal(got, x.astype(np.float32).astype(np.float64))


def test_level_index_matches_string_lookup():
    m, Xtr, _, _ = _mixed_model()
    idx = m._level_index(Xtr, True)
    want = [m.perm_dict[0][str(row.tolist())] for row in Xtr[:, [0, 5]].type(torch.int64)]
    assert idx.tolist() == want
    zeta = m.zeta[0]
    assert zeta.shape == (25, 10) and int(zeta.sum()) == 50
    assert torch.equal(m.transform_categorical(Xtr[:7, [0, 5]], m.perm_dict[0], zeta), zeta[idx[:7]])
    bad = Xtr.clone()
    bad[0, 0] = 7
    with pytest.raises(ValueError):
        m._level_index(bad, True)


def test_sample_from_prior_is_seeded_and_ordered():
    m, *_ = _mixed_model()
    torch.manual_seed(0)
    a = _sample_from_prior(m)
    torch.manual_seed(0)
    b = _sample_from_prior(m)
    assert a.shape == (29,) and np.array_equal(a, b)
    assert a[21] > 0  # LogNormal draw used as the RAW outputscale start, as in the reference
    lo, hi = get_bounds(MLLObjective(m, True, [0, 0]), a)
    assert lo.shape == (29,) and np.all(np.isinf(lo[:21])) and np.all(lo[21:28] == -10) and lo[28] == -1.5
    assert np.all(hi[21:28] == 3) and hi[28] == 1.5


def test_prior_log_densities():
    v = torch.tensor([-6.0, -2.0], dtype=torch.float64)
    hs = LogHalfHorseshoePrior(0.01, 1e-8)
    want = torch.log(torch.log(1 + 3 * (0.01 / (1e-8 + torch.exp(v))) ** 2)) + v
    assert torch.allclose(hs.log_prob(v), want, rtol=1e-14)
    assert float(hs.expand([3]).lb[0]) == pytest.approx(1e-6)  # expand drops lb (reference quirk)
    mu = MollifiedUniformPrior(math.log(0.1), math.log(10))
    inside = mu.log_prob(torch.tensor([0.0, 1.0, -2.0], dtype=torch.float64))
    assert torch.allclose(inside, inside[0].expand(3))
    c = -math.log(1 + (math.log(10) - math.log(0.1)) / (math.sqrt(2 * math.pi) * 0.1))
    assert float(inside[0]) == pytest.approx(c - math.log(0.1) - 0.5 * math.log(2 * math.pi), rel=1e-12)
    out = mu.log_prob(torch.tensor([math.log(10) + 0.2], dtype=torch.float64))
    assert float(out) == pytest.approx(float(inside[0]) - 0.5 * (0.2 / 0.1) ** 2, rel=1e-12)
    torch.manual_seed(1)
    s = mu.expand([1000]).sample()
    assert s.min() >= math.log(0.1) and s.max() < math.log(10)
    x = torch.tensor([0.3, 2.0, 30.0], dtype=torch.float64)
    assert torch.allclose(inv_softplus(softplus(x)), x, rtol=1e-12)


def test_preprocessing_and_generators():
    X, y = wing(n=64, random_state=0)
    assert X.shape == (64, 10) and y.shape == (64,)
    Xs, mean, std = standard(torch.tensor(X.copy()), {})
    assert torch.allclose(Xs.mean(0), torch.zeros(10, dtype=torch.float64), atol=1e-12)
    lv = setlevels(np.array([[3.5, 1.0], [1.5, 1.0], [3.5, 2.0]]), qual_index=[0])
    assert lv[:, 0].tolist() == [1.0, 0.0, 1.0]
    Xb, yb = borehole(n=32, random_state=1)
    assert np.all(yb > 0)


def test_fit_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpplus_b200._engine import EngineError
    m, *_ = _mixed_model()
    with pytest.raises(EngineError):
        fit_model_scipy(m, num_restarts=0)


WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "gp-plus_b200"))
import numpy as np, torch, torch.distributed as dist
from scipy.optimize import OptimizeResult
from gpplus_b200 import parallel
from gpplus_b200._engine import NotPSDError
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
thetas = [np.full(3, float(i)) for i in range(5)] if rank == 0 else [np.zeros(3) for _ in range(5)]
thetas = parallel.broadcast_theta_list(thetas)
assert all(np.all(t == i) for i, t in enumerate(thetas))
mine = parallel.shard_indices(5)
assert mine == list(range(rank, 5, 2))
# cross-rank work queue: every index is claimed exactly once, whatever the interleaving
import threading, time
q = parallel.RestartQueue(23)
claimed = []
def _drain():
    while True:
        i = q.claim()
        if i is None:
            break
        claimed.append(i)
        time.sleep(0.001 * (1 + rank))
ts = [threading.Thread(target=_drain) for _ in range(3)]
[t.start() for t in ts]
[t.join() for t in ts]
flags = torch.zeros(23)
flags[claimed] = 1.0
assert len(claimed) == len(set(claimed))
dist.all_reduce(flags)
assert bool((flags == 1.0).all()), flags
q2 = parallel.RestartQueue(1)   # a second queue uses a fresh counter
got = q2.claim()
total = torch.tensor([0.0 if got is None else 1.0])
dist.all_reduce(total)
assert float(total) == 1.0
# ownership by claim order instead of i %% world: rank 0 owns {0, 1, 4}, rank 1 owns {2, 3}
mine = [0, 1, 4] if rank == 0 else [2, 3]
local = {}
for i in mine:
    local[i] = NotPSDError("x") if i == 3 else OptimizeResult(x=thetas[i] + 0.5, fun=10.0 - i, nit=i, nfev=2 * i,
                                                             njev=2 * i, status=0, success=True)
out = parallel.gather_restarts(local, 5, 3)
assert len(out) == 5 and isinstance(out[3], NotPSDError)
funs = [np.inf if isinstance(r, Exception) else r.fun for r in out]
assert funs == [10.0, 9.0, 8.0, np.inf, 6.0] and int(np.argmin(funs)) == 4
assert np.all(out[4].x == 4.5) and out[2].nfev == 4
lo, hi = parallel.shard_range(11)
assert (lo, hi) == ((0, 6) if rank == 0 else (6, 11))
s, i = parallel.global_argmax(1.0 if rank == 0 else 1.0, 7 if rank == 0 else 3)
assert (s, i) == (1.0, 3)
s, i = parallel.global_argmax(2.0 if rank == 0 else 1.5, 7 if rank == 0 else 3)
assert (s, i) == (2.0, 7)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_restart_and_candidate_sharding_world2_gloo(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    for r, pr in enumerate(procs):
        out, _ = pr.communicate(timeout=240)
        assert pr.returncode == 0, out.decode()
        assert ("rank %d ok" % r) in out.decode()


def test_acquisition_functions_match_the_oracle_formulas():
    """AF_*_Engineering (AFs.py:102-159) against the oracle's acquisition(); AF_LF/HF/EI point objectives
    (AFs.py:1-99) against the same formulas through a stub model."""
    from gpplus_b200.bayesian_optimizations import AF_EI, AF_HF, AF_HF_Engineering, AF_LF, AF_LF_Engineering
    from oracle import gp_oracle as O
    rng = np.random.default_rng(5)
    m = 40
    mean = torch.tensor(rng.standard_normal((m, 1)))
    std = torch.tensor(0.1 + rng.random((m, 1)))
    xval = torch.tensor(np.hstack([rng.standard_normal((m, 3)), rng.integers(0, 3, (m, 1)).astype(float)]))
    costs = {"0": 1000.0, "1": 100.0, "2": 10.0}
    cost_fun = lambda s: costs[str(int(s))]  # noqa: E731
    cvec = np.array([cost_fun(s) for s in xval[:, -1]])
    for maximize in (True, False):
        for best_f in (0.3, -0.7):
            hf = AF_HF_Engineering(best_f, mean, std, xval, cost_fun, maximize=maximize, si=0.05)
            lf = AF_LF_Engineering(best_f, mean, std, xval, cost_fun, maximize=maximize, si=0.05)
            assert hf.shape == (m,) and lf.shape == (m,)
            ref_hf = O.acquisition(mean.reshape(-1), std.reshape(-1), 0, best_f, cvec, maximize=maximize, si=0.05)
            ref_lf = O.acquisition(mean.reshape(-1), std.reshape(-1), 1, best_f, cvec, maximize=maximize, si=0.05)
            assert np.allclose(hf.numpy(), ref_hf, rtol=1e-13, atol=1e-15)
            assert np.allclose(lf.numpy(), ref_lf, rtol=1e-13, atol=1e-15)

    class Stub:
        def predict(self, x, return_std=True, include_noise=True):
            assert include_noise and x.shape == (1, 4)
            self.seen = x.clone()
            return torch.tensor([0.4]), torch.tensor([0.2])

    stub = Stub()
    xmean, xstd = np.array([1.0, 2.0, 3.0]), np.array([2.0, 4.0, 8.0])
    sample = np.array([3.0, 6.0, 11.0, 2.0])
    for fn, kind in ((AF_HF, 0), (AF_LF, 1), (AF_EI, 2)):
        val = fn(sample, 0.1, stub, xmean, xstd, cost_fun, maximize=False)
        ref = -O.acquisition(np.array([0.4]), np.array([0.2]), kind, 0.1, np.array([10.0]), maximize=False)
        assert np.allclose(np.asarray(val).reshape(-1), ref, rtol=1e-13)
        assert np.allclose(stub.seen.numpy(), [[1.0, 1.0, 1.0, 2.0]])


@pytest.mark.parametrize("kwargs", [
    dict(), dict(quant_correlation_class="Matern52Kernel"), dict(quant_correlation_class="RBFKernel"),
    dict(qual=True), dict(qual=True, multiple_noise=True, m_gp="multiple_constant"), dict(qual=True, fix_noise=True),
    dict(m_gp="single_zero"),
])
def test_closed_form_host_objective_equals_torch_path(kwargs):
    """optim/_fast_objective.py (the layout gpp_objective evaluates natively) against MLLObjective.fun's torch
    path -- float32 cast, transforms, every prior family, chain rule -- through a stub engine, on prior draws."""
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.optim import _fast_objective as FO
    from gpplus_b200.optim.mll_scipy import MLLObjective, _sample_from_prior
    rng = np.random.default_rng(2)
    n = 40
    kw = dict(kwargs)
    qual = kw.pop("qual", False)
    X = np.hstack([rng.standard_normal((n, 3)), rng.integers(0, 3, (n, 1)).astype(float)])
    y = np.sin(X[:, 0]) + 0.3 * X[:, 3] + 0.05 * rng.standard_normal(n)
    if qual:
        kw["qual_dict"] = {3: 3}
    m = GP_Plus(torch.tensor(X), torch.tensor(y), dtype=torch.float64, **kw)
    obj = MLLObjective(m, True, [0, 0])
    fast = FO.build(m, True, [0, 0])
    assert fast is not None
    assert FO.self_check(obj, fast, trials=3, tol=1e-12)
    spec = fast.layout_spec()
    assert spec["p"] == obj.pack_parameters().shape[0]
    table = m._latent_table()
    n_mean, _ = m._mean_layout()
    stub = FO._StubEngine(len(m._quant_columns()), 0 if table is None else int(table.shape[1]),
                          0 if table is None else int(table.shape[0]),
                          int(m.likelihood.noise_covar.raw_noise.numel()), n_mean)
    assert FO.self_check(obj, fast, trials=3, tol=1e-12)   # second call: answered from the per-structure cache
    assert len(m.__dict__["_fast_check_cache"]) == 1
    torch.manual_seed(3)
    with m.engine_override(stub):
        assert m._get_engine() is stub
        for k in range(20):
            th = _sample_from_prior(m) * (1.0 if k < 15 else 2.5)
            f_ref, g_ref = obj.fun(th)
            f, g = fast.fun(th, stub.mll_grad)
            if np.isfinite(f_ref):
                assert abs(f - f_ref) <= 1e-12 * max(1.0, abs(f_ref))
                assert np.max(np.abs(g - g_ref)) <= 1e-11 * max(1.0, np.max(np.abs(g_ref)))
    assert getattr(m, "_engine_stub", None) is None and "_get_engine" not in m.__dict__   # nothing left patched


def test_host_objective_falls_back_for_unsupported_models():
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.optim import _fast_objective as FO
    rng = np.random.default_rng(0)
    X = np.hstack([rng.standard_normal((30, 2)), rng.integers(0, 3, (30, 1)).astype(float)])
    y = X[:, 0] + 0.1 * X[:, 2]
    m = GP_Plus(torch.tensor(X), torch.tensor(y), qual_dict={2: 3}, dtype=torch.float64, NN_layers_embedding=[4])
    assert FO.build(m, True, [0, 0]) is None            # tanh MLP latent map: torch path only
    m2 = GP_Plus(torch.tensor(X), torch.tensor(y), qual_dict={2: 3}, dtype=torch.float64)
    assert FO.build(m2, True, [0.1, 0]) is None         # weight regularisation: torch path only
    assert FO.build(m2, False, [0, 0]) is not None


def test_float32_models_take_the_closed_form_path():
    """Default-dtype (float32) models: the closed forms run in float64, the torch path in float32; they agree to
    float32 accuracy and the fit does not fall back to the 1.4 ms torch objective."""
    from gpplus_b200.models import GP_Plus
    from gpplus_b200.optim import _fast_objective as FO
    from gpplus_b200.optim.mll_scipy import MLLObjective
    rng = np.random.default_rng(0)
    X = np.hstack([rng.standard_normal((50, 3)), rng.integers(0, 3, (50, 1)).astype(float)])
    y = np.sin(X[:, 0]) + 0.2 * X[:, 3]
    m = GP_Plus(torch.tensor(X), torch.tensor(y), qual_dict={3: 3})
    assert next(m.parameters()).dtype == torch.float32
    obj = MLLObjective(m, True, [0, 0])
    fast = FO.build(m, True, [0, 0])
    assert fast is not None and FO.self_check(obj, fast, trials=4)
    assert not FO.self_check(obj, fast, trials=4, tol=1e-9)   # float32 rounding of the torch path is visible



def test_get_params_names():
    from gpplus_b200.models import GP_Plus
    rng = np.random.default_rng(0)
    X = np.hstack([rng.standard_normal((30, 3)), rng.integers(0, 3, (30, 1)).astype(float)])
    m = GP_Plus(torch.tensor(X), torch.tensor(X[:, 0]), qual_dict={3: 3}, dtype=torch.float64)
    assert m.get_params("Omega").shape == (1, 3)
    assert m.get_params("Sigma") is m.covar_module.raw_outputscale
    assert m.get_params("Noise") is m.likelihood.noise_covar.raw_noise
    assert m.get_params("Mean") is m.mean_module.constant
    assert "latent[3]" in m.get_params()
