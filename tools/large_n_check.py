"""Consistency of the engine beyond the headline size: look-ahead vs blocked schedule, and a central finite
difference of the outputscale gradient, at N = 32768 and 49152 (index arithmetic past 2^31 elements)."""
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'gp-plus_b200')
import numpy as np
import bench
from gpplus_b200 import _engine as E

for n in (32768, 49152):
    X, y = bench.W.c4_workload(n)
    ys = (y - y.min()) / (y.max() - y.min())
    h = bench.W.c4_natural(0.05 * np.random.RandomState(3).randn(13))
    res = {}
    for mode in ("lookahead", "blocked"):
        os.environ["GPP_CHOL"] = mode
        eng = E.Engine(xq=X, y=ys, kernel=E.KERNEL_MATERN52, n_noise=1, n_mean=1, device=0)
        t0 = time.time()
        out = eng.mll_grad(h, want_grad=True)
        dt = time.time() - t0
        res[mode] = out
        if mode == "lookahead":
            eps = 1e-4
            hp, hm = dict(h), dict(h)
            hp["sigma_f2"] = h["sigma_f2"] + eps
            hm["sigma_f2"] = h["sigma_f2"] - eps
            fd = (eng.mll_grad(hp, want_grad=False)["nll"] - eng.mll_grad(hm, want_grad=False)["nll"]) / (2 * eps)
            print("n=%d %s: %.2f s/eval  nll %.10e  d_sf2 %.8e  fd %.8e  rel %.2e  timings %s" % (
                n, mode, dt, out["nll"], out["d_sigma_f2"], fd, abs(fd - out["d_sigma_f2"]) / abs(fd),
                {k: round(v, 1) for k, v in eng.timings().items()}), flush=True)
        eng.close()
    a, b = res["lookahead"], res["blocked"]
    print("n=%d lookahead vs blocked: nll rel %.2e  d_w rel %.2e" % (
        n, abs(a["nll"] - b["nll"]) / abs(a["nll"]), np.max(np.abs(a["d_w"] - b["d_w"])) / np.max(np.abs(a["d_w"]))), flush=True)
