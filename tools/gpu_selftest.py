"""Stage-by-stage GPU diagnostics against the CPU oracle (development tool; prints, never asserts).

Run on the GPU box:  python tools/gpu_selftest.py [--big]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gp-plus_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gpplus_b200 import _engine as E  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from problems import engine_kwargs, make_candidates, make_hyper, make_problem  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def run_case(name, p, h, fetch=True, mode="expansion"):
    print("=== %s: n=%d dq=%d dz=%d kernel=%d n_noise=%d n_mean=%d" % (
        name, p["n"], p["dq"], p["dz"], p["kernel"], p["n_noise"], p["n_mean"]), flush=True)
    t0 = time.time()
    ref = O.mll(p, h, want_grad=True, mode=mode, return_mats=True)
    t_ref = time.time() - t0
    eng = E.Engine(**engine_kwargs(p))
    K = eng.covariance(h)
    print("  K      rel err %.3e" % relerr(K, ref["K"]), flush=True)
    t0 = time.time()
    out = eng.mll_grad(h, want_grad=True)
    t_gpu = time.time() - t0
    print("  nll    gpu %.12e ref %.12e rel %.3e (jitter %g/%g)" % (
        out["nll"], ref["nll"], abs(out["nll"] - ref["nll"]) / abs(ref["nll"]), out["jitter"], ref["jitter"]))
    print("  quad   rel %.3e  logdet rel %.3e" % (abs(out["quad"] - ref["quad"]) / abs(ref["quad"]),
                                                  abs(out["logdet"] - ref["logdet"]) / max(1e-300, abs(ref["logdet"]))))
    if fetch:
        L = eng.fetch("L")
        print("  L      rel err %.3e" % relerr(L, ref["L"]))
        Li = eng.fetch("Linv")
        Li_ref = np.linalg.inv(ref["L"])
        print("  Linv   rel err %.3e" % relerr(Li, Li_ref))
        al = eng.fetch("alpha")
        print("  alpha  rel err %.3e" % relerr(al, ref["alpha"]))
        Ki = eng.fetch("Kinv")
        print("  Kinv   rel err %.3e" % relerr(Ki, ref["Kinv"]))
    for k in ("d_sigma_f2", "d_w", "d_noise", "d_z", "d_beta"):
        if k in ref:
            print("  %-10s rel err %.3e  (ref max %.3e)" % (k, relerr(out[k], ref[k]), np.max(np.abs(ref[k]))))
    print("  timings %s" % eng.timings())
    print("  wall: oracle %.3fs gpu first call %.3fs" % (t_ref, t_gpu), flush=True)
    # prediction
    c = make_candidates(p, 300)
    mu_ref, var_ref = O.predict(p, h, c, include_noise=True, mode=mode)
    eng.factorize(h)
    mu, var = eng.predict(c["xq"], c["level_idx"], c["noise_idx"], c["mean_idx"], include_noise=True)
    print("  predict mean rel %.3e  var rel %.3e" % (relerr(mu, mu_ref), relerr(var, var_ref)))
    best, idx, scores = eng.acq_argmax(c["xq"], np.zeros(c["m"], dtype=np.int32), [1.0], [E.ACQ_EI], [0.5],
                                       level_idx=c["level_idx"], mean_idx=c["mean_idx"], return_scores=True)
    mu2, var2 = O.predict(p, h, c, include_noise=False, mode=mode)
    sc_ref = O.acquisition(mu2, np.sqrt(var2), 2, 0.5, np.ones(c["m"]))
    print("  acq EI scores rel %.3e argmax gpu %d ref %d" % (relerr(scores, sc_ref), idx, int(np.argmax(sc_ref))),
          flush=True)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    print("devices:", E.device_count(), flush=True)
    for (m, n, k) in ((1024, 1024, 1024), (4096, 4096, 4096), (8192, 8192, 8192)):
        ms = E.probe_dgemm(m, n, k, iters=5)
        print("dgemm probe %dx%dx%d: %.3f ms  %.2f TFLOP/s" % (m, n, k, ms, 2.0 * m * n * k / ms / 1e9), flush=True)
    cases = [
        ("tiny-expsq", make_problem(5, 3, E.KERNEL_EXPSQ, seed=3)),
        ("one-tile-m52", make_problem(100, 4, E.KERNEL_MATERN52, seed=4)),
        ("two-tile-m32", make_problem(200, 8, E.KERNEL_MATERN32, seed=5)),
        ("latent-expsq", make_problem(300, 6, E.KERNEL_EXPSQ, dz=2, n_combo=25, seed=6)),
        ("mf-m52", make_problem(700, 10, E.KERNEL_MATERN52, dz=2, n_combo=4, n_noise=4, n_mean=4, seed=7,
                                zero_mean_group=True)),
        ("t9-expsq", make_problem(1100, 10, E.KERNEL_EXPSQ, seed=8)),
    ]
    for name, p in cases:
        try:
            run_case(name, p, make_hyper(p))
        except Exception as e:  # keep going: this is a diagnostic
            print("  !! %s failed: %r" % (name, e), flush=True)
    if args.big:
        p = make_problem(4096, 10, E.KERNEL_MATERN52, seed=9)
        try:
            run_case("n4096-m52", p, make_hyper(p), fetch=True)
        except Exception as e:
            print("  !! big failed: %r" % (e,), flush=True)
        # throughput at the headline size (no oracle: 16384 is minutes on the CPU)
        p = make_problem(16384, 10, E.KERNEL_MATERN52, seed=10)
        h = make_hyper(p)
        eng = E.Engine(**engine_kwargs(p))
        for it in range(3):
            t0 = time.time()
            out = eng.mll_grad(h, want_grad=True)
            print("n16384 eval %d: wall %.1f ms nll %.9e timings %s" % (it, 1e3 * (time.time() - t0), out["nll"],
                                                                      eng.timings()), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
