"""Generate the golden fixtures in this directory from the CPU oracle (seeded, float64).

    python tests/golden/make_golden.py            # the small natural-parameter cases
    python tests/golden/make_golden.py --c4 16384  # headline workload (bench.py asserts on it): ~3 min and ~25 GB
                                                   # of host memory per evaluation point on 8 cores

The reference itself cannot be imported in this image (gpytorch / botorch are not installed), so these
vectors pin the ORACLE (regression) and give the GPU tests fixed targets; they are not outputs of the
reference.  Each .npz stores the problem, the hyper-parameters and the oracle's results.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import gp_oracle as O  # noqa: E402
from problems import make_candidates, make_hyper, make_problem  # noqa: E402

CASES = {
    "c1_borehole_like_n200_expsq": dict(n=200, dq=8, kernel=0, seed=11),
    "c2_mixed_n500_expsq_latent25": dict(n=500, dq=6, kernel=0, dz=2, n_combo=25, seed=12),
    "c3_mf_n350_expsq_4src": dict(n=350, dq=10, kernel=0, dz=2, n_combo=4, n_noise=4, n_mean=4, seed=13,
                                  zero_mean_group=True),
    "c4_small_n640_matern52": dict(n=640, dq=10, kernel=2, seed=14),
    "edge_n1_matern32": dict(n=1, dq=2, kernel=1, seed=15),
    "edge_n129_matern32_latent": dict(n=129, dq=3, kernel=1, dz=2, n_combo=3, seed=16),
}


def main():
    for name, kw in CASES.items():
        p = make_problem(**kw)
        h = make_hyper(p, seed=kw["seed"] + 100)
        c = make_candidates(p, 64, seed=kw["seed"] + 200)
        res = O.mll(p, h, want_grad=True, mode="direct")
        mu, var = O.predict(p, h, c, include_noise=True, mode="direct")
        blob = {}
        for k, v in p.items():
            if v is not None:
                blob["p_" + k] = np.asarray(v)
        for k, v in h.items():
            if v is not None:
                blob["h_" + k] = np.asarray(v)
        for k, v in c.items():
            if v is not None:
                blob["c_" + k] = np.asarray(v)
        for k, v in res.items():
            blob["r_" + k] = np.asarray(v)
        blob["r_pred_mean"] = mu
        blob["r_pred_var"] = var
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, "nll", res["nll"])


def main_c4(n, points=(0, 1, 2)):
    """NLL + 13 gradient entries of the CPU oracle on the C4 workload of bench.py (bench_workloads.py) at
    theta_init (point 0) and the first prior draws: the numbers bench.py and the -m gpu tests assert against."""
    import json
    import time
    import bench_workloads as W
    model = W.c4_model(256)  # the theta points depend on the priors only
    thetas = W.c4_theta_points(model)
    prob = W.c4_oracle_problem(n)
    out = {"n": n, "workload": "bench_workloads.c4_workload", "oracle": "oracle.gp_oracle.mll(mode='expansion')",
           "points": []}
    for k in points:
        t0 = time.time()
        r = O.mll(prob, W.c4_natural(thetas[k]), want_grad=True)
        out["points"].append({"index": int(k), "theta": [float(v) for v in thetas[k]], "nll": r["nll"],
                              "quad": r["quad"], "logdet": r["logdet"], "jitter": r["jitter"],
                              "d_w": [float(v) for v in r["d_w"]], "d_sigma_f2": r["d_sigma_f2"],
                              "d_noise": [float(v) for v in r["d_noise"]], "d_beta": [float(v) for v in r["d_beta"]],
                              "oracle_seconds": time.time() - t0})
        print("c4 n=%d point %d nll %.9f (%.1f s)" % (n, k, r["nll"], time.time() - t0), flush=True)
        with open(os.path.join(HERE, "c4_n%d_matern52.json" % n), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--c4":
        sys.path.insert(0, os.path.join(ROOT, "gp-plus_b200"))
        main_c4(int(sys.argv[2]))
    else:
        main()
