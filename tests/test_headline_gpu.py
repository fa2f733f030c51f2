"""Parity at the HEADLINE workload (BASELINE configs[3]: synthetic exact GP, D=10, Matern-5/2, FP64):

* engine vs the CPU oracle (oracle/gp_oracle.mll, autograd through torch's Cholesky) at N=4096 and N=8192 on the
  bench workload, at theta_init and at two seeded prior draws (the starts a 64-restart fit would use);
* engine at N=16384 vs the committed oracle output tests/golden/c4_n16384_matern52.json (generated once by
  ``python tests/golden/make_golden.py --c4 16384``; bench.py asserts on the same file).

Tolerances: NLL 1e-9 relative, every gradient entry 1e-8 of the gradient's max-norm.
"""
import json
import os

import numpy as np
import pytest

import bench_workloads as W
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_n16384_matern52.json")


def _engine(n):
    from gpplus_b200 import _engine as E
    X, y = W.c4_workload(n)
    ys = (y - y.min()) / (y.max() - y.min())
    return E.Engine(xq=X, y=ys, kernel=E.KERNEL_MATERN52, n_noise=1, n_mean=1, device=0)


def _check(out, ref, where):
    assert abs(out["nll"] - ref["nll"]) <= 1e-9 * abs(ref["nll"]), (where, out["nll"], ref["nll"])
    assert out["jitter"] == ref["jitter"], (where, out["jitter"], ref["jitter"])
    g = np.concatenate([out["d_w"], [out["d_sigma_f2"]], out["d_noise"], out["d_beta"]])
    gr = np.concatenate([np.asarray(ref["d_w"]), [ref["d_sigma_f2"]], np.asarray(ref["d_noise"]),
                         np.asarray(ref["d_beta"])])
    assert np.max(np.abs(g - gr)) <= 1e-8 * np.max(np.abs(gr)), (where, np.max(np.abs(g - gr)), np.max(np.abs(gr)))


@pytest.mark.parametrize("n", [4096, 8192])
def test_engine_matches_oracle_on_the_bench_workload(n):
    thetas = W.c4_theta_points(W.c4_model(256))
    prob = W.c4_oracle_problem(n)
    eng = _engine(n)
    try:
        for k in (0, 1, 2):
            hyp = W.c4_natural(thetas[k])
            _check(eng.mll_grad(hyp, want_grad=True), O.mll(prob, hyp, want_grad=True), "n=%d point %d" % (n, k))
    finally:
        eng.close()


def test_engine_matches_the_committed_oracle_output_at_n16384():
    gold = json.load(open(GOLD))
    assert gold["n"] == 16384
    thetas = W.c4_theta_points(W.c4_model(256))
    eng = _engine(16384)
    try:
        for pt in gold["points"]:
            np.testing.assert_array_equal(np.asarray(pt["theta"]), thetas[pt["index"]])  # same seeded points
            _check(eng.mll_grad(W.c4_natural(thetas[pt["index"]]), want_grad=True), pt, "n=16384 point %d" % pt["index"])
    finally:
        eng.close()
