"""Parity at the HEADLINE workload (BASELINE configs[3]: synthetic exact GP, D=10, Matern-5/2, FP64):

* engine vs the CPU oracle (oracle/gp_oracle.mll, autograd through torch's Cholesky) at N=4096 and N=8192 on the
  bench workload, at theta_init and at two seeded prior draws (the starts a 64-restart fit would use);
* engine at N=16384 vs the committed oracle output tests/golden/c4_n16384_matern52.json (generated once by
  ``python tests/golden/make_golden.py --c4 16384``; bench.py asserts on the same file).

From N = 3072 up the engine's O(N^3) stages run as exact integer GEMMs on the INT8 tcgen05 tensor cores
(csrc/oz_gemm.cuh: 7 digit planes per operand); the tests above therefore exercise that path.  Two more cases pin it:
the same points with the exact-DMMA arithmetic (gpp_set_fp64_mode(0)), and an ill-conditioned covariance
(condition ~1e8: noise at its 1e-8 floor, long lengthscales) on both arithmetics.

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


def test_int8_sliced_and_dmma_arithmetic_agree_with_the_oracle_at_n4096():
    from gpplus_b200 import _engine as E
    n = 4096
    thetas = W.c4_theta_points(W.c4_model(256))
    prob = W.c4_oracle_problem(n)
    hyp = W.c4_natural(thetas[1])
    ref = O.mll(prob, hyp, want_grad=True)
    outs = {}
    for mode in (E.FP64_INT8, E.FP64_DMMA):
        prev = E.set_fp64_mode(mode)
        try:
            eng = _engine(n)
            try:
                assert eng.fp64_mode() == mode
                outs[mode] = eng.mll_grad(hyp, want_grad=True)
            finally:
                eng.close()
        finally:
            E.set_fp64_mode(prev)
        _check(outs[mode], ref, "n=4096 mode %d" % mode)
    a, b = outs[E.FP64_INT8], outs[E.FP64_DMMA]
    assert abs(a["nll"] - b["nll"]) <= 1e-11 * abs(b["nll"])


def test_ill_conditioned_covariance_on_both_arithmetics():
    """Condition number ~1e8 (the regime SURVEY 8(d) / the round-1 review name for the jitter ladder): noise at its
    lower bound 1e-8, lengthscales 20x longer than theta_init.  Both arithmetics against the oracle."""
    from gpplus_b200 import _engine as E
    n = 4096
    thetas = W.c4_theta_points(W.c4_model(256))
    prob = W.c4_oracle_problem(n)
    hyp = W.c4_natural(thetas[0])
    hyp["noise"] = np.array([1e-8])
    hyp["w"] = hyp["w"] * 0.002
    ref = O.mll(prob, hyp, want_grad=True, return_mats=True)
    ev = np.linalg.eigvalsh(np.asarray(ref["K"]) + 1e-8 * np.eye(n))
    assert ev[-1] / ev[0] > 1e7, ev[-1] / ev[0]
    gr = np.concatenate([np.asarray(ref["d_w"]), [ref["d_sigma_f2"]], np.asarray(ref["d_noise"]), np.asarray(ref["d_beta"])])
    for mode in (E.FP64_INT8, E.FP64_DMMA):
        prev = E.set_fp64_mode(mode)
        try:
            eng = _engine(n)
            try:
                out = eng.mll_grad(hyp, want_grad=True)
            finally:
                eng.close()
        finally:
            E.set_fp64_mode(prev)
        assert out["jitter"] == ref["jitter"], (mode, out["jitter"], ref["jitter"])
        assert abs(out["nll"] - ref["nll"]) <= 1e-9 * abs(ref["nll"]), (mode, out["nll"], ref["nll"])
        g = np.concatenate([out["d_w"], [out["d_sigma_f2"]], out["d_noise"], out["d_beta"]])
        # the noise gradient is 0.5 tr(K^-1 - alpha alpha^T) ~ 1e10 here and dominates the max-norm
        assert np.max(np.abs(g - gr)) <= 1e-8 * np.max(np.abs(gr)), (mode, np.max(np.abs(g - gr)), np.max(np.abs(gr)))


def test_block_dataflow_factorisation_matches_the_launch_chain():
    """potrf_lazy factors the diagonal block of every panel in one dataflow launch (block_potrf_kernel: left-looking
    tiles handed over through flags); GPP_BLOCK_CTAS=0 selects the launch-per-step chain it replaces.  Same arithmetic
    in a different summation order: the two must agree far inside the parity tolerance."""
    import os
    n = 5120   # T = 40; GPP_OZ_LAZY_MIN lowers the size threshold (96 tiles) of the lazy-panel path for this test
    thetas = W.c4_theta_points(W.c4_model(256))
    hyp = W.c4_natural(thetas[1])
    outs = []
    saved = {k: os.environ.get(k) for k in ("GPP_BLOCK_CTAS", "GPP_OZ_LAZY_MIN")}
    os.environ["GPP_OZ_LAZY_MIN"] = "36"
    try:
        for ctas in ("24", "0"):
            os.environ["GPP_BLOCK_CTAS"] = ctas
            eng = _engine(n)
            try:
                outs.append(eng.mll_grad(hyp, want_grad=True))
            finally:
                eng.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    a, b = outs
    assert abs(a["nll"] - b["nll"]) <= 1e-11 * abs(b["nll"]), (a["nll"], b["nll"])
    ga = np.concatenate([a["d_w"], [a["d_sigma_f2"]], a["d_noise"], a["d_beta"]])
    gb = np.concatenate([b["d_w"], [b["d_sigma_f2"]], b["d_noise"], b["d_beta"]])
    assert np.max(np.abs(ga - gb)) <= 1e-10 * np.max(np.abs(gb))
    ref = O.mll(W.c4_oracle_problem(n), hyp, want_grad=True)
    _check(a, ref, "n=5120 block dataflow")
