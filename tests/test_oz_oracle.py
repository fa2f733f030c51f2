"""CPU checks of the INT8-sliced FP64 product (oracle/oz_oracle.py restates csrc/oz_split.cuh + csrc/oz_gemm.cuh):

* digit planes: signed base-256 digits in [-128, 127], leading digit within +-64, exact reconstruction to 2^-56 of the
  row scale, scale = the power of two that maps the row maximum into [1/8, 1/4);
* the 28-pair product against a long-double product: error <= K * 2^-50 * max|row| * max|col| (operand truncation 2^-56
  twice, dropped pairs 2^-54), and what 6 or 5 significance levels cost (the K^-1 product of the engine keeps 6);
* a whole evaluation (blocked Cholesky -> recursive-doubling inverse -> K^-1 -> objective and gradient traces) with every
  large product replaced, on a covariance of condition ~1e10 (harsher than the 1e8 the round-1 review asked for):
  objective 1e-9 / gradient 1e-8 against LAPACK (observed 1.4e-10 / 2.0e-9), and the same with 6 planes FAILING it
  (4e-6: why 7).
The GPU kernels are checked against the same definitions in tests/test_oz_gemm_gpu.py (tools/oz_lab check).
"""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import oz_oracle as OZ


def _rand(rows, k, seed, spread=3.0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((rows, k)) * 10.0 ** (spread * rng.uniform(-1, 1, (rows, 1))) * \
        np.where(rng.uniform(size=(rows, k)) < 0.15, 1e-7, 1.0)


def test_digit_planes_are_exact_balanced_base_256():
    A = _rand(37, 200, 0)
    A[5] = 0.0                                   # a zero row keeps scale 1 and zero digits
    d, scale = OZ.split_rows(A)
    assert d.shape == (7, 37, 200)
    assert d.min() >= -128 and d.max() <= 127 and np.all(d == np.rint(d))
    assert np.abs(d[0]).max() <= 64
    amax = np.abs(A).max(axis=1)
    nz = amax > 0
    assert np.all((amax[nz] / scale[nz] >= 0.125) & (amax[nz] / scale[nz] < 0.25))
    assert scale[5] == 1.0 and not d[:, 5].any()
    rec = sum(d[p].astype(np.longdouble) * np.longdouble(256.0) ** -(p + 1) for p in range(7)) * scale[:, None]
    t = (A.astype(np.longdouble) - rec) / scale[:, None]
    assert t.min() >= 0 and t.max() < 2.0 ** -56


@pytest.mark.parametrize("k", [64, 1536])
def test_product_error_bound_and_the_cost_of_fewer_levels(k):
    A, B = _rand(48, k, 1), _rand(40, k, 2)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    unit = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :]
    err7 = np.max(np.abs(OZ.abt(A, B) - ref) / unit)
    err6 = np.max(np.abs(OZ.abt(A, B, levels=6) - ref) / unit)
    err5 = np.max(np.abs(OZ.abt(A, B, levels=5) - ref) / unit)
    assert err7 <= k * 2.0 ** -50
    # random digits: the sums grow like sqrt(K); each dropped level costs about 2^8
    assert err7 <= 40 * np.sqrt(k) * 2.0 ** -54
    assert err7 < err6 < err5 and err6 <= k * 2.0 ** -42 and err5 <= k * 2.0 ** -34
    # the same operands in plain float64 are no better than the 28-pair product by more than a small factor
    err64 = np.max(np.abs(A @ B.T - ref) / unit)
    assert err7 <= 64 * max(err64, 2.0 ** -60)


def _chol_blocked(K, nb, mm):
    A = K.copy()
    n = A.shape[0]
    for k0 in range(0, n, nb):
        k1 = min(n, k0 + nb)
        A[k0:k1, k0:k1] = np.linalg.cholesky(A[k0:k1, k0:k1])
        if k1 < n:
            A[k1:, k0:k1] = sla.solve_triangular(A[k0:k1, k0:k1], A[k1:, k0:k1].T, lower=True).T
            P = A[k1:, k0:k1]
            A[k1:, k1:] -= mm(P, P)
    return np.tril(A)


def _trtri_doubling(L, nb, mm):
    n = L.shape[0]
    M = np.zeros_like(L)
    for k0 in range(0, n, nb):
        k1 = min(n, k0 + nb)
        M[k0:k1, k0:k1] = sla.solve_triangular(L[k0:k1, k0:k1], np.eye(k1 - k0), lower=True)
    h = nb
    while h < n:
        for g0 in range(0, n, 2 * h):
            m0, m1 = g0 + h, min(n, g0 + 2 * h)
            if m0 >= n:
                break
            X = mm(L[m0:m1, g0:m0], M[g0:m0, g0:m0].T)
            M[m0:m1, g0:m0] = -mm(M[m0:m1, m0:m1], X.T)
        h *= 2
    return M


def _evaluate(K, K0, r, mm, mm_kinv):
    L = _chol_blocked(K, 64, mm)
    M = _trtri_doubling(L, 64, mm)
    Kinv = mm_kinv(M.T, M.T)
    u = M @ r
    alpha = M.T @ u
    nll = 0.5 * (u @ u) + np.sum(np.log(np.diag(L))) + 0.5 * len(r) * np.log(2 * np.pi)
    W = np.outer(alpha, alpha) - Kinv
    return nll, np.array([0.5 * np.sum(W * K0), 0.5 * np.trace(W)])      # d/d log sigma_f^2 and d/d noise traces


def test_whole_evaluation_at_condition_1e10_seven_planes_pass_six_fail():
    n = 384
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (n, 3))
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    rr = np.sqrt(5.0 * d2 / 9.0)
    K0 = (1.0 + rr + rr * rr / 3.0) * np.exp(-rr)            # Matern-5/2, long lengthscale
    K = K0 + 1e-8 * np.eye(n)
    ev = np.linalg.eigvalsh(K)
    assert ev[-1] / ev[0] > 1e8
    r = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(n)
    fp64 = lambda A, B: A @ B.T
    n0, g0 = _evaluate(K, K0, r, fp64, fp64)
    # the engine's configuration: 7 planes / 28 pairs everywhere, 6 levels in K^-1 (gradient only)
    n1, g1 = _evaluate(K, K0, r, OZ.abt, lambda A, B: OZ.abt(A, B, levels=6))
    assert abs(n1 - n0) <= 1e-9 * abs(n0)
    assert np.max(np.abs(g1 - g0)) <= 1e-8 * np.max(np.abs(g0))
    # six planes (21 pairs) everywhere miss the objective tolerance on this matrix
    six = lambda A, B: OZ.abt(A, B, levels=6, planes=6)
    n2, _ = _evaluate(K, K0, r, six, six)
    assert abs(n2 - n0) > 1e-9 * abs(n0) > abs(n1 - n0)
