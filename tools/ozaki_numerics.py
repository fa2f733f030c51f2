"""CPU feasibility study of the INT8-sliced ("Ozaki") FP64 GEMM for the three O(N^3) stages (numpy, no GPU).

Every large GEMM of the evaluation (trailing updates of the factorisation, the levels of the triangular inverse,
K^-1 = L^-T L^-1) is replaced by  sum_{i+j<=LEVELS-1} A_i B_j^T 2^-8(i+j+2)  on row-scaled signed base-256 digit
planes (exact integer products; emulated here with float64 BLAS on the digit planes, which is exact below 2^53).
Prints the objective and gradient differences against plain float64 LAPACK on the C4 workload.

    python tools/ozaki_numerics.py [n] [slices]
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gp-plus_b200")]

S = int(sys.argv[2]) if len(sys.argv) > 2 else 7      # digit planes per operand
LEVELS = int(sys.argv[3]) if len(sys.argv) > 3 else S  # keep pairs with i + j < LEVELS


from oracle import oz_oracle as OZ  # noqa: E402

LEVELS_KINV = int(os.environ.get("LEVELS_KINV", LEVELS))    # levels kept in K^-1 = M^T M only
LEVELS_TRTRI = int(os.environ.get("LEVELS_TRTRI", LEVELS))  # levels kept in the triangular inverse only


def ozaki_abt(A, B, levels=None):
    """A [m,K] @ B[n,K]^T through digit planes (oracle/oz_oracle.py)."""
    return OZ.abt(A, B, LEVELS if levels is None else levels, S)


def chol_blocked(K, nb, mm):
    """right-looking blocked Cholesky; the trailing update runs through `mm` (A, B) -> A B^T"""
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


def trtri_doubling(L, nb, mm):
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
            X = mm(L[m0:m1, g0:m0], M[g0:m0, g0:m0].T)          # L21 M11
            M[m0:m1, g0:m0] = -mm(M[m0:m1, m0:m1], X.T)         # -M22 X
        h *= 2
    return M


def evaluate(K, r, mm, nb_chol, nb_inv):
    L = chol_blocked(K, nb_chol, mm)
    M = trtri_doubling(L, nb_inv, mm if mm is not ozaki_abt else (lambda A, B: ozaki_abt(A, B, LEVELS_TRTRI)))
    Kinv = mm(M.T, M.T) if mm is not ozaki_abt else ozaki_abt(M.T, M.T, LEVELS_KINV)   # M^T M
    u = M @ r
    alpha = M.T @ u
    nll = 0.5 * (u @ u) + np.sum(np.log(np.diag(L))) + 0.5 * len(r) * np.log(2 * np.pi)
    W = np.outer(alpha, alpha) - Kinv
    return nll, W, L


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    import bench_workloads as W
    from oracle import gp_oracle as O
    import torch
    prob = W.c4_oracle_problem(n)
    model = W.c4_model(n)
    thetas = W.c4_theta_points(model)
    noises = [np.exp(np.float32(t[0])) for t in thetas[1:]]
    pick = [0, 1 + int(np.argmin(noises)), 1 + int(np.argmax([t[2:12].max() for t in thetas[1:]]))]
    fp64 = lambda A, B: A @ B.T
    for idx in pick + ["harsh"]:
        if idx == "harsh":
            h = W.c4_natural(thetas[0])
            h["noise"] = np.array([1e-8])
            h["w"] = h["w"] * float(os.environ.get("HARSH_W", "0.02"))
        else:
            h = W.c4_natural(thetas[idx])
        X = torch.as_tensor(prob["xq"])
        K0t = O.covariance(X, None, X, None, torch.as_tensor(h["w"]), None, float(h["sigma_f2"]), 2)
        K = np.asarray(K0t, dtype=np.float64) + h["noise"][0] * np.eye(n)
        r = prob["y"] - h["beta"][0]
        ev = np.linalg.eigvalsh(K)
        n0, W0, L0 = evaluate(K, r, fp64, 128, 128)
        n1, W1, L1 = evaluate(K, r, ozaki_abt, 128, 128)
        # a gradient-like functional: tr(W dK/dsigma) = sum(W * K0) / sigma_f2 and the noise gradient tr(W)
        K0 = K - h["noise"][0] * np.eye(n)
        g0 = np.array([0.5 * np.sum(W0 * K0), 0.5 * np.trace(W0)])
        g1 = np.array([0.5 * np.sum(W1 * K0), 0.5 * np.trace(W1)])
        print("point %-6s cond %.2e  nll %.12e  |d nll|/|nll| %.2e   |dL|max/|L|max %.2e   |dW|max/|W|max %.2e   "
              "grad rel %.2e %.2e" % (idx, ev[-1] / ev[0], n0, abs(n1 - n0) / abs(n0),
                                      np.max(np.abs(L1 - L0)) / np.max(np.abs(L0)),
                                      np.max(np.abs(W1 - W0)) / np.max(np.abs(W0)),
                                      abs(g1[0] - g0[0]) / np.max(np.abs(g0)), abs(g1[1] - g0[1]) / np.max(np.abs(g0))),
              flush=True)


if __name__ == "__main__":
    main()
