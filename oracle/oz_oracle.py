"""CPU restatement of the INT8-sliced ("Ozaki") FP64 product used by the engine's O(N^3) stages -- TEST INFRASTRUCTURE.

Only ``tests/`` and ``tools/ozaki_numerics.py`` import this module; the product package never does.  It restates, in
numpy, what ``gp-plus_b200/csrc/oz_split.cuh`` (digit planes) and ``gp-plus_b200/csrc/oz_gemm.cuh`` (28 plane pairs,
one exact integer accumulator per significance level, FP64 recombination) compute, so that the accuracy claims of
DESIGN.md section 4.1 can be checked without a GPU.  The reference (Bostanabad-Research-Group/GP-Plus) has no
counterpart: it reaches torch.linalg.cholesky / cholesky_backward through MultivariateNormal.log_prob
(optim/mll_scipy.py:37-39,123); these products replace the FP64 GEMMs inside those factorisations.

The integer GEMMs are emulated with float64 BLAS on the digit planes, which is exact while every partial sum stays
below 2^53 (|digit| <= 128, K <= 2^20: 2^14 * 2^20 * 7 < 2^53)."""
from __future__ import annotations

import numpy as np

S = 7           # digit planes per operand
LEVELS = 7      # significance levels kept: plane pairs (i, j) with i + j < LEVELS (28 pairs)


def split_rows(A: np.ndarray, planes: int = S):
    """A [rows, K] -> (digits [planes, rows, K] as float64 holding integers in [-128, 127], scale [rows]) with
    A = scale * sum_p digits[p] 256^-(p+1) + t, 0 <= t < scale 2^-56 (planes = 7).  scale = 2^(e+2) for a row maximum in
    [2^(e-1), 2^e); v = floor(A / scale * 2^(8 planes)) is split into balanced base-256 digits by integer arithmetic,
    least significant first (oz_split.cuh::oz_digits)."""
    A = np.asarray(A, dtype=np.float64)
    amax = np.max(np.abs(A), axis=1)
    _, ex = np.frexp(amax)
    e = np.where(amax > 0, ex + 2, 0).astype(np.int64)
    scale = np.ldexp(1.0, e)
    v = np.floor(np.ldexp(A, (8 * planes - e)[:, None])).astype(np.int64)
    digits = np.empty((planes,) + A.shape)
    for p in range(planes - 1, -1, -1):
        lo = ((v & 0xFF) ^ 0x80) - 0x80          # sign-extended low byte
        digits[p] = lo
        v = (v - lo) >> 8
    assert not np.any(v), "leading digit overflow"
    return digits, scale


def abt(A: np.ndarray, B: np.ndarray, levels: int = LEVELS, planes: int = S) -> np.ndarray:
    """A [m, K] @ B [n, K]^T through digit planes: sum over the levels of exact integer plane-pair products."""
    pa, sa = split_rows(A, planes)
    pb, sb = split_rows(B, planes)
    acc = np.zeros((A.shape[0], B.shape[0]))
    for lvl in range(levels - 1, -1, -1):       # Horner over the levels, least significant first (oz_gemm.cuh epilogue)
        t = np.zeros_like(acc)
        for i in range(min(lvl, planes - 1) + 1):
            j = lvl - i
            if j < planes:
                t += pa[i] @ pb[j].T
        acc = acc / 256.0 + t
    return acc / 65536.0 * sa[:, None] * sb[None, :]
