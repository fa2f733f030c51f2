// Triangular sweeps with the explicit inverse factor M = L^-1 (lower, row-major, ld = np):
//   v     = M r          (whitened residual;  quad = |v|^2,  MultivariateNormal.log_prob's
//                          Mahalanobis term, optim/mll_scipy.py:38-39 / SURVEY A.5)
//   alpha = M^T v        (= K_y^-1 (y - m), DefaultPredictionStrategy.mean_cache)
// Both are HBM-bound sweeps over the lower triangle (4*N^2 bytes each) with a fixed summation
// order, so results are bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>

namespace gpp {

// one warp per row, 8 rows per CTA
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* M, long long ld, const double* r, int np,
                                                         double* v) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= np) return;
    const double* Mr = M + (long long)row * ld;
    double s0 = 0.0, s1 = 0.0;
    const int npair = (row + 2) >> 1;  // columns [0, 2*npair) cover 0..row (entry row+1 of M is 0 if present)
    for (int p = lane; p < npair; p += 32) {
        double2 m = *reinterpret_cast<const double2*>(Mr + 2 * p);
        double2 x = *reinterpret_cast<const double2*>(r + 2 * p);
        s0 = fma(m.x, x.x, s0);
        if (2 * p + 1 <= row) s1 = fma(m.y, x.y, s1);
    }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) v[row] = s;
}

// part[tb][j] = sum_{i in row block tb, i >= j} M[i][j] v[i]; grid = lower (tb, cb) block pairs.  512 threads: four
// row groups of 32 rows per column with eight independent accumulators each, so that a tile is four batches of
// eight loads in flight per thread instead of a 32-deep dependent chain (37.5 -> 6.9 us for the 10 tiles of n = 500,
// where this sweep is pure latency); the groups are combined in a fixed order.
__global__ void __launch_bounds__(512) trmv_lower_t_part_kernel(const double* M, long long ld, const double* v,
                                                                int np, double* part) {
    __shared__ double sv[128];
    __shared__ double sp[3][128];
    int bid = blockIdx.x;
    int tb = (int)((sqrt(8.0 * (double)bid + 1.0) - 1.0) * 0.5);
    while ((long long)(tb + 1) * (tb + 2) / 2 <= bid) tb++;
    while ((long long)tb * (tb + 1) / 2 > bid) tb--;
    const int cb = bid - (int)((long long)tb * (tb + 1) / 2);
    const int tid = threadIdx.x;
    const int c = tid & 127, rg = tid >> 7;
    if (tid < 128) sv[tid] = v[tb * 128 + tid];
    __syncthreads();
    const int j = cb * 128 + c;
    const double* Mp = M + ((long long)tb * 128 + rg * 32) * ld + j;
    const double* svp = sv + rg * 32;
    double s[8];
#pragma unroll
    for (int u = 0; u < 8; u++) s[u] = 0.0;
    // on the diagonal block rows i < j hold zeros in M (strict upper part of L^-1), so no masking is needed
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        double m[8];
#pragma unroll
        for (int u = 0; u < 8; u++) m[u] = Mp[(long long)(i + u) * ld];
#pragma unroll
        for (int u = 0; u < 8; u++) s[u] = fma(m[u], svp[i + u], s[u]);
    }
    const double tot = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    if (rg > 0) sp[rg - 1][c] = tot;
    __syncthreads();
    if (rg == 0) part[(long long)tb * np + j] = (tot + sp[0][c]) + (sp[1][c] + sp[2][c]);
}

// alpha[j] = sum_{tb >= block(j)} part[tb][j]
__global__ void trmv_lower_t_reduce_kernel(const double* part, int np, int T, double* alpha) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np) return;
    double s = 0.0;
    for (int tb = j >> 7; tb < T; tb++) s += part[(long long)tb * np + j];
    alpha[j] = s;
}

// upper triangle (strict) of every diagonal 128-block and everything above it must be zero in M for the sweeps
// and GEMMs that skip by tile; this clears the strict upper tiles of a buffer once.
__global__ void zero_upper_tiles_kernel(double* A, long long ld, int T) {
    int ti = blockIdx.y, tj = blockIdx.x;
    if (tj <= ti) return;
    double* p = A + (long long)ti * 128 * ld + (long long)tj * 128;
    for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) p[(long long)(idx >> 7) * ld + (idx & 127)] = 0.0;
}

}  // namespace gpp
