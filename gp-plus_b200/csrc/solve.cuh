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

// part[tb][j] = sum_{i in row block tb, i >= j} M[i][j] v[i]; grid = lower (tb, cb) block pairs, 128 threads
__global__ void __launch_bounds__(128) trmv_lower_t_part_kernel(const double* M, long long ld, const double* v,
                                                                int np, double* part) {
    __shared__ double sv[128];
    int bid = blockIdx.x;
    int tb = (int)((sqrt(8.0 * (double)bid + 1.0) - 1.0) * 0.5);
    while ((long long)(tb + 1) * (tb + 2) / 2 <= bid) tb++;
    while ((long long)tb * (tb + 1) / 2 > bid) tb--;
    const int cb = bid - (int)((long long)tb * (tb + 1) / 2);
    const int tid = threadIdx.x;
    sv[tid] = v[tb * 128 + tid];
    __syncthreads();
    const int j = cb * 128 + tid;
    const double* Mp = M + (long long)tb * 128 * ld + j;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    // on the diagonal block rows i < j hold zeros in M (strict upper part of L^-1), so no masking is needed
#pragma unroll 4
    for (int i = 0; i < 128; i += 4) {
        s0 = fma(Mp[(long long)(i + 0) * ld], sv[i + 0], s0);
        s1 = fma(Mp[(long long)(i + 1) * ld], sv[i + 1], s1);
        s2 = fma(Mp[(long long)(i + 2) * ld], sv[i + 2], s2);
        s3 = fma(Mp[(long long)(i + 3) * ld], sv[i + 3], s3);
    }
    part[(long long)tb * np + j] = (s0 + s1) + (s2 + s3);
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
