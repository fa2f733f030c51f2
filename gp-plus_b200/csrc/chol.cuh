// Blocked FP64 Cholesky, triangular inverse and K^-1 on top of the DMMA tile GEMM.
//
// Replaces psd_safe_cholesky / torch.linalg.cholesky_ex and the autograd cholesky_backward
// the reference reaches through MultivariateNormal.log_prob (optim/mll_scipy.py:37-39,123;
// SURVEY Appendix A.5).  All matrices are row-major with leading dimension ld = Np, Np a
// multiple of 128 (padded rows/cols carry an identity block, so they change neither the
// factor, log|K| nor the solves).
//
//   1. potrf: right-looking, two-level blocking.  Inside a panel of PANEL_BLOCKS 128-columns:
//        leaf  (one CTA): L_kk = chol(A_kk) and L_kk^-1 (written into the diagonal block of M)
//        TRSM  (DMMA GEMM, K=128): A_ik <- A_ik * L_kk^-T  via the explicit 128x128 inverse
//        panel update (DMMA GEMM, K=128) of the remaining panel columns
//      then one trailing SYRK with K = 128*PANEL_BLOCKS (DMMA GEMM, lower tiles only).
//   2. trtri: M = L^-1 by recursive doubling.  With the diagonal 128-blocks already inverted,
//      level h combines aligned groups of h blocks:  M21 = -M22 * (L21 * M11); every group of
//      a level is independent, so a level is two batched GEMM launches (ceil(log2 T) levels).
//   3. lauum: K^-1 = M^T M, one launch over the lower tiles, mirrored into the upper triangle.
#pragma once
#include "dgemm_dmma.cuh"

namespace gpp {

constexpr int LEAF_THREADS = 256;
constexpr int LEAF_LDS = TILE + 1;
constexpr int LEAF_SMEM_BYTES = (TILE * LEAF_LDS + 2 * TILE) * 8;
constexpr int PANEL_BLOCKS = 4;

// info[0]: 0 ok, 1 = non-positive pivot, 2 = NaN pivot.  logdet_part[kb] = sum_j log L_jj of block kb.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_trinv_kernel(double* A, int ld, int kb, double* M, double* logdet_part, int* info) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                      // [128][129]: lower = L, strict upper = (L^-1)^T
    double* dsq = sm + TILE * LEAF_LDS;  // [128] L_jj
    double* xd = dsq + TILE;             // [128] 1/L_jj
    const int tid = threadIdx.x;
    double* Ab = A + (long long)kb * TILE * ld + (long long)kb * TILE;
    double* Mb = M + (long long)kb * TILE * ld + (long long)kb * TILE;

    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        S[i * LEAF_LDS + j] = (j <= i) ? Ab[(long long)i * ld + j] : 0.0;
    }
    __syncthreads();

    // right-looking factorisation without per-column scaling:
    //   S_ik -= S_ij * S_kj / S_jj  for j < k <= i ; column j is final after step j-1
    const int row = tid >> 1, half = tid & 1;
    int bad = 0;
    for (int j = 0; j < TILE - 1; j++) {
        double d = S[j * LEAF_LDS + j];
        if (!(d > 0.0)) bad |= (d != d) ? 2 : 1;
        if (row > j) {
            double lij = S[row * LEAF_LDS + j] / d;
            for (int k = j + 1 + half; k <= row; k += 2)
                S[row * LEAF_LDS + k] = fma(-lij, S[k * LEAF_LDS + j], S[row * LEAF_LDS + k]);
        }
        __syncthreads();
    }
    {
        double d = S[(TILE - 1) * LEAF_LDS + TILE - 1];
        if (!(d > 0.0)) bad |= (d != d) ? 2 : 1;
    }
    if (bad && tid == 0) atomicOr(info, bad);

    if (tid < TILE) {
        double d = S[tid * LEAF_LDS + tid];
        double r = sqrt(d);
        dsq[tid] = r;
        xd[tid] = 1.0 / r;
    }
    __syncthreads();
    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        if (j < i) S[i * LEAF_LDS + j] *= xd[j];
    }
    __syncthreads();
    if (tid < TILE) S[tid * LEAF_LDS + tid] = dsq[tid];
    if (tid < 32) {
        double s = 0.0;
        for (int j = tid; j < TILE; j += 32) s += log(dsq[j]);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) logdet_part[kb] = s;
    }
    __syncthreads();
    // L back to the lower triangle of A
    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        if (j <= i) Ab[(long long)i * ld + j] = S[i * LEAF_LDS + j];
    }

    // X = L^-1, column c by forward substitution, two lanes per column; X[i][c] lives at S[c][i]
    {
        const int c = tid >> 1, h = tid & 1;
        const unsigned pmask = 3u << ((tid & 31) & ~1);
        const double xc = xd[c];
        for (int i = c + 1; i < TILE; i++) {
            double p = (h == 0) ? S[i * LEAF_LDS + c] * xc : 0.0;
            for (int k = c + 1 + h; k < i; k += 2) p = fma(S[i * LEAF_LDS + k], S[c * LEAF_LDS + k], p);
            p += __shfl_xor_sync(pmask, p, 1);
            double xi = -p * xd[i];
            if (h == 0) S[c * LEAF_LDS + i] = xi;
            __syncwarp(pmask);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        int i = idx >> 7, c = idx & 127;
        double v = (c < i) ? S[c * LEAF_LDS + i] : ((c == i) ? xd[i] : 0.0);
        Mb[(long long)i * ld + c] = v;
    }
}

inline cudaError_t chol_set_attributes() {
    cudaError_t e = cudaFuncSetAttribute(leaf_potrf_trinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         LEAF_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return gemm_set_attributes();
}

#define GPP_TRY(x)                      \
    do {                                \
        cudaError_t _e = (x);           \
        if (_e != cudaSuccess) return _e; \
    } while (0)

// A (lower) -> L in place; diagonal blocks of M <- inverse of the diagonal blocks of L.
inline cudaError_t potrf_blocked(double* A, double* M, int ld, int T, double* logdet_part, int* info,
                                 cudaStream_t st) {
    for (int p0 = 0; p0 < T; p0 += PANEL_BLOCKS) {
        const int pw = (T - p0 < PANEL_BLOCKS) ? (T - p0) : PANEL_BLOCKS;
        const int pend = p0 + pw;
        for (int col = p0; col < pend; col++) {
            leaf_potrf_trinv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, st>>>(A, ld, col, M, logdet_part, info);
            GPP_TRY(cudaGetLastError());
            const int below = T - col - 1;
            if (below <= 0) continue;
            {   // TRSM: A[i,col] <- A[i,col] * Linv_col^T   (in place; a CTA only reads its own rows)
                GemmOp op = gemm_default();
                op.A = A + (long long)(col + 1) * TILE * ld + (long long)col * TILE;
                op.lda = ld;
                op.B = M + (long long)col * TILE * ld + (long long)col * TILE;
                op.ldb = ld;
                op.C = const_cast<double*>(op.A);
                op.ldc = ld;
                op.tiles_m = op.tiles_m_last = below;
                op.tiles_n = 1;
                op.klo_c = 0;
                op.khi_c = 1;
                GPP_TRY(launch_gemm(op, true, true, 1, st));
            }
            const int pc = pend - col - 1;  // panel columns still to update
            if (pc > 0) {
                // A[i,j] -= A[i,col] * A[j,col]^T  for col < j < pend, i >= j
                GemmOp op = gemm_default();
                op.A = A + (long long)(col + 1) * TILE * ld + (long long)col * TILE;
                op.lda = ld;
                op.B = op.A;
                op.ldb = ld;
                op.C = A + (long long)(col + 1) * TILE * ld + (long long)(col + 1) * TILE;
                op.ldc = ld;
                op.tiles_m = op.tiles_m_last = below;
                op.tiles_n = pc;
                op.lower_filter = 1;
                op.lower_off = 0;
                op.klo_c = 0;
                op.khi_c = 1;
                op.alpha = -1.0;
                op.beta = 1.0;
                GPP_TRY(launch_gemm(op, true, true, 1, st));
            }
        }
        const int rest = T - pend;
        if (rest > 0) {
            // trailing SYRK: A[i,j] -= sum_{k in panel} A[i,k] A[j,k]^T , i >= j >= pend
            GemmOp op = gemm_default();
            op.A = A + (long long)pend * TILE * ld + (long long)p0 * TILE;
            op.lda = ld;
            op.B = op.A;
            op.ldb = ld;
            op.C = A + (long long)pend * TILE * ld + (long long)pend * TILE;
            op.ldc = ld;
            op.map = MAP_TRI;
            op.tiles_m = op.tiles_m_last = rest;
            op.tiles_n = rest;
            op.klo_c = 0;
            op.khi_c = pw;
            op.alpha = -1.0;
            op.beta = 1.0;
            GPP_TRY(launch_gemm(op, true, true, 1, st));
        }
    }
    return cudaSuccess;
}

// M (diagonal 128-blocks already hold L_kk^-1) <- L^-1 (lower); X is an N x N scratch.
inline cudaError_t trtri_doubling(const double* L, double* M, double* X, int ld, int T, cudaStream_t st) {
    for (int hb = 1; hb < T; hb *= 2) {
        // groups of 2*hb blocks; first half [g0, g0+hb), second half [g0+hb, min(g0+2hb, T))
        const int ngroups = (T + 2 * hb - 1) / (2 * hb);
        int nb = 0, last_s2 = 0;  // groups with a non-empty second half
        for (int gidx = 0; gidx < ngroups; gidx++) {
            int s2 = T - (gidx * 2 * hb + hb);
            if (s2 <= 0) break;
            if (s2 > hb) s2 = hb;
            nb++;
            last_s2 = s2;
        }
        if (nb == 0) continue;
        const long long zs = (long long)2 * hb * TILE * ld + (long long)2 * hb * TILE;  // next group, diagonal step
        const long long off21 = (long long)hb * TILE * ld;                              // block (hb, 0) of the group
        const long long off22 = (long long)hb * TILE * ld + (long long)hb * TILE;
        {   // X21 = L21 * M11 ; M11 lower: k-blocks [tj, hb)
            GemmOp op = gemm_default();
            op.A = L + off21;
            op.lda = ld;
            op.a_zs = zs;
            op.B = M;
            op.ldb = ld;
            op.b_zs = zs;
            op.C = X + off21;
            op.ldc = ld;
            op.c_zs = zs;
            op.tiles_m = hb;
            op.tiles_m_last = last_s2;
            op.tiles_n = hb;
            op.klo_sel = KSEL_TJ;
            op.klo_c = 0;
            op.khi_sel = KSEL_CONST;
            op.khi_c = hb;
            GPP_TRY(launch_gemm(op, true, false, nb, st));
        }
        {   // M21 = -M22 * X21 ; M22 lower: k-blocks [0, ti+1)
            GemmOp op = gemm_default();
            op.A = M + off22;
            op.lda = ld;
            op.a_zs = zs;
            op.B = X + off21;
            op.ldb = ld;
            op.b_zs = zs;
            op.C = M + off21;
            op.ldc = ld;
            op.c_zs = zs;
            op.tiles_m = hb;
            op.tiles_m_last = last_s2;
            op.tiles_n = hb;
            op.klo_sel = KSEL_CONST;
            op.klo_c = 0;
            op.khi_sel = KSEL_TI;
            op.khi_c = 1;
            op.alpha = -1.0;
            GPP_TRY(launch_gemm(op, true, false, nb, st));
        }
    }
    return cudaSuccess;
}

// Kinv = M^T M (full symmetric storage); M lower: k-blocks [ti, T)
inline cudaError_t lauum_full(const double* M, double* Kinv, int ld, int T, cudaStream_t st) {
    GemmOp op = gemm_default();
    op.A = M;
    op.lda = ld;
    op.B = M;
    op.ldb = ld;
    op.C = Kinv;
    op.ldc = ld;
    op.map = MAP_TRI;
    op.tiles_m = op.tiles_m_last = T;
    op.tiles_n = T;
    op.klo_sel = KSEL_TI;
    op.klo_c = 0;
    op.khi_sel = KSEL_CONST;
    op.khi_c = T;
    op.mirror = 1;
    return launch_gemm(op, false, false, 1, st);
}

}  // namespace gpp
