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
constexpr int LB = 32;     // sub-block edge inside a 128x128 leaf
constexpr int LLD = 132;   // smem leading dimension: 132 = 4 (mod 16) keeps every DMMA fragment load conflict-free
constexpr int TLD = 36;    // per-warp scratch leading dimension (same residue)
constexpr int LEAF_SMEM_BYTES = (TILE * LLD + TILE + 3 * LB * TLD) * 8;
constexpr int PANEL_BLOCKS = 4;
constexpr unsigned FULL = 0xffffffffu;

// C(32x32) += A(32x32) * B(32x32) by one warp on DMMA; A(r,k) and B(k,n) are element getters.
// acc[mi][ni][e] holds C(mi*8+g, ni*8+2t+e).
template <class FA, class FB>
__device__ __forceinline__ void warp_mm32(double (&acc)[4][4][2], FA getA, FB getB, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < LB; kk += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) af[mi] = getA(mi * 8 + g, kk + t);
#pragma unroll
        for (int ni = 0; ni < 4; ni++) bf[ni] = getB(kk + t, ni * 8 + g);
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
}

__device__ __forceinline__ void acc_zero(double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// One warp factors the 32x32 diagonal sub-block at offset b (lane i owns row i in registers, pivots travel by
// shuffle) and inverts the factor (lane c owns column c of the inverse; L is re-read from shared memory with
// warp-uniform addresses).  On exit the lower triangle of the sub-block holds L, its strict upper triangle holds
// (L^-1)^T and xd[b+i] = 1/L_ii.  log L_jj is accumulated as (mantissa product, exponent sum): one log per block.
__device__ __forceinline__ void warp_potrf_inv32(double* S, double* xd, int b, int lane, int& bad, double& logacc) {
    double mydinv = 0.0;
    {
        double a[LB];
#pragma unroll
        for (int k = 0; k < LB; k++) a[k] = (k <= lane) ? S[(b + lane) * LLD + b + k] : 0.0;
        double mant = 1.0;
        int esum = 0;
#pragma unroll
        for (int j = 0; j < LB; j++) {
            const double d = __shfl_sync(FULL, a[j], j);
            if (!(d > 0.0)) bad |= (d != d) ? 2 : 1;
            const double l = sqrt(d);
            const double inv = 1.0 / l;
            int ex;
            mant *= frexp(l, &ex);
            esum += ex;
            if (lane == j) mydinv = inv;
            const double lij = (lane == j) ? l : ((lane > j) ? a[j] * inv : 0.0);
            a[j] = lij;
#pragma unroll
            for (int k = j + 1; k < LB; k++) {
                const double lkj = __shfl_sync(FULL, lij, k);
                if (lane >= k) a[k] = fma(-lij, lkj, a[k]);
            }
        }
        logacc += log(mant) + (double)esum * 0.6931471805599453;
#pragma unroll
        for (int k = 0; k < LB; k++)
            if (k <= lane) S[(b + lane) * LLD + b + k] = a[k];
        xd[b + lane] = mydinv;
    }
    __syncwarp();
    // X = L^-1: X[i][c] = -(sum_{k=c}^{i-1} L[i][k] X[k][c]) / L[i][i]
    double x[LB];
#pragma unroll
    for (int i = 0; i < LB; i++) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) {
            const double lik = S[(b + i) * LLD + b + k];
            if (k & 1) s1 = fma(lik, x[k], s1);
            else s0 = fma(lik, x[k], s0);
        }
        const double di = xd[b + i];
        x[i] = (lane == i) ? di : ((lane < i) ? -(s0 + s1) * di : 0.0);
    }
#pragma unroll
    for (int k = 0; k < LB; k++)
        if (k > lane) S[(b + lane) * LLD + b + k] = x[k];
}

// info[0]: 0 ok, 1 = non-positive pivot, 2 = NaN pivot.  logdet_part[kb] = sum_j log L_jj of block kb.
// Blocked (4 x 32) right-looking factorisation of the 128x128 diagonal block kb of A plus its explicit inverse
// (written to the diagonal block of M): diagonal sub-blocks by one warp in registers, every 32^3 product
// (TRSM through the sub-block inverse, SYRK updates, block forward substitution of the inverse) by one warp on DMMA.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_trinv_kernel(double* A, int ld, int kb, double* M, double* logdet_part, int* info) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                 // [128][LLD]: lower = L, strict upper = (L^-1)^T
    double* xd = sm + TILE * LLD;   // [128] 1/L_jj
    double* Tall = xd + TILE;       // [3][32][TLD] per-warp scratch
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    double* Ab = A + (long long)kb * TILE * ld + (long long)kb * TILE;
    double* Mb = M + (long long)kb * TILE * ld + (long long)kb * TILE;

    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        S[i * LLD + j] = (j <= i) ? Ab[(long long)i * ld + j] : 0.0;
    }
    __syncthreads();

    int bad = 0;
    double logacc = 0.0;
    double acc[4][4][2];
    for (int q = 0; q < 4; q++) {
        const int b = q * LB;
        if (warp == 0) warp_potrf_inv32(S, xd, b, lane, bad, logacc);
        __syncthreads();
        if (warp < 3 - q) {  // TRSM: L_ib,q = A_ib,q * X_q^T
            const int r0 = (q + 1 + warp) * LB;
            acc_zero(acc);
            warp_mm32(acc, [&](int r, int k) { return S[(r0 + r) * LLD + b + k]; },
                      [&](int k, int c) { return k < c ? S[(b + k) * LLD + b + c] : (k == c ? xd[b + c] : 0.0); }, g, t);
            __syncwarp();
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
#pragma unroll
                    for (int e = 0; e < 2; e++) S[(r0 + mi * 8 + g) * LLD + b + ni * 8 + 2 * t + e] = acc[mi][ni][e];
        }
        __syncthreads();
        {   // SYRK: A_ib,jb -= L_ib,q L_jb,q^T for q < jb <= ib <= 3 (one warp per block pair)
            const int m = 3 - q;  // remaining block rows
            if (warp < m * (m + 1) / 2) {
                int u = 0, w = warp;
                while (w > u) { w -= u + 1; u++; }  // warp -> (u, w), w <= u
                const int i0 = (q + 1 + u) * LB, j0 = (q + 1 + w) * LB;
                acc_zero(acc);
                warp_mm32(acc, [&](int r, int k) { return S[(i0 + r) * LLD + b + k]; },
                          [&](int k, int c) { return S[(j0 + c) * LLD + b + k]; }, g, t);
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = mi * 8 + g, c = ni * 8 + 2 * t + e;
                            if (i0 != j0 || c <= r) S[(i0 + r) * LLD + j0 + c] -= acc[mi][ni][e];
                        }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        logdet_part[kb] = logacc;
        if (bad) atomicOr(info, bad);
    }

    // off-diagonal 32-blocks of X = L^-1 by block forward substitution, one block-diagonal distance at a time:
    //   X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj        (X[r][c] is stored at S[c][r])
    for (int dist = 1; dist < 4; dist++) {
        if (warp < 4 - dist) {
            const int j = warp, i = warp + dist;
            double* T = Tall + warp * LB * TLD;
            acc_zero(acc);
            for (int k = j; k < i; k++) {
                if (k == j) {
                    warp_mm32(acc, [&](int r, int kk) { return S[(i * LB + r) * LLD + k * LB + kk]; },
                              [&](int kk, int n) {
                                  return n < kk ? S[(j * LB + n) * LLD + j * LB + kk] : (n == kk ? xd[j * LB + kk] : 0.0);
                              }, g, t);
                } else {
                    warp_mm32(acc, [&](int r, int kk) { return S[(i * LB + r) * LLD + k * LB + kk]; },
                              [&](int kk, int n) { return S[(j * LB + n) * LLD + k * LB + kk]; }, g, t);
                }
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
#pragma unroll
                    for (int e = 0; e < 2; e++) T[(mi * 8 + g) * TLD + ni * 8 + 2 * t + e] = acc[mi][ni][e];
            __syncwarp();
            acc_zero(acc);
            warp_mm32(acc, [&](int r, int kk) {
                          return kk < r ? S[(i * LB + kk) * LLD + i * LB + r] : (kk == r ? xd[i * LB + r] : 0.0);
                      },
                      [&](int kk, int n) { return T[kk * TLD + n]; }, g, t);
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
#pragma unroll
                    for (int e = 0; e < 2; e++)
                        S[(j * LB + ni * 8 + 2 * t + e) * LLD + i * LB + mi * 8 + g] = -acc[mi][ni][e];
        }
        __syncthreads();
    }

    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        int i = idx >> 7, c = idx & 127;
        if (c <= i) Ab[(long long)i * ld + c] = S[i * LLD + c];
        Mb[(long long)i * ld + c] = (c < i) ? S[c * LLD + i] : ((c == i) ? xd[i] : 0.0);
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
            count_launch();
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
