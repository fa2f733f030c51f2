// K1 fused covariance builder and K3 fused MLL-gradient reduction for sm_100a.
//
// K1 replaces the chain  transform_categorical -> Linear_MAP -> cat -> covar_module(x).evaluate()
// -> likelihood(...)  (models/gp_plus.py:410-474, likelihoods_noise/multifidelity.py:63-136; the
// gpytorch Kernel.covar_dist quadratic expansion): one pass that writes K_y.
// K3 replaces autograd's backward through that chain and through cholesky
// (optim/mll_scipy.py:123): one streaming pass over W = alpha alpha^T - K_y^-1 that regenerates
// K and dK/dtheta tile by tile and reduces  -1/2 tr(W dK/dtheta)  for every hyper-parameter.
//
// Both kernels work on 128x128 tiles.  Per tile, the row / column point panels (centred inputs
// pre-scaled by sqrt(w_d), squared norms, latent coordinates) are staged with bulk TMA copies;
// the cross term x_i . x_j of the squared distance runs on the FP64 tensor cores (DMMA m8n8k4),
// the rest (clamp, exp / Matern polynomial, latent RBF factor, noise) stays in registers.
#pragma once
#include "dgemm_dmma.cuh"
#include "fastmath.cuh"
#include "tma.cuh"

namespace gpp {

constexpr int CT = 128;          // tile edge (points)
constexpr int COV_THREADS = 512;  // 16 warps, 4 x 4 grid of 32x32 warp tiles: 64 accumulator registers per thread
                                  // (the 8-warp / 32x64 layout ran at 12 % warp occupancy and was latency-bound)
constexpr int ZP = 4;            // latent-coordinate stride per point (GPP_MAX_DZ)

enum { KERNEL_EXPSQ = 0, KERNEL_MATERN32 = 1, KERNEL_MATERN52 = 2 };

// padded feature count: 4*odd so that fragment loads with row stride dqp are bank-conflict free
inline __host__ __device__ int pad_dq(int dq) {
    int q = (dq + 3) / 4;
    if (q < 1) q = 1;
    if ((q & 1) == 0) q++;
    return 4 * q;
}

// quantitative correlation f(s) and df/ds at s = sum_d w_d (x_id - x_jd)^2
template <int KIND>
__device__ __forceinline__ void kq_eval(double s, double& f, double& fp) {
    if (KIND == KERNEL_EXPSQ) {
        f = exp_nonpos(-s);
        fp = -f;
    } else if (KIND == KERNEL_MATERN32) {
        const double c = 1.7320508075688772;
        double r = sqrt(s);
        double e = exp_nonpos(-c * r);
        f = (1.0 + c * r) * e;
        fp = -1.5 * e;
    } else {
        const double c = 2.23606797749979;
        double r = sqrt(s);
        double e = exp_nonpos(-c * r);
        f = ((c * r + 1.0) + (5.0 / 3.0) * s) * e;
        fp = -(5.0 / 6.0) * (1.0 + c * r) * e;
    }
}

// ---------------------------------------------------------------------------------------------
// per-point preparation: xs = (x - centre) * sqrt(w), nrm = |xs|^2, zpt = Z[level_idx]
// (x.mul(sqrt(lengthscale)) / x.div(lengthscale) + the mean-centering of covar_dist)
struct PrepArgs {
    const double* xq;      // [n*dq]
    const int* level_idx;  // [n] or NULL
    const double* w;       // [dq]
    const double* centre;  // [dq]
    const double* ztab;    // [n_combo*dz]
    int n, np, dq, dqp, dz, n_combo;
    int n_pass;            // latent tables (multi-pass ensemble, gp_plus.py:387-399); ztab / zpt hold n_pass blocks
    double* xs;            // [np*dqp]
    double* xst;           // [dqp*np] the same values feature-major (gradient kernel panels) or NULL
    double* nrm;           // [np]
    double* zpt;           // [n_pass][np*ZP]
};

__global__ void prep_points_kernel(const PrepArgs a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.np) return;
    double nr = 0.0;
    double* xo = a.xs + (long long)i * a.dqp;
    if (i < a.n) {
        for (int d = 0; d < a.dq; d++) {
            double v = (a.xq[(long long)i * a.dq + d] - a.centre[d]) * sqrt(a.w[d]);
            xo[d] = v;
            if (a.xst) a.xst[(long long)d * a.np + i] = v;
            nr = fma(v, v, nr);
        }
        for (int d = a.dq; d < a.dqp; d++) {
            xo[d] = 0.0;
            if (a.xst) a.xst[(long long)d * a.np + i] = 0.0;
        }
    } else {
        for (int d = 0; d < a.dqp; d++) {
            xo[d] = 0.0;
            if (a.xst) a.xst[(long long)d * a.np + i] = 0.0;
        }
    }
    a.nrm[i] = nr;
    int lv = (i < a.n && a.level_idx && a.dz > 0) ? a.level_idx[i] : -1;
    const int passes = a.n_pass > 0 ? a.n_pass : 1;
    for (int p = 0; p < passes; p++)
        for (int k = 0; k < ZP; k++) {
            double z = 0.0;
            if (lv >= 0 && lv < a.n_combo && k < a.dz) z = a.ztab[((long long)p * a.n_combo + lv) * a.dz + k];
            a.zpt[((long long)p * a.np + i) * ZP + k] = z;
        }
}

// r = y - m(x), diag_add = noise[group] + jitter   (means: gp_plus.py:509-544, noise: multifidelity.py:105-136)
__global__ void prep_targets_kernel(const double* y, const int* mean_idx, const double* beta, int n_mean,
                                    const int* noise_idx, const double* noise, int n_noise, double jitter,
                                    const double* jitter_dev, int n, int np, double* r, double* diag_add) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    if (jitter_dev) jitter = *jitter_dev;  // graph replays read the per-evaluation scalars from device memory
    if (i < n) {
        int mi = mean_idx ? mean_idx[i] : 0;
        double m = (n_mean > 0 && mi >= 0 && mi < n_mean) ? beta[mi] : 0.0;
        r[i] = y[i] - m;
        int g = noise_idx ? noise_idx[i] : 0;
        double nz = (g >= 0 && g < n_noise) ? noise[g] : 0.0;
        diag_add[i] = nz + jitter;
    } else {
        r[i] = 0.0;
        diag_add[i] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
struct CovArgs {
    const double *xs_r, *nrm_r, *zpt_r;  // row points (padded to 128)
    const double *xs_c, *nrm_c, *zpt_c;  // column points
    double* out;
    long long ld;
    int n_r, n_c;            // valid rows / cols
    int tiles_r, tiles_c;
    int tri;                 // 1: lower tiles only (tiles_r == tiles_c)
    int same;                // rows and columns are the same point set
    int pad_identity;        // padded diagonal entries = 1 (training K_y)
    int dqp, dz;
    double sf2;
    const double* sf2_dev;   // when set, overrides sf2 (CUDA-graph replays keep kernel arguments fixed)
    const double* diag_add;  // [n] added on the diagonal when same (noise + jitter) or NULL
    const double* alpha;     // optional [cols]: mean_part[tj*ld_part + row] = sum_col K[row,col]*alpha[col]
    double* mean_part;
    long long ld_part;
    // multi-pass ensemble covariance (gp_plus.py:387-399, 474-482): K = (1/k) sum_p K_p; one launch per latent table
    int accum;               // add to what `out` / `mean_part` already hold
    int last;                // last pass: the diagonal additions / identity padding are applied now
    double scale;            // 1/k (1.0 for the single-pass path: multiplying by it is skipped)
};

inline size_t cov_smem_bytes(int dqp) {
    return 128 + sizeof(double) * (size_t)(2 * CT * dqp + 2 * CT + 2 * CT * ZP + CT + 4 * CT);
}

__device__ __forceinline__ void tri_decode(int bid, int& ti, int& tj) {
    int t = (int)((sqrt(8.0 * (double)bid + 1.0) - 1.0) * 0.5);
    while ((long long)(t + 1) * (t + 2) / 2 <= bid) t++;
    while ((long long)t * (t + 1) / 2 > bid) t--;
    ti = t;
    tj = bid - (int)((long long)t * (t + 1) / 2);
}

// cross term acc[mi][ni][e] = xs_i . xs_j on DMMA; warp (4 x 4 grid) owns rows wm0..+32, cols wn0..+32
__device__ __forceinline__ void tile_cross_dmma(const double* Xi, const double* Xj, int dqp, int wm0, int wn0, int g,
                                                int t, double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int kk = 0; kk < dqp; kk += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) af[mi] = Xi[(wm0 + mi * 8 + g) * dqp + kk + t];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) bf[ni] = Xj[(wn0 + ni * 8 + g) * dqp + kk + t];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
}

// sqrt of a positive, normal argument without the IEEE slow path of sqrt(): MUFU.RSQ64H seed (2^-22) and two coupled
// Goldschmidt steps (2^-44, 2^-88 before rounding): <= 2 ulp, 2 DMUL + 5 DFMA, no branch.
__device__ __forceinline__ double sqrt_pos(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double gq = s * y, hq = 0.5 * y;
    double rq = fma(-gq, hq, 0.5);
    gq = fma(gq, rq, gq);
    hq = fma(hq, rq, hq);
    rq = fma(-gq, hq, 0.5);
    return fma(gq, rq, gq);
}

// Squared distances come out of  |x_i|^2 + |x_j|^2 - 2 x_i.x_j  and can be slightly negative (clamp_min(0) in
// gpytorch's covar_dist).  For the Matern kernels the clamp floor is 1e-300 instead of 0 (sqrt_pos needs
// a normal argument): r = 1e-150 gives f = 1 and f' = f'(0) exactly, the values at s = 0.  Integer compare on the
// high word: negative, zero and sub-floor values all take the branch-free select, no FP64-pipe instruction.
template <int KIND>
__device__ __forceinline__ double clamp_dist(double s) {
    if (KIND == KERNEL_EXPSQ) return (__double2hiint(s) < 0) ? 0.0 : s;
    return (__double2hiint(s) < 0x01a00000) ? 1.0e-300 : s;
}

// quantitative correlation f(s) and df/ds for a clamped s (see clamp_dist); same formulas as kq_eval
template <int KIND>
__device__ __forceinline__ void kq_eval_fast(double s, double& f, double& fp) {
    if (KIND == KERNEL_EXPSQ) {
        f = exp_nonpos_dev(-s);
        fp = -f;
    } else if (KIND == KERNEL_MATERN32) {
        const double c = 1.7320508075688772;
        const double cr = c * sqrt_pos(s);
        const double e = exp_nonpos_dev(-cr);
        f = (1.0 + cr) * e;
        fp = -1.5 * e;
    } else {
        const double c = 2.23606797749979;
        const double cr = c * sqrt_pos(s);
        const double e = exp_nonpos_dev(-cr);
        const double q = cr + 1.0;
        f = fma(5.0 / 3.0, s, q) * e;
        fp = (-(5.0 / 6.0) * q) * e;
    }
}

// One 128x128 tile.  Each warp owns a 32x32 block and walks it in four 8-row slabs with a ROLLED loop: the slab's
// cross terms (DMMA), its 8 kernel values per thread and their stores.  Rolling keeps the live state at 8
// accumulators (the fully unrolled version held 64 plus re-materialised every FP64 constant at every use: 34 of ~125
// issued instructions per pair were LDCU / UMOV), so two 512-thread CTAs fit on an SM and one CTA's panel load
// overlaps the other's arithmetic.  GENERAL = false is the interior tile: no diagonal, no padding, single pass.
template <int KIND, bool HAS_Z, bool GENERAL>
__device__ __forceinline__ void cov_tile_body(const CovArgs& a, const int ti, const int tj, const double* Xi,
                                              const double* Xj, const double* sni, const double* snj, const double* zi,
                                              const double* zj, const double* sal, double* red) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
    const int dqp = a.dqp, dz = HAS_Z ? a.dz : 0;
    const bool diag_tile = GENERAL && a.same && (ti == tj);
    const bool has_alpha = a.alpha != nullptr;
    double sf2 = a.sf2_dev ? *a.sf2_dev : a.sf2;
    if (GENERAL && a.scale != 1.0) sf2 *= a.scale;
    double* outp = a.out ? a.out + (long long)ti * CT * a.ld + (long long)tj * CT + wn0 + 2 * t : nullptr;
    const double* xjp = Xj + (wn0 + g) * dqp + t;
    // column-side scalars of this thread's 8 columns (col = wn0 + ni * 8 + 2 t + e)
    double ncol[4][2];
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
        for (int e = 0; e < 2; e++) ncol[ni][e] = snj[wn0 + ni * 8 + 2 * t + e];
#pragma unroll 1
    for (int mi = 0; mi < 4; mi++) {
        const int row = wm0 + mi * 8 + g;
        double acc[4][2];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) { acc[ni][0] = 0.0; acc[ni][1] = 0.0; }
        const double* xip = Xi + row * dqp + t;
        for (int kk = 0; kk < dqp; kk += 4) {
            const double af = xip[kk];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) dmma884(acc[ni][0], acc[ni][1], af, xjp[ni * 8 * dqp + kk]);
        }
        const double nri = sni[row];
        const int gi = ti * CT + row;
        double zr[ZP];
#pragma unroll
        for (int k = 0; k < ZP; k++) zr[k] = (HAS_Z && k < dz) ? zi[row * ZP + k] : 0.0;
        double msum = 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double kv[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = wn0 + ni * 8 + 2 * t + e;
                double s = fma(-2.0, acc[ni][e], nri + ncol[ni][e]);
                if (GENERAL && diag_tile && row == col) s = 0.0;
                s = clamp_dist<KIND>(s);
                double sz = 0.0;
                if (HAS_Z) {
#pragma unroll
                    for (int k = 0; k < ZP; k++)
                        if (k < dz) {
                            const double dd = zr[k] - zj[col * ZP + k];
                            sz = fma(dd, dd, sz);
                        }
                }
                double kval;
                if (KIND == KERNEL_EXPSQ) {
                    kval = sf2 * exp_nonpos_dev(HAS_Z ? -fma(0.5, sz, s) : -s);
                } else {
                    double f, fp;
                    kq_eval_fast<KIND>(s, f, fp);
                    kval = sf2 * f;
                    if (HAS_Z) kval *= exp_nonpos_dev(-0.5 * sz);
                }
                if (GENERAL && (gi >= a.n_r || tj * CT + col >= a.n_c)) kval = 0.0;
                if (has_alpha) msum = fma(kval, sal[col], msum);
                kv[e] = kval;
            }
            if (outp) {
                double2* dst = reinterpret_cast<double2*>(outp + (long long)row * a.ld + ni * 8);
                double2 v;
                v.x = kv[0];
                v.y = kv[1];
                if (GENERAL) {
                    if (a.accum) {
                        const double2 o = *dst;
                        v.x += o.x;
                        v.y += o.y;
                    }
                    if (a.last) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int col = wn0 + ni * 8 + 2 * t + e;
                            const int gj = tj * CT + col;
                            const bool on_diag = diag_tile && (row == col);
                            double& x = e == 0 ? v.x : v.y;
                            if (on_diag && a.diag_add) x += a.diag_add[gi < a.n_r ? gi : 0];
                            if (gi >= a.n_r || gj >= a.n_c) x = (a.pad_identity && on_diag) ? 1.0 : 0.0;
                        }
                    }
                }
                *dst = v;
            }
        }
        if (has_alpha) {
            msum += __shfl_xor_sync(0xffffffffu, msum, 1);
            msum += __shfl_xor_sync(0xffffffffu, msum, 2);
            if (t == 0) red[(warp & 3) * CT + row] = msum;
        }
    }
}

template <int KIND, bool HAS_Z>
__global__ void __launch_bounds__(COV_THREADS, 2) cov_tile_kernel(const CovArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int dqp = a.dqp, dz = HAS_Z ? a.dz : 0;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
    double* Xi = reinterpret_cast<double*>(smraw + 128);
    double* Xj = Xi + CT * dqp;
    double* sni = Xj + CT * dqp;
    double* snj = sni + CT;
    double* zi = snj + CT;
    double* zj = zi + CT * ZP;
    double* sal = zj + CT * ZP;  // alpha of the column points
    double* red = sal + CT;      // [4][CT]

    const int tid = threadIdx.x;
    int ti, tj;
    if (a.tri) {
        tri_decode(blockIdx.x, ti, tj);
    } else {
        ti = blockIdx.x / a.tiles_c;
        tj = blockIdx.x - ti * a.tiles_c;
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t xb = (uint32_t)(CT * dqp * 8), nb = CT * 8, zb = CT * ZP * 8;
        uint32_t total = 2 * xb + 2 * nb + (dz > 0 ? 2 * zb : 0) + (a.alpha ? nb : 0);
        mbar_arrive_expect_tx(bar, total);
        tma_load_1d(Xi, a.xs_r + (long long)ti * CT * dqp, xb, bar);
        tma_load_1d(Xj, a.xs_c + (long long)tj * CT * dqp, xb, bar);
        tma_load_1d(sni, a.nrm_r + (long long)ti * CT, nb, bar);
        tma_load_1d(snj, a.nrm_c + (long long)tj * CT, nb, bar);
        if (dz > 0) {
            tma_load_1d(zi, a.zpt_r + (long long)ti * CT * ZP, zb, bar);
            tma_load_1d(zj, a.zpt_c + (long long)tj * CT * ZP, zb, bar);
        }
        if (a.alpha) tma_load_1d(sal, a.alpha + (long long)tj * CT, nb, bar);
    }
    mbar_wait(bar, 0);

    const bool general = (a.same && ti == tj) || (ti + 1) * CT > a.n_r || (tj + 1) * CT > a.n_c || a.accum || !a.last ||
                         a.scale != 1.0;
    if (general) cov_tile_body<KIND, HAS_Z, true>(a, ti, tj, Xi, Xj, sni, snj, zi, zj, sal, red);
    else cov_tile_body<KIND, HAS_Z, false>(a, ti, tj, Xi, Xj, sni, snj, zi, zj, sal, red);

    if (a.alpha) {
        __syncthreads();
        if (tid < CT) {
            double* mp = a.mean_part + (long long)tj * a.ld_part + (long long)ti * CT + tid;
            const double v = (red[tid] + red[CT + tid]) + (red[2 * CT + tid] + red[3 * CT + tid]);
            *mp = a.accum ? (*mp + v) : v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct GradArgs {
    const double *xst, *nrm, *zpt;  // training points: scaled inputs feature-major [dqp][np], norms, latent coords
    const double* alpha;            // [np]
    const double* Kinv;            // [np*ld] lower tiles of K_y^-1 (the upper triangle is never read)
    long long ld;
    int n, np, T;
    int dq, dqp, dz;
    double sf2;
    const double* sf2_dev;  // when set, overrides sf2
    double* tile_part;  // [ntiles * (1 + dqp)]: sum W*Kc, sum Q*dx_d^2 (d < dq)
    double* zpart;      // [T * np * ZP]: slot (t, p): partial of sum_j P_pj (z_p - z_j)
};

// column slabs of a warp block handled per phase-1 / phase-2 round: all four when the Q stash (32 doubles per thread)
// fits beside the point panels, two for wide inputs (dqp > 20)
inline __host__ __device__ int grad_slabs(int dqp) { return dqp <= 20 ? 4 : 2; }

// Row stride of the feature-major point panels in shared memory.  132 = 4 (mod 16): the DMMA fragment loads
// (lane (g, t) reads feature kk + t of point row0 + g) hit 32 distinct banks per half warp, and the per-feature reads
// of phase 2 (consecutive points) are contiguous, so neither phase has a bank conflict (the point-major panels gave
// 2.8 wavefronts per shared load: r02 ncu, "Shared is the highest-utilized pipeline (56.7 %)").
constexpr int GLD = 132;

inline size_t grad_smem_bytes(int dqp) {
    // bar | XiT XjT [dqp][GLD] | ni nj | zi zj | ai aj | rowacc[4][CT][ZP] | colacc[4][CT][ZP] | wred[16][1+dqp] |
    // qst[8 slabs][512]
    return 128 + sizeof(double) * (size_t)(2 * GLD * dqp + 2 * CT + 2 * CT * ZP + 2 * CT + 4 * CT * ZP + 4 * CT * ZP +
                                           16 * (1 + dqp) + 8 * grad_slabs(dqp) * COV_THREADS);
}

// Phase 1 of a gradient tile (rolled over the four 8-column slabs of the warp's 32x32 block): cross terms on DMMA,
// kernel value and derivative per pair, W = alpha_i alpha_j - (K^-1)_ij from the prefetched slab of the W tile,
// sum W K, the latent-coordinate sums, and Q = W sf2 f' parked in shared memory (thread-private slots) for phase 2.
// GENERAL = false is the interior tile (no diagonal, no padding).
template <int KIND, int ZD, bool GENERAL>
__device__ __forceinline__ double grad_tile_pairs(const GradArgs& a, const int ti, const int tj, const double* Xi,
                                                  const double* Xj, const double* sni, const double* snj,
                                                  const double* zi, const double* zj, const double* sai,
                                                  const double* saj, double* rowacc, double* colacc, double* qst,
                                                  const int ni0, const int ni1) {
    constexpr bool HAS_Z = ZD > 0;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
    const int dqp = a.dqp, dz = HAS_Z ? a.dz : 0;
    const bool diag_tile = GENERAL && (ti == tj);
    const double sf2 = a.sf2_dev ? *a.sf2_dev : a.sf2;
    const double* Wp = a.Kinv + (long long)(ti * CT + wm0 + g) * a.ld + (long long)tj * CT + wn0 + 2 * t;
    const long long ld8 = 8 * a.ld;

    double nrow[4], arow[4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        nrow[mi] = sni[wm0 + mi * 8 + g];
        arow[mi] = sai[wm0 + mi * 8 + g];
    }
    double rowz[4][HAS_Z ? ZD : 1];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int k = 0; k < (HAS_Z ? ZD : 1); k++) rowz[mi][k] = 0.0;
    double sumWK = 0.0;
    double2 kin[4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) kin[mi] = *reinterpret_cast<const double2*>(Wp + mi * ld8 + ni0 * 8);

#pragma unroll 1
    for (int ni = ni0; ni < ni1; ni++) {
        double2 kcur[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) kcur[mi] = kin[mi];
        if (ni + 1 < ni1) {
#pragma unroll
            for (int mi = 0; mi < 4; mi++) kin[mi] = *reinterpret_cast<const double2*>(Wp + mi * ld8 + (ni + 1) * 8);
        }
        double acc[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) { acc[mi][0] = 0.0; acc[mi][1] = 0.0; }
        const double* xjp = Xj + t * GLD + wn0 + ni * 8 + g;
        const double* xip = Xi + t * GLD + wm0 + g;
        for (int kk = 0; kk < dqp; kk += 4) {
            const double bf = xjp[kk * GLD];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) dmma884(acc[mi][0], acc[mi][1], xip[kk * GLD + mi * 8], bf);
        }
        const int col0 = wn0 + ni * 8 + 2 * t;
        double ncol[2], acol[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            ncol[e] = snj[col0 + e];
            acol[e] = saj[col0 + e];
        }
        double colz[2][HAS_Z ? ZD : 1];
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int k = 0; k < (HAS_Z ? ZD : 1); k++) colz[e][k] = 0.0;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int row = wm0 + mi * 8 + g;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = col0 + e;
                double s = fma(-2.0, acc[mi][e], nrow[mi] + ncol[e]);
                if (GENERAL && diag_tile && row == col) s = 0.0;
                s = clamp_dist<KIND>(s);
                double dzk[HAS_Z ? ZD : 1];
                double sz = 0.0;
#pragma unroll
                for (int k = 0; k < ZD; k++) {
                    dzk[k] = (k < dz) ? (zi[row * ZP + k] - zj[col * ZP + k]) : 0.0;
                    sz = fma(dzk[k], dzk[k], sz);
                }
                double f, fp;
                if (KIND == KERNEL_EXPSQ) {
                    f = exp_nonpos_dev(HAS_Z ? -fma(0.5, sz, s) : -s);
                    fp = -f;
                } else {
                    kq_eval_fast<KIND>(s, f, fp);
                    if (HAS_Z) {
                        const double kz = exp_nonpos_dev(-0.5 * sz);
                        f *= kz;
                        fp *= kz;
                    }
                }
                double W = fma(arow[mi], acol[e], -(e == 0 ? kcur[mi].x : kcur[mi].y));
                if (GENERAL && (ti * CT + row >= a.n || tj * CT + col >= a.n)) W = 0.0;
                const double WK = W * f;
                sumWK += WK;
                qst[(((ni - ni0) * 4 + mi) * 2 + e) * COV_THREADS + tid] = (W * sf2) * fp;  // Q
                if (HAS_Z) {
                    const double P = WK * sf2;
#pragma unroll
                    for (int k = 0; k < ZD; k++)
                        if (k < dz) {
                            rowz[mi][k] = fma(P, dzk[k], rowz[mi][k]);
                            colz[e][k] = fma(-P, dzk[k], colz[e][k]);
                        }
                }
            }
        }
        if (HAS_Z && dz > 0 && !(ti == tj)) {
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int k = 0; k < ZD; k++)
                    if (k < dz) {
                        double v = colz[e][k];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (g == 0) colacc[((warp >> 2) * CT + col0 + e) * ZP + k] = v;
                    }
        }
    }
    if (HAS_Z && dz > 0) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int k = 0; k < ZD; k++)
                if (k < dz) {
                    double v = rowz[mi][k];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (t == 0) {
                        double* ra = rowacc + ((warp & 3) * CT + wm0 + mi * 8 + g) * ZP + k;
                        *ra = ni0 == 0 ? v : *ra + v;
                    }
                }
    }
    return sumWK;
}

// Phase 2: sum Q (xs_id - xs_jd)^2 per input dimension for the NI column slabs parked by phase 1
template <int NI>
__device__ __forceinline__ void grad_tile_dims(const GradArgs& a, const double* Xi, const double* Xj, const double* qst,
                                               double* wred, const int ni0) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
    double q[NI][4][2];
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int e = 0; e < 2; e++) q[ni][mi][e] = qst[((ni * 4 + mi) * 2 + e) * COV_THREADS + tid];
#pragma unroll 1
    for (int d = 0; d < a.dq; d++) {
        double xi_d[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) xi_d[mi] = Xi[d * GLD + wm0 + mi * 8 + g];
        double sdm[4] = {0.0, 0.0, 0.0, 0.0};  // one chain per row slab: four independent DFMA chains of 2 NI links
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const double2 xj2 = *reinterpret_cast<const double2*>(Xj + d * GLD + wn0 + (ni0 + ni) * 8 + 2 * t);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const double xj = e == 0 ? xj2.x : xj2.y;
#pragma unroll
                for (int mi = 0; mi < 4; mi++) {
                    const double df = xi_d[mi] - xj;
                    sdm[mi] = fma(q[ni][mi][e], df * df, sdm[mi]);
                }
            }
        }
        double sd = (sdm[0] + sdm[1]) + (sdm[2] + sdm[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
        if (lane == 0) {
            double* wr = wred + warp * (1 + a.dqp) + 1 + d;
            *wr = ni0 == 0 ? sd : *wr + sd;
        }
    }
}

// ZD: compile-time bound on the latent dimension (0 = no latent map, 2 = the default two-dimensional map, 4 = GPP_MAX_DZ)
template <int KIND, int ZD>
__global__ void __launch_bounds__(COV_THREADS, 1) grad_tile_kernel(const GradArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int dqp = a.dqp, dz = ZD > 0 ? a.dz : 0;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
    double* Xi = reinterpret_cast<double*>(smraw + 128);  // [dqp][GLD] feature-major
    double* Xj = Xi + GLD * dqp;
    double* sni = Xj + GLD * dqp;
    double* snj = sni + CT;
    double* zi = snj + CT;
    double* zj = zi + CT * ZP;
    double* sai = zj + CT * ZP;
    double* saj = sai + CT;
    double* rowacc = saj + CT;              // [4][CT][ZP]
    double* colacc = rowacc + 4 * CT * ZP;  // [4][CT][ZP]
    double* wred = colacc + 4 * CT * ZP;    // [16][1+dqp]
    double* qst = wred + 16 * (1 + dqp);    // [32][COV_THREADS]

    const int tid = threadIdx.x;
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t xb = (uint32_t)(CT * dqp * 8), nb = CT * 8, zb = CT * ZP * 8;
        uint32_t total = 2 * xb + 4 * nb + (dz > 0 ? 2 * zb : 0);
        mbar_arrive_expect_tx(bar, total);
        for (int d = 0; d < dqp; d++) {  // one 1 KB row per feature and panel
            tma_load_1d(Xi + d * GLD, a.xst + (long long)d * a.np + (long long)ti * CT, nb, bar);
            tma_load_1d(Xj + d * GLD, a.xst + (long long)d * a.np + (long long)tj * CT, nb, bar);
        }
        tma_load_1d(sni, a.nrm + (long long)ti * CT, nb, bar);
        tma_load_1d(snj, a.nrm + (long long)tj * CT, nb, bar);
        tma_load_1d(sai, a.alpha + (long long)ti * CT, nb, bar);
        tma_load_1d(saj, a.alpha + (long long)tj * CT, nb, bar);
        if (dz > 0) {
            tma_load_1d(zi, a.zpt + (long long)ti * CT * ZP, zb, bar);
            tma_load_1d(zj, a.zpt + (long long)tj * CT * ZP, zb, bar);
        }
    }
    mbar_wait(bar, 0);

    const int warp = tid >> 5, lane = tid & 31;
    const bool diag_tile = (ti == tj);
    const bool general = diag_tile || (ti + 1) * CT > a.n;  // tj <= ti: the row bound covers the columns
    const int slabs = grad_slabs(dqp);
    double sumWK = 0.0;
    for (int ni0 = 0; ni0 < 4; ni0 += slabs) {
        if (ni0 > 0) __syncwarp();  // the stash slots are thread-private; keeps the warp converged between rounds
        if (general)
            sumWK += grad_tile_pairs<KIND, ZD, true>(a, ti, tj, Xi, Xj, sni, snj, zi, zj, sai, saj, rowacc, colacc,
                                                        qst, ni0, ni0 + slabs);
        else
            sumWK += grad_tile_pairs<KIND, ZD, false>(a, ti, tj, Xi, Xj, sni, snj, zi, zj, sai, saj, rowacc, colacc,
                                                         qst, ni0, ni0 + slabs);
        if (slabs == 4) grad_tile_dims<4>(a, Xi, Xj, qst, wred, ni0);
        else grad_tile_dims<2>(a, Xi, Xj, qst, wred, ni0);
    }
    // sum W*Kc over the warp
    {
        double v = sumWK;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wred[warp * (1 + dqp)] = v;
    }
    __syncthreads();
    if (tid < 1 + a.dq) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 16; w++) v += wred[w * (1 + dqp) + tid];
        a.tile_part[(long long)blockIdx.x * (1 + dqp) + tid] = v;
    }
    if (dz > 0 && tid < CT) {
        // slot (tj, point of block ti): row partials; slot (ti, point of block tj): column partials
        for (int k = 0; k < dz; k++) {
            double v = (rowacc[(0 * CT + tid) * ZP + k] + rowacc[(1 * CT + tid) * ZP + k]) +
                       (rowacc[(2 * CT + tid) * ZP + k] + rowacc[(3 * CT + tid) * ZP + k]);
            a.zpart[((long long)tj * a.np + (long long)ti * CT + tid) * ZP + k] = v;
            if (!diag_tile) {
                double c = colacc[(0 * CT + tid) * ZP + k] + colacc[(1 * CT + tid) * ZP + k] +
                           colacc[(2 * CT + tid) * ZP + k] + colacc[(3 * CT + tid) * ZP + k];
                a.zpart[((long long)ti * a.np + (long long)tj * CT + tid) * ZP + k] = c;
            }
        }
    }
}

// gz[p][k] = sum_t zpart[t][p][k]
__global__ void zpart_reduce_kernel(const double* zpart, int T, int np, int n, int dz, double* gz, double scale) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * dz) return;
    int p = idx / dz, k = idx - p * dz;
    double s = 0.0;
    for (int t = 0; t < T; t++) s += zpart[((long long)t * np + p) * ZP + k];
    gz[idx] = (scale != 1.0) ? s * scale : s;
}

// ---------------------------------------------------------------------------------------------
// deterministic block-wide sum (blockDim.x == 256)
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) r += sh[w];
    return r;
}

struct FinishArgs {
    const double* v;            // [np]  L^-1 r
    const double* alpha;        // [np]
    const double* logdet_part;  // [T]
    const double* Kinv;         // diagonal read, or NULL when no gradient
    long long ld;
    const double* tile_part;    // or NULL
    const double* w;            // [dq]
    const int* noise_idx;
    const int* mean_idx;
    const int* info;
    int n, np, T, dq, dqp, n_noise, n_mean, want_grad;
    int n_pass;                 // tile_part holds n_pass blocks (one gradient pass per latent table); averaged here
    double* res;  // [0]=quad [1]=logdet [2]=info [3]=d_sf2 [4..4+dq)=d_w, then d_noise[n_noise], d_beta[n_mean]
};

__global__ void __launch_bounds__(256) finish_kernel(const FinishArgs a) {
    __shared__ double sh[8];
    const int tid = threadIdx.x;
    double q = 0.0;
    for (int i = tid; i < a.n; i += 256) q = fma(a.v[i], a.v[i], q);
    q = block_sum_256(q, sh);
    double ld = 0.0;
    for (int k = tid; k < a.T; k += 256) ld += a.logdet_part[k];
    ld = block_sum_256(ld, sh);
    if (tid == 0) {
        a.res[0] = q;
        a.res[1] = 2.0 * ld;
        a.res[2] = (double)(*a.info);
    }
    if (!a.want_grad) return;
    const int ntiles = a.T * (a.T + 1) / 2;
    const int tp = 1 + a.dqp;
    const int passes = a.n_pass > 0 ? a.n_pass : 1;
    for (int c = 0; c < 1 + a.dq; c++) {
        double s = 0.0;
        // tile id b = ti(ti+1)/2 + tj ; diagonal tiles (tj == ti) weigh 1, the others 2
        for (int p = 0; p < passes; p++)
            for (int b = tid; b < ntiles; b += 256) {
                int ti, tj;
                tri_decode(b, ti, tj);
                double wgt = (ti == tj) ? 1.0 : 2.0;
                s = fma(wgt, a.tile_part[((long long)p * ntiles + b) * tp + c], s);
            }
        s = block_sum_256(s, sh);
        if (passes > 1) s /= (double)passes;
        if (tid == 0) {
            if (c == 0) a.res[3] = -0.5 * s;
            else a.res[3 + c] = -0.5 * s / a.w[c - 1];
        }
    }
    for (int gsel = 0; gsel < a.n_noise; gsel++) {
        double s = 0.0;
        for (int i = tid; i < a.n; i += 256) {
            int gi = a.noise_idx ? a.noise_idx[i] : 0;
            if (gi == gsel) s += a.alpha[i] * a.alpha[i] - a.Kinv[(long long)i * a.ld + i];
        }
        s = block_sum_256(s, sh);
        if (tid == 0) a.res[4 + a.dq + gsel] = -0.5 * s;
    }
    for (int msel = 0; msel < a.n_mean; msel++) {
        double s = 0.0;
        for (int i = tid; i < a.n; i += 256) {
            int mi = a.mean_idx ? a.mean_idx[i] : 0;
            if (mi == msel) s += a.alpha[i];
        }
        s = block_sum_256(s, sh);
        if (tid == 0) a.res[4 + a.dq + a.n_noise + msel] = -s;
    }
}

inline cudaError_t cov_set_attributes() {
    // the largest request over the supported widths (the Q stash makes grad_smem_bytes non-monotonic in dqp)
    int maxb = 0;
    for (int dq = 0; dq <= 32; dq++) {
        const int b = (int)grad_smem_bytes(pad_dq(dq));
        if (b > maxb) maxb = b;
    }
    cudaError_t e;
#define GPP_SET(k)                                                                    \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb); \
    if (e != cudaSuccess) return e;
    GPP_SET((cov_tile_kernel<KERNEL_EXPSQ, true>))
    GPP_SET((cov_tile_kernel<KERNEL_MATERN32, true>))
    GPP_SET((cov_tile_kernel<KERNEL_MATERN52, true>))
    GPP_SET((cov_tile_kernel<KERNEL_EXPSQ, false>))
    GPP_SET((cov_tile_kernel<KERNEL_MATERN32, false>))
    GPP_SET((cov_tile_kernel<KERNEL_MATERN52, false>))
    GPP_SET((grad_tile_kernel<KERNEL_EXPSQ, 2>))
    GPP_SET((grad_tile_kernel<KERNEL_EXPSQ, 4>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN32, 2>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN32, 4>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN52, 2>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN52, 4>))
    GPP_SET((grad_tile_kernel<KERNEL_EXPSQ, 0>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN32, 0>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN52, 0>))
#undef GPP_SET
    return cudaSuccess;
}

inline cudaError_t launch_cov(const CovArgs& a, int kind, cudaStream_t st) {
    int nt = a.tri ? a.tiles_r * (a.tiles_r + 1) / 2 : a.tiles_r * a.tiles_c;
    if (nt <= 0) return cudaSuccess;
    size_t sm = cov_smem_bytes(a.dqp);
    count_launch();
    if (a.dz > 0) {
        if (kind == KERNEL_EXPSQ) cov_tile_kernel<KERNEL_EXPSQ, true><<<nt, COV_THREADS, sm, st>>>(a);
        else if (kind == KERNEL_MATERN32) cov_tile_kernel<KERNEL_MATERN32, true><<<nt, COV_THREADS, sm, st>>>(a);
        else cov_tile_kernel<KERNEL_MATERN52, true><<<nt, COV_THREADS, sm, st>>>(a);
    } else {
        if (kind == KERNEL_EXPSQ) cov_tile_kernel<KERNEL_EXPSQ, false><<<nt, COV_THREADS, sm, st>>>(a);
        else if (kind == KERNEL_MATERN32) cov_tile_kernel<KERNEL_MATERN32, false><<<nt, COV_THREADS, sm, st>>>(a);
        else cov_tile_kernel<KERNEL_MATERN52, false><<<nt, COV_THREADS, sm, st>>>(a);
    }
    return cudaGetLastError();
}

template <int ZD>
inline void launch_grad_zd(const GradArgs& a, int kind, int nt, size_t sm, cudaStream_t st) {
    if (kind == KERNEL_EXPSQ) grad_tile_kernel<KERNEL_EXPSQ, ZD><<<nt, COV_THREADS, sm, st>>>(a);
    else if (kind == KERNEL_MATERN32) grad_tile_kernel<KERNEL_MATERN32, ZD><<<nt, COV_THREADS, sm, st>>>(a);
    else grad_tile_kernel<KERNEL_MATERN52, ZD><<<nt, COV_THREADS, sm, st>>>(a);
}

inline cudaError_t launch_grad(const GradArgs& a, int kind, cudaStream_t st) {
    int nt = a.T * (a.T + 1) / 2;
    size_t sm = grad_smem_bytes(a.dqp);
    count_launch();
    if (a.dz <= 0) launch_grad_zd<0>(a, kind, nt, sm, st);
    else if (a.dz <= 2) launch_grad_zd<2>(a, kind, nt, sm, st);
    else launch_grad_zd<4>(a, kind, nt, sm, st);
    return cudaGetLastError();
}

}  // namespace gpp
