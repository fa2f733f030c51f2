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
            nr = fma(v, v, nr);
        }
        for (int d = a.dq; d < a.dqp; d++) xo[d] = 0.0;
    } else {
        for (int d = 0; d < a.dqp; d++) xo[d] = 0.0;
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

template <int KIND, bool HAS_Z>
__global__ void __launch_bounds__(COV_THREADS, 1) cov_tile_kernel(const CovArgs a) {
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

    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
    double acc[4][4][2];
    tile_cross_dmma(Xi, Xj, dqp, wm0, wn0, g, t, acc);

    const bool diag_tile = a.same && (ti == tj);
    const double sf2 = a.sf2_dev ? *a.sf2_dev : a.sf2;
    double* outp = a.out + (long long)ti * CT * a.ld + (long long)tj * CT;
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int row = wm0 + mi * 8 + g;
        const int gi = ti * CT + row;
        const double nri = sni[row];
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
                const int gj = tj * CT + col;
                double s = fmax(nri + snj[col] - 2.0 * acc[mi][ni][e], 0.0);
                const bool on_diag = diag_tile && (row == col);
                if (on_diag) s = 0.0;
                double sz = 0.0;
                if (HAS_Z) {
#pragma unroll
                    for (int k = 0; k < ZP; k++) {
                        if (k < dz) {
                            double dd = zr[k] - zj[col * ZP + k];
                            sz = fma(dd, dd, sz);
                        }
                    }
                }
                double kval;
                if (KIND == KERNEL_EXPSQ) {
                    kval = sf2 * exp_nonpos(-(s + 0.5 * sz));
                } else {
                    double f, fp;
                    kq_eval<KIND>(s, f, fp);
                    kval = sf2 * f;
                    if (HAS_Z) kval *= exp_nonpos(-0.5 * sz);
                }
                if (a.scale != 1.0) kval *= a.scale;
                if (gi >= a.n_r || gj >= a.n_c) kval = 0.0;
                if (a.alpha) msum = fma(kval, sal[col], msum);
                kv[e] = kval;
            }
            if (a.out) {
                double2* dst = reinterpret_cast<double2*>(outp + (long long)row * a.ld + wn0 + ni * 8 + 2 * t);
                double2 v;
                v.x = kv[0];
                v.y = kv[1];
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
                *dst = v;
            }
        }
        if (a.alpha) {
            msum += __shfl_xor_sync(0xffffffffu, msum, 1);
            msum += __shfl_xor_sync(0xffffffffu, msum, 2);
            if (t == 0) red[(warp & 3) * CT + row] = msum;
        }
    }
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
    const double *xs, *nrm, *zpt;  // training points
    const double* alpha;           // [np]
    const double* Kinv;            // [np*ld] lower tiles of K_y^-1 (the upper triangle is never read)
    long long ld;
    int n, np, T;
    int dq, dqp, dz;
    double sf2;
    const double* sf2_dev;  // when set, overrides sf2
    double* tile_part;  // [ntiles * (1 + dqp)]: sum W*Kc, sum Q*dx_d^2 (d < dq)
    double* zpart;      // [T * np * ZP]: slot (t, p): partial of sum_j P_pj (z_p - z_j)
};

inline size_t grad_smem_bytes(int dqp) {
    // bar | Xi Xj | ni nj | zi zj | ai aj | rowacc[4][CT][ZP] | colacc[4][CT][ZP] | wred[16][1+dqp]
    return 128 + sizeof(double) * (size_t)(2 * CT * dqp + 2 * CT + 2 * CT * ZP + 2 * CT + 4 * CT * ZP + 4 * CT * ZP +
                                           16 * (1 + dqp));
}

template <int KIND, bool HAS_Z>
__global__ void __launch_bounds__(COV_THREADS, 1) grad_tile_kernel(const GradArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int dqp = a.dqp, dz = HAS_Z ? a.dz : 0;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
    double* Xi = reinterpret_cast<double*>(smraw + 128);
    double* Xj = Xi + CT * dqp;
    double* sni = Xj + CT * dqp;
    double* snj = sni + CT;
    double* zi = snj + CT;
    double* zj = zi + CT * ZP;
    double* sai = zj + CT * ZP;
    double* saj = sai + CT;
    double* rowacc = saj + CT;              // [4][CT][ZP]
    double* colacc = rowacc + 4 * CT * ZP;  // [4][CT][ZP]
    double* wred = colacc + 4 * CT * ZP;    // [16][1+dqp]

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
        tma_load_1d(Xi, a.xs + (long long)ti * CT * dqp, xb, bar);
        tma_load_1d(Xj, a.xs + (long long)tj * CT * dqp, xb, bar);
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
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
    double acc[4][4][2];
    tile_cross_dmma(Xi, Xj, dqp, wm0, wn0, g, t, acc);

    const bool diag_tile = (ti == tj);
    const double sf2 = a.sf2_dev ? *a.sf2_dev : a.sf2;
    const double* Wp = a.Kinv + (long long)ti * CT * a.ld + (long long)tj * CT;

    double sumWK = 0.0;
    double rowz[4][ZP];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int k = 0; k < ZP; k++) rowz[mi][k] = 0.0;

#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
        double colz[2][ZP];
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int k = 0; k < ZP; k++) colz[e][k] = 0.0;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int row = wm0 + mi * 8 + g;
            const int gi = ti * CT + row;
            const double2 kin = *reinterpret_cast<const double2*>(Wp + (long long)row * a.ld + wn0 + ni * 8 + 2 * t);
            const double ai = sai[row];
            const double nri = sni[row];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = wn0 + ni * 8 + 2 * t + e;
                const int gj = tj * CT + col;
                double s = fmax(nri + snj[col] - 2.0 * acc[mi][ni][e], 0.0);
                if (diag_tile && row == col) s = 0.0;
                double dzk[ZP];
                double sz = 0.0;
#pragma unroll
                for (int k = 0; k < ZP; k++) {
                    dzk[k] = (HAS_Z && k < dz) ? (zi[row * ZP + k] - zj[col * ZP + k]) : 0.0;
                    if (HAS_Z) sz = fma(dzk[k], dzk[k], sz);
                }
                double f, fp;
                if (KIND == KERNEL_EXPSQ) {
                    f = exp_nonpos(-(s + 0.5 * sz));
                    fp = -f;
                } else {
                    kq_eval<KIND>(s, f, fp);
                    if (HAS_Z) {
                        double kz = exp_nonpos(-0.5 * sz);
                        f *= kz;
                        fp *= kz;
                    }
                }
                double W = ai * saj[col] - (e == 0 ? kin.x : kin.y);
                if (gi >= a.n || gj >= a.n) W = 0.0;
                const double WK = W * f;
                sumWK += WK;
                acc[mi][ni][e] = W * sf2 * fp;  // Q
                const double P = WK * sf2;
#pragma unroll
                for (int k = 0; k < ZP; k++) {
                    if (k < dz) {
                        rowz[mi][k] = fma(P, dzk[k], rowz[mi][k]);
                        colz[e][k] = fma(-P, dzk[k], colz[e][k]);
                    }
                }
            }
        }
        if (dz > 0 && !diag_tile) {
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int k = 0; k < ZP; k++) {
                    if (k < dz) {
                        double v = colz[e][k];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (g == 0) colacc[((warp >> 2) * CT + wn0 + ni * 8 + 2 * t + e) * ZP + k] = v;
                    }
                }
        }
    }
    if (dz > 0) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int k = 0; k < ZP; k++) {
                if (k < dz) {
                    double v = rowz[mi][k];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (t == 0) rowacc[((warp & 3) * CT + wm0 + mi * 8 + g) * ZP + k] = v;
                }
            }
    }
    // sum W*Kc over the warp
    {
        double v = sumWK;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wred[warp * (1 + dqp)] = v;
    }
    // sum Q * (xs_id - xs_jd)^2 per input dimension
    for (int d = 0; d < a.dq; d++) {
        double xi_d[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) xi_d[mi] = Xi[(wm0 + mi * 8 + g) * dqp + d];
        double sd = 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const double xj = Xj[(wn0 + ni * 8 + 2 * t + e) * dqp + d];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) {
                    const double df = xi_d[mi] - xj;
                    sd = fma(acc[mi][ni][e], df * df, sd);
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
        if (lane == 0) wred[warp * (1 + dqp) + 1 + d] = sd;
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
    const int maxb = (int)grad_smem_bytes(pad_dq(32));
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
    GPP_SET((grad_tile_kernel<KERNEL_EXPSQ, true>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN32, true>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN52, true>))
    GPP_SET((grad_tile_kernel<KERNEL_EXPSQ, false>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN32, false>))
    GPP_SET((grad_tile_kernel<KERNEL_MATERN52, false>))
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

inline cudaError_t launch_grad(const GradArgs& a, int kind, cudaStream_t st) {
    int nt = a.T * (a.T + 1) / 2;
    size_t sm = grad_smem_bytes(a.dqp);
    count_launch();
    if (a.dz > 0) {
        if (kind == KERNEL_EXPSQ) grad_tile_kernel<KERNEL_EXPSQ, true><<<nt, COV_THREADS, sm, st>>>(a);
        else if (kind == KERNEL_MATERN32) grad_tile_kernel<KERNEL_MATERN32, true><<<nt, COV_THREADS, sm, st>>>(a);
        else grad_tile_kernel<KERNEL_MATERN52, true><<<nt, COV_THREADS, sm, st>>>(a);
    } else {
        if (kind == KERNEL_EXPSQ) grad_tile_kernel<KERNEL_EXPSQ, false><<<nt, COV_THREADS, sm, st>>>(a);
        else if (kind == KERNEL_MATERN32) grad_tile_kernel<KERNEL_MATERN32, false><<<nt, COV_THREADS, sm, st>>>(a);
        else grad_tile_kernel<KERNEL_MATERN52, false><<<nt, COV_THREADS, sm, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace gpp
