// K4: finishing pass of the batched predictive mean / variance and the fused acquisition + argmax.
//
// For a chunk of candidates the engine builds K* (cov_tile_kernel, with the K* alpha partials fused),
// runs V = K* L^-T on the FP64 tensor cores with a row-sum-of-squares epilogue (dgemm_dmma.cuh,
// EPI_ROWSQ), and this kernel combines the per-column-tile partials:
//   mu_i  = m(x_i) + sum_t mean_part[t][i]                       (GPR.predict, models/gpregression.py:122-149)
//   var_i = max(sf2 - sum_t rowsq[t][i] (+ noise_i), min_var)    (exact_prediction)
// and optionally the acquisition value (bayesian_optimizations/AFs.py:102-159) with a per-CTA
// arg-max (first index wins ties, like torch.argmax in BO_GP_plus.py:192-194).
#pragma once
#include <cuda_runtime.h>

namespace gpp {

enum { ACQ_HF = 0, ACQ_LF = 1, ACQ_EI = 2 };

struct PredFinishArgs {
    const double* mean_part;  // [tiles_c][ldp]
    const double* rowsq;      // [tiles_c][ldp]
    long long ldp;
    int tiles_c;
    int m;                    // candidates in this chunk
    long long base;           // global index of the chunk's first candidate
    double sf2;
    const int* mean_idx;      // [m] or NULL
    const double* beta;
    int n_mean;
    const int* noise_idx;     // [m] or NULL
    const double* noise;
    int n_noise;
    int include_noise;
    double min_var;
    double* mean_out;         // [m] (chunk-local) or NULL
    double* var_out;
    // acquisition (enabled when cost != NULL)
    const int* cost_idx;      // [m] or NULL (all 0)
    const double* cost;       // [n_cost]
    const int* kind_by_cost;  // [n_cost]
    const double* best_f;     // [n_cost]
    int n_cost;
    int maximize;
    double si, y_min, y_std;
    double* score_out;        // [m] or NULL
    double* blk_best;         // [gridDim.x]
    long long* blk_idx;       // [gridDim.x]
};

__global__ void __launch_bounds__(256) predict_finish_kernel(const PredFinishArgs a) {
    __shared__ double sb[256];
    __shared__ long long si_[256];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double score = -INFINITY;
    long long gidx = 0x7fffffffffffffffLL;
    if (i < a.m) {
        double mu = 0.0, q = 0.0;
        for (int t = 0; t < a.tiles_c; t++) {
            mu += a.mean_part[(long long)t * a.ldp + i];
            q += a.rowsq[(long long)t * a.ldp + i];
        }
        int mi = a.mean_idx ? a.mean_idx[i] : 0;
        if (a.n_mean > 0 && mi >= 0 && mi < a.n_mean) mu += a.beta[mi];
        double var = a.sf2 - q;
        if (a.include_noise) {
            int g = a.noise_idx ? a.noise_idx[i] : 0;
            if (g >= 0 && g < a.n_noise) var += a.noise[g];
        }
        var = fmax(var, a.min_var);
        if (a.mean_out) a.mean_out[i] = mu;
        if (a.var_out) a.var_out[i] = var;
        if (a.cost) {
            int c = a.cost_idx ? a.cost_idx[i] : 0;
            if (c < 0 || c >= a.n_cost) c = 0;
            const double mean = a.y_min + a.y_std * mu;
            const double sigma = sqrt(var) * a.y_std;
            const double bf = a.best_f[c];
            const double sg = (bf > 0.0) ? 1.0 : ((bf < 0.0) ? -1.0 : 0.0);
            double u = (mean - bf - sg * a.si) / sigma;
            if (!a.maximize) u = -u;
            const int kind = a.kind_by_cost[c];
            double ei;
            if (kind == ACQ_HF) {
                ei = sigma * u;
            } else {
                const double pdf = 0.3989422804014327 * exp(-0.5 * u * u);
                if (kind == ACQ_LF) ei = sigma * pdf;
                else ei = sigma * (pdf + u * 0.5 * erfc(-u * 0.7071067811865476));
            }
            score = ei / a.cost[c];
            gidx = a.base + i;
            if (a.score_out) a.score_out[i] = score;
        }
    }
    if (!a.cost) return;
    sb[threadIdx.x] = score;
    si_[threadIdx.x] = gidx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double s2 = sb[threadIdx.x + o];
            long long i2 = si_[threadIdx.x + o];
            double s1 = sb[threadIdx.x];
            long long i1 = si_[threadIdx.x];
            if (s2 > s1 || (s2 == s1 && i2 < i1)) {
                sb[threadIdx.x] = s2;
                si_[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.blk_best[blockIdx.x] = sb[0];
        a.blk_idx[blockIdx.x] = si_[0];
    }
}

}  // namespace gpp
