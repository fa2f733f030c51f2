// Host-side O(p) part of MLLObjective.fun (optim/mll_scipy.py:112-127 of the reference) in plain C++:
// float32 rounding of theta, raw -> natural transforms, log-priors and the chain rule back to theta.
// Mirrors gpplus_b200/optim/_fast_objective.py term by term (that module compiles the layout and is itself
// validated against the torch path).
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/gpplus_b200.h"

namespace gpp {

struct PriorTerm {
    int kind = 0, off = 0, len = 0;
    std::vector<double> a, b, c;
};

struct ThetaLayout {
    bool set = false;
    int p = 0;
    int off_latent = -1, n_onehot = 0;
    std::vector<double> zeta, latent_const;
    double latent_ls = 1.0;
    int off_noise = -1;
    std::vector<double> noise_const;
    double noise_lb = 0.0;
    int off_os = -1;
    double os_const = 0.0;
    int off_ls = -1;
    std::vector<double> ls_const;
    int ls_kind = 0;
    double w_num = 0.5;
    std::vector<int> off_mean;
    std::vector<double> mean_const;
    std::vector<PriorTerm> priors;
    // scratch (natural parameters and chain-rule factors of the current evaluation)
    std::vector<double> theta, w, dw, z, noise, dnoise, beta;
    double sf2 = 0.0, dos = 0.0;
};

static inline void copy_arr(std::vector<double>& dst, const double* src, size_t n) {
    dst.assign(n, 0.0);
    if (src)
        for (size_t i = 0; i < n; i++) dst[i] = src[i];
}

// returns an error message or nullptr
inline const char* layout_copy(ThetaLayout& L, const gpp_theta_layout* in, int dq, int dz, int n_combo, int n_noise,
                               int n_mean) {
    if (!in || in->p <= 0) return "theta layout: p must be positive";
    L = ThetaLayout();
    L.p = in->p;
    auto in_range = [&](int off, int len) { return off < 0 || (off + len <= in->p); };
    if (dz > 0) {
        if (in->n_onehot <= 0 || !in->zeta) return "theta layout: latent map needs zeta and n_onehot";
        if (!in_range(in->off_latent, dz * in->n_onehot)) return "theta layout: latent block out of range";
        if (in->off_latent < 0 && !in->latent_const) return "theta layout: frozen latent map needs latent_const";
        L.off_latent = in->off_latent;
        L.n_onehot = in->n_onehot;
        copy_arr(L.zeta, in->zeta, (size_t)n_combo * in->n_onehot);
        copy_arr(L.latent_const, in->latent_const, (size_t)dz * in->n_onehot);
        L.latent_ls = in->latent_ls;
        if (!(L.latent_ls > 0.0)) return "theta layout: latent lengthscale must be positive";
    }
    if (!in_range(in->off_noise, n_noise)) return "theta layout: noise block out of range";
    if (in->off_noise < 0 && !in->noise_const) return "theta layout: frozen noise needs noise_const";
    L.off_noise = in->off_noise;
    copy_arr(L.noise_const, in->noise_const, (size_t)n_noise);
    L.noise_lb = in->noise_lb;
    if (!in_range(in->off_os, 1)) return "theta layout: outputscale out of range";
    L.off_os = in->off_os;
    L.os_const = in->os_const;
    if (dq > 0) {
        if (!in_range(in->off_ls, dq)) return "theta layout: lengthscale block out of range";
        if (in->off_ls < 0 && !in->ls_const) return "theta layout: frozen lengthscale needs ls_const";
        L.off_ls = in->off_ls;
        copy_arr(L.ls_const, in->ls_const, (size_t)dq);
        L.ls_kind = in->ls_kind;
        L.w_num = in->w_num;
        if (L.ls_kind != 0 && L.ls_kind != 1) return "theta layout: unknown ls_kind";
    }
    L.off_mean.assign(n_mean, -1);
    copy_arr(L.mean_const, in->mean_const, (size_t)n_mean);
    for (int k = 0; k < n_mean; k++) {
        int off = in->off_mean ? in->off_mean[k] : -1;
        if (!in_range(off, 1)) return "theta layout: mean constant out of range";
        if (off < 0 && !in->mean_const) return "theta layout: frozen mean needs mean_const";
        L.off_mean[k] = off;
    }
    for (int i = 0; i < in->n_priors; i++) {
        const gpp_prior& q = in->priors[i];
        PriorTerm t;
        t.kind = q.kind;
        t.off = q.off;
        t.len = q.len;
        switch (q.kind) {
            case GPP_PRIOR_NORMAL:
            case GPP_PRIOR_HORSESHOE:
                if (q.off < 0 || q.off + q.len > in->p || !q.a || !q.b) return "theta layout: bad prior block";
                copy_arr(t.a, q.a, q.len);
                copy_arr(t.b, q.b, q.len);
                break;
            case GPP_PRIOR_MOLLIFIED:
                if (q.off < 0 || q.off + q.len > in->p || !q.a || !q.b || !q.c) return "theta layout: bad prior block";
                copy_arr(t.a, q.a, q.len);
                copy_arr(t.b, q.b, q.len);
                copy_arr(t.c, q.c, q.len);
                break;
            case GPP_PRIOR_LOGNORMAL_OS:
                if (!q.a || !q.b) return "theta layout: bad prior block";
                copy_arr(t.a, q.a, 1);
                copy_arr(t.b, q.b, 1);
                break;
            case GPP_PRIOR_CONST:
                if (!q.a) return "theta layout: bad prior block";
                copy_arr(t.a, q.a, 1);
                break;
            default:
                return "theta layout: unknown prior kind";
        }
        L.priors.push_back(t);
    }
    L.theta.assign(L.p, 0.0);
    L.w.assign(dq, 0.0);
    L.dw.assign(dq, 0.0);
    L.z.assign((size_t)n_combo * dz, 0.0);
    L.noise.assign(n_noise, 0.0);
    L.dnoise.assign(n_noise, 0.0);
    L.beta.assign(n_mean, 0.0);
    L.set = true;
    return nullptr;
}

// theta -> natural parameters (into L's scratch); theta is rounded through float32 like mll_scipy.py:97
inline void layout_natural(ThetaLayout& L, const double* theta_in, int dq, int dz, int n_combo, int n_noise,
                           int n_mean) {
    for (int i = 0; i < L.p; i++) L.theta[i] = (double)(float)theta_in[i];
    const double* th = L.theta.data();
    if (dz > 0) {
        const double* A = L.off_latent >= 0 ? th + L.off_latent : L.latent_const.data();
        for (int c = 0; c < n_combo; c++)
            for (int k = 0; k < dz; k++) {
                double s = 0.0;
                for (int j = 0; j < L.n_onehot; j++) s = fma(L.zeta[(size_t)c * L.n_onehot + j], A[k * L.n_onehot + j], s);
                L.z[(size_t)c * dz + k] = s / L.latent_ls;
            }
    }
    for (int k = 0; k < n_noise; k++) {
        const double raw = L.off_noise >= 0 ? th[L.off_noise + k] : L.noise_const[k];
        const double e = exp(raw);
        L.noise[k] = L.noise_lb + e;
        L.dnoise[k] = e;
    }
    {
        const double raw = L.off_os >= 0 ? th[L.off_os] : L.os_const;
        if (raw > 20.0) {  // torch.nn.Softplus(beta=1, threshold=20)
            L.sf2 = raw;
            L.dos = 1.0;
        } else {
            const double e = exp(raw);
            L.sf2 = log1p(e);
            L.dos = e / (1.0 + e);
        }
    }
    const double ln10 = 2.302585092994046;
    for (int d = 0; d < dq; d++) {
        const double raw = L.off_ls >= 0 ? th[L.off_ls + d] : L.ls_const[d];
        double ls, dfac, dself;
        if (L.ls_kind == 1) {
            ls = 0.7071067811865476 * pow(10.0, -raw / 2.0);
            dfac = ln10;
            dself = -0.5 * ln10;
        } else {
            ls = exp(raw);
            dfac = -2.0;
            dself = 1.0;
        }
        if (L.w_num == 0.0) {
            L.w[d] = ls;
            L.dw[d] = dself * ls;
        } else {
            L.w[d] = L.w_num / (ls * ls);
            L.dw[d] = dfac * L.w[d];
        }
    }
    for (int k = 0; k < n_mean; k++) L.beta[k] = L.off_mean[k] >= 0 ? th[L.off_mean[k]] : L.mean_const[k];
}

// grad (length p) receives the chain rule of the data term; returns nothing
inline void layout_chain(const ThetaLayout& L, const gpp_mll_result& r, int dq, int dz, int n_combo, int n_noise,
                         int n_mean, double* grad) {
    for (int i = 0; i < L.p; i++) grad[i] = 0.0;
    if (dz > 0 && L.off_latent >= 0) {
        for (int k = 0; k < dz; k++)
            for (int j = 0; j < L.n_onehot; j++) {
                double s = 0.0;
                for (int c = 0; c < n_combo; c++) s = fma(r.d_z[(size_t)c * dz + k], L.zeta[(size_t)c * L.n_onehot + j], s);
                grad[L.off_latent + k * L.n_onehot + j] = s / L.latent_ls;
            }
    }
    if (L.off_noise >= 0)
        for (int k = 0; k < n_noise; k++) grad[L.off_noise + k] = r.d_noise[k] * L.dnoise[k];
    if (L.off_os >= 0) grad[L.off_os] = r.d_sigma_f2 * L.dos;
    if (L.off_ls >= 0)
        for (int d = 0; d < dq; d++) grad[L.off_ls + d] = r.d_w[d] * L.dw[d];
    for (int k = 0; k < n_mean; k++)
        if (L.off_mean[k] >= 0) grad[L.off_mean[k]] = r.d_beta[k];
}

// sum of log-priors; when grad != nullptr subtracts d(sum log p)/d theta from it (objective = nll - sum log p)
inline double layout_priors(const ThetaLayout& L, double* grad) {
    const double half_log_2pi = 0.9189385332046727;
    const double* th = L.theta.data();
    double total = 0.0;
    for (const PriorTerm& t : L.priors) {
        switch (t.kind) {
            case GPP_PRIOR_NORMAL:
                for (int i = 0; i < t.len; i++) {
                    const double zed = (th[t.off + i] - t.a[i]) / t.b[i];
                    total += -0.5 * zed * zed - log(t.b[i]) - half_log_2pi;
                    if (grad) grad[t.off + i] += zed / t.b[i];
                }
                break;
            case GPP_PRIOR_LOGNORMAL_OS: {
                const double s = L.sf2, ls = log(s), zed = (ls - t.a[0]) / t.b[0];
                total += -ls - log(t.b[0]) - half_log_2pi - 0.5 * zed * zed;
                if (grad && L.off_os >= 0) grad[L.off_os] -= (-1.0 / s - zed / (t.b[0] * s)) * L.dos;
                break;
            }
            case GPP_PRIOR_HORSESHOE:
                for (int i = 0; i < t.len; i++) {
                    const double v = th[t.off + i], ev = exp(v), tt = t.b[i] + ev, r = t.a[i] / tt;
                    const double u = 1.0 + 3.0 * r * r, lu = log(u);
                    total += log(lu) + v;
                    if (grad) grad[t.off + i] -= (6.0 * r / (u * lu)) * (-r * ev / tt) + 1.0;
                }
                break;
            case GPP_PRIOR_MOLLIFIED:
                for (int i = 0; i < t.len; i++) {
                    const double a = t.a[i], b = t.b[i], ts = t.c[i];
                    const double dev = th[t.off + i] - 0.5 * (a + b);
                    double out = fabs(dev) - 0.5 * (b - a);
                    if (out < 0.0) out = 0.0;
                    total += -0.5 * (out / ts) * (out / ts) - log(ts) - half_log_2pi -
                             log(1.0 + (b - a) / (2.5066282746310002 * ts));
                    if (grad) grad[t.off + i] += (out / (ts * ts)) * (dev > 0.0 ? 1.0 : (dev < 0.0 ? -1.0 : 0.0));
                }
                break;
            default:  // GPP_PRIOR_CONST
                total += t.a[0];
        }
    }
    return total;
}

}  // namespace gpp
