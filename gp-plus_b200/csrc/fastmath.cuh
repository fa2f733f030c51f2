// exp of a non-positive argument for the covariance / gradient tile kernels.
//
// Every kernel entry K_ij = sigma_f^2 * f(s_ij) * exp(-|z_i - z_j|^2 / 2) needs one or two exponentials of a
// NON-POSITIVE argument.  CUDA's exp() is inlined per call site with its polynomial coefficients as 64-bit
// immediates, which sm_100a materialises with two UMOV / MOV instructions per coefficient (FP64 instructions cannot
// carry a 64-bit immediate): in the fully unrolled 32-pairs-per-thread tile loops that was ~22 of ~50 instructions
// per exponential, plus range-check branches that these arguments never take.  The issue slots, not the FP64 pipe,
// bound those kernels (profiles/r02_sass_histogram.txt), so this version keeps the coefficients in constant memory
// (operands of the DFMAs, no extra instructions) and has no branches:
//
//   n = rint(x / ln2) by the 1.5 * 2^52 shift, r = x - n ln2 (two-constant Cody-Waite), e^r by its degree-13 Taylor
//   polynomial (|r| <= ln2 / 2: truncation 4e-18), scaled by 2^n through the exponent field.  Arguments below -708
//   are clamped (result 3e-308 instead of a denormal / 0): K entries are O(1), the difference is 1e-308.
//
// Maximum error measured against libm on 4e6 points of [-708, 0]: < 1 ulp (tests/test_fastmath.py compiles this
// header for the host and checks it).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define GPP_FM_HD __host__ __device__ __forceinline__
#else
#define GPP_FM_HD inline
#endif

namespace gpp {

#define GPP_EXP_COEFFS                                                                                        \
    {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, \
     1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5}

#if defined(__CUDACC__)
__constant__ double c_exp_taylor_dev[12] = GPP_EXP_COEFFS;  // constant-bank operands of the DFMAs
#endif
static const double c_exp_taylor_host[12] = GPP_EXP_COEFFS;

GPP_FM_HD double exp_nonpos(double x) {
    const double L2E = 1.4426950408889634074;       // 1 / ln 2
    const double LN2_HI = 6.93147180369123816490e-01;  // ln 2 with its low 21 mantissa bits cleared: n * LN2_HI exact
    const double LN2_LO = 1.90821492927058770002e-10;
    const double SHIFT = 6755399441055744.0;        // 1.5 * 2^52
    x = fmax(x, -708.0);
    const double t = fma(x, L2E, SHIFT);
    const double nd = t - SHIFT;
    double r = fma(nd, -LN2_HI, x);
    r = fma(nd, -LN2_LO, r);
#if defined(__CUDA_ARCH__)
    const double* c = c_exp_taylor_dev;
#else
    const double* c = c_exp_taylor_host;
#endif
    double p = c[0];
#pragma unroll
    for (int k = 1; k < 12; k++) p = fma(p, r, c[k]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // multiply by 2^n: n sits in the low word of t; p is in [0.70, 1.42] and n >= -1022, so the sum stays normal
#if defined(__CUDA_ARCH__)
    const int n = __double2loint(t);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#else
    uint64_t tb, pb;
    memcpy(&tb, &t, 8);
    memcpy(&pb, &p, 8);
    const int32_t n = (int32_t)(uint32_t)(tb & 0xffffffffu);
    pb += (uint64_t)((int64_t)n << 52);
    double out;
    memcpy(&out, &pb, 8);
    return out;
#endif
}

#if defined(__CUDACC__)
// Device form used by the tile kernels: same arithmetic, the clamp at -708 done on the high word (x <= 0, so a larger
// magnitude is a larger unsigned high word; 0xC0862000 is the high word of -708.0) instead of fmax(): one ISETP + two
// SEL, nothing on the FP64 pipe.
__device__ __forceinline__ double exp_nonpos_dev(double x) {
    const double L2E = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double SHIFT = 6755399441055744.0;
    if ((unsigned)__double2hiint(x) > 0xC0862000u) x = -708.0;
    const double t = fma(x, L2E, SHIFT);
    const double nd = t - SHIFT;
    double r = fma(nd, -LN2_HI, x);
    r = fma(nd, -LN2_LO, r);
    const double* c = c_exp_taylor_dev;
    double p = c[0];
#pragma unroll
    for (int k = 1; k < 12; k++) p = fma(p, r, c[k]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int n = __double2loint(t);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
#endif

}  // namespace gpp
