// sb_math.h -- deterministic fp32 elementary functions, bit-identical on host (gcc,
// -ffp-contract=off -mfma) and device (nvcc sm_100a), built from IEEE-754 RN add/mul/fma,
// integer ops and bit casts only (no MUFU approximations, no libm, no FTZ).
//
// Why: the reference's boxplus-phi rule evaluates phi(x) = log(e^x+1) - log(e^x-1) in fp32
// (/root/reference/src/sionna/phy/fec/ldpc/decoding.py:1110-1120). That expression is
// cancellation-limited, so a +-1 ulp difference between two libms changes small phi values by
// ~1 %. Using ONE arithmetic definition on both sides of every parity test makes the CUDA
// path bit-comparable with the CPU oracle (oracle "kernel-math" mode), while the oracle's
// "libm" mode (glibc expf/logf) measures how far any <=1 ulp libm -- such as the
// TensorFlow/Eigen one the reference runs on -- sits from it.
//
// Accuracy (tools/check_math.c against float64: exhaustive over all floats when run by hand, a stride of 997 in
// tests/test_sb_math.py):
//   sb_expf : <= 0.99 ulp on [-87.3, 88.7]          (exhaustive: 0.9876)
//   sb_logf : <= 0.87 ulp on all positive normal floats  (exhaustive: 0.8623; sb_tanhf 1.509, sb_atanhf 2.768, sb_logf_tab 0.977)
// Both are exact at the two points the reference's phi clipping constants rely on:
//   sb_expf(8.5e-8f) == 1+2^-23   and   sb_logf(2^24) == sb_logf(2^24-1).
#pragma once
#include <stdint.h>
#include "sb_logtab.h"

#if defined(__CUDA_ARCH__)
#define SB_HD __host__ __device__ __forceinline__
#define SB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define SB_MUL(a, b) __fmul_rn((a), (b))
#define SB_ADD(a, b) __fadd_rn((a), (b))
#define SB_SUB(a, b) __fsub_rn((a), (b))
#define SB_DIV(a, b) __fdiv_rn((a), (b))
#define SB_F2I(x) __float_as_int(x)
#define SB_I2F(x) __int_as_float(x)
#else
#include <math.h>
#include <string.h>
#if defined(__CUDACC__)
#define SB_HD __host__ __device__ inline
#else
#define SB_HD static inline
#endif
#define SB_FMA(a, b, c) fmaf((a), (b), (c))
// Host builds must use -ffp-contract=off so these stay separate roundings.
#define SB_MUL(a, b) ((a) * (b))
#define SB_ADD(a, b) ((a) + (b))
#define SB_SUB(a, b) ((a) - (b))
#define SB_DIV(a, b) ((a) / (b))
static inline int32_t sb_f2i_(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float sb_i2f_(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
#define SB_F2I(x) sb_f2i_(x)
#define SB_I2F(x) sb_i2f_(x)
#endif

#define SB_INF_BITS 0x7f800000

// e^x.  x < -87.3 -> 0 (results below FLT_MIN are flushed), x > 88.7 -> +inf.
SB_HD float sb_expf(float x) {
    if (x < -87.3f) return 0.0f;
    if (x > 88.7f) return SB_I2F(SB_INF_BITS);
    // n = rint(x*log2(e)) by the 1.5*2^23 magic constant; r = x - n*ln2 (Cody-Waite, 2 parts)
    float t = SB_FMA(x, 1.44269504088896341f, 12582912.0f);
    float nf = SB_SUB(t, 12582912.0f);
    int32_t n = SB_F2I(t) - 0x4B400000;
    float r = SB_FMA(nf, -0.693145751953125f, x);          // ln2_hi: 15 significant bits
    r = SB_FMA(nf, -1.42860677e-06f, r);                   // ln2_lo
    // e^r = 1 + r + r^2 G(r), |r| <= ln2/2   (tools/gen_poly.py, "exp G deg 5")
    float g = 0x1.a124e4p-13f;
    g = SB_FMA(g, r, 0x1.6d4316p-10f);
    g = SB_FMA(g, r, 0x1.1110e0p-7f);
    g = SB_FMA(g, r, 0x1.5554eap-5f);
    g = SB_FMA(g, r, 0x1.555556p-3f);
    g = SB_FMA(g, r, 0.5f);
    float r2 = SB_MUL(r, r);
    float s = SB_FMA(r2, g, r);
    float p = SB_ADD(1.0f, s);
    return SB_I2F(SB_F2I(p) + (n << 23));                  // exact scaling by 2^n
}

// log(y) for positive normal y (y <= 0, denormals, inf, nan are outside the contract:
// callers clip first, as the reference does).
SB_HD float sb_logf(float y) {
    int32_t ix = SB_F2I(y);
    int32_t e = (ix - 0x3f3504f3) >> 23;                   // m = y*2^-e in [sqrt(.5), sqrt(2))
    float m = SB_I2F(ix - (e << 23));
    float r = SB_SUB(m, 1.0f);                             // exact
    float ef = (float)e;
    // log1p(r) = r - r^2/2 + r^3 P(r)   (tools/gen_poly.py, "log P deg 8")
    float p = 0x1.1d8ea6p-4f;
    p = SB_FMA(p, r, -0x1.d635bcp-4f);
    p = SB_FMA(p, r, 0x1.dea282p-4f);
    p = SB_FMA(p, r, -0x1.fcf4c6p-4f);
    p = SB_FMA(p, r, 0x1.23d21ap-3f);
    p = SB_FMA(p, r, -0x1.555b4ap-3f);
    p = SB_FMA(p, r, 0x1.999d5ap-3f);
    p = SB_FMA(p, r, -0x1.fffffcp-3f);
    p = SB_FMA(p, r, 0x1.555554p-2f);
    float r2 = SB_MUL(r, r);
    float r3 = SB_MUL(r2, r);
    float h = SB_MUL(0.5f, r2);
    float tl = SB_FMA(r3, p, -h);                          // r^3 P - r^2/2
    float lo = SB_FMA(ef, 1.42860677e-06f, tl);            // + e*ln2_lo
    float t2 = SB_ADD(r, lo);                              // small part first ...
    return SB_FMA(ef, 0.693145751953125f, t2);             // ... one rounding for e*ln2_hi + t2
}

// Table-driven log(y) for positive normal y, used by phi only (two logs per phi, two phi per edge per iteration: the
// hot spot of the boxplus-phi decoder). y = 2^(E-127) m, m in [1, 2) straight from the bit fields; the 64-entry table
// (sb_logtab.h, tools/gen_logtab.py) gives {1/c, log c} for the piece of m, r = m/c - 1 (one FMA, |r| <= 1/64) and
// log1p(r) = r - r^2/2 + r^3/3 - r^4/4 (truncation < 2e-10). E-127 is obtained as a float without a conversion:
// float(0x4B400000 | E) = 12582912 + E exactly. 9 floating-point and 5 integer operations (sb_logf: 16 + 5 + I2F).
// Accuracy (tools/check_math.c, exhaustive): <= 1 ulp for y >= 1; for y < 1 the absolute error stays <= 6e-8 (the
// result is not relatively accurate just below 1, which phi does not need: log(t-1) ~ 0 is subtracted from log(t+1) ~ 1).
#if defined(__CUDACC__)
static __device__ const float sb_logtab_dev[2 * SB_LOGTAB_N] = {SB_LOGTAB_VALUES};
#endif
static const float sb_logtab_host[2 * SB_LOGTAB_N] = {SB_LOGTAB_VALUES};
#if defined(__CUDA_ARCH__)
#define SB_LOGTAB_DEFAULT sb_logtab_dev
#else
#define SB_LOGTAB_DEFAULT sb_logtab_host
#endif

SB_HD float sb_logf_tab_core(int32_t ix, float inv_c, float logc) {
    float m = SB_I2F((ix & 0x007fffff) | 0x3f800000);
    float ef = SB_ADD(SB_I2F((int32_t)((uint32_t)ix >> 23) | 0x4B400000), -12583039.0f);   // E - 127, exact
    float r = SB_FMA(m, inv_c, -1.0f);
    float q = SB_FMA(r, -0.25f, 0x1.555556p-2f);
    q = SB_FMA(q, r, -0.5f);
    float r2 = SB_MUL(r, r);
    float s = SB_FMA(r2, q, r);
    float lo = SB_FMA(ef, 1.42860677e-06f, s);
    float t2 = SB_ADD(logc, lo);
    return SB_FMA(ef, 0.693145751953125f, t2);
}

SB_HD float sb_logf_tab(float y) {
    int32_t ix = SB_F2I(y);
    int32_t i = (ix >> SB_LOGTAB_SHIFT) & (SB_LOGTAB_N - 1);
    const float* tab = SB_LOGTAB_DEFAULT;
    return sb_logf_tab_core(ix, tab[2 * i], tab[2 * i + 1]);
}

// phi(x) = log(e^x + 1) - log(e^x - 1) with the reference's fp32 clipping constants
// (/root/reference/src/sionna/phy/fec/ldpc/decoding.py:1110-1120).
SB_HD float sb_phif(float x) {
    x = x < 8.5e-8f ? 8.5e-8f : x;
    x = x > 16.635532f ? 16.635532f : x;
    float t = sb_expf(x);
    return SB_SUB(sb_logf_tab(SB_ADD(t, 1.0f)), sb_logf_tab(SB_SUB(t, 1.0f)));
}

// tanh(z): odd; |z| < 2^-12 -> z (exact to fp32), else (1-q)/(1+q), q = e^{-2|z|}
// using expm1-free form with an IEEE division. Used by the "boxplus" rule
// (/root/reference/src/sionna/phy/fec/ldpc/decoding.py:1003-1005).
SB_HD float sb_tanhf(float z) {
    float a = z < 0.0f ? -z : z;
    float res;
    if (a < 0.55f) {
        // tanh(a) = a + a^3 T(a^2), |a| < 0.55
        float a2 = SB_MUL(a, a);
        float q = 0x1.4b239ap-9f;                          // tools/gen_poly.py "tanh T(a^2) deg 5"
        q = SB_FMA(q, a2, -0x1.176410p-7f);
        q = SB_FMA(q, a2, 0x1.657936p-6f);
        q = SB_FMA(q, a2, -0x1.ba142ep-5f);
        q = SB_FMA(q, a2, 0x1.111104p-3f);
        q = SB_FMA(q, a2, -0x1.555556p-2f);
        float a3 = SB_MUL(a2, a);
        res = SB_FMA(a3, q, a);
    } else if (a > 9.02f) {
        res = 1.0f;
    } else {
        float q = sb_expf(SB_MUL(-2.0f, a));
        res = SB_DIV(SB_SUB(1.0f, q), SB_ADD(1.0f, q));
    }
    return z < 0.0f ? -res : res;
}

// atanh(y) for |y| < 1: 0.5*log((1+y)/(1-y)); for |y| < 0.25 the odd series form
// y + y^3 A(y^2) keeps relative accuracy near 0
// (/root/reference/src/sionna/phy/fec/ldpc/decoding.py:1036).
SB_HD float sb_atanhf(float y) {
    float a = y < 0.0f ? -y : y;
    float res;
    if (a < 0.25f) {
        float a2 = SB_MUL(a, a);
        float q = 0x1.3b13b2p-4f;                          // 1/13
        q = SB_FMA(q, a2, 0x1.745d18p-4f);                 // 1/11
        q = SB_FMA(q, a2, 0x1.c71c72p-4f);                 // 1/9
        q = SB_FMA(q, a2, 0x1.249250p-3f);                 // 1/7
        q = SB_FMA(q, a2, 0x1.99999ap-3f);                 // 1/5
        q = SB_FMA(q, a2, 0x1.555556p-2f);                 // 1/3
        float a3 = SB_MUL(a2, a);
        res = SB_FMA(a3, q, a);
    } else {
        float ratio = SB_DIV(SB_ADD(1.0f, a), SB_SUB(1.0f, a));
        res = SB_MUL(0.5f, sb_logf(ratio));
    }
    return y < 0.0f ? -res : res;
}
