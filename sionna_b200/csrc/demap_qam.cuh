// demap_qam.cuh -- LLRs of ONE received symbol for a separable (square QAM) constellation without prior, shared by the
// stand-alone demapper kernel (phy_kernels.cu, sb_demap_qam) and the fused receive front-end (frontend.cu).
// Reference: Demapper.call + SymbolLogits2LLRs.call, /root/reference/src/sionna/phy/mapping.py:664-691, 927-967.
// Even label bits select the real PAM level, odd label bits the imaginary one (mapping.py:104-117): the 2-D sums factor,
// exp(e_j) = exp(e_re) exp(e_im), and the factor of the other dimension cancels in the LLR, so bit i only needs the
// 2^(M/2) exponents of its own dimension:
//   LLR_(2u+d) = logsumexp_{t: bit u of t = 1} e_d(t) - logsumexp_{t: bit u = 0} e_d(t),  e_d(t) = -(y_d - a_d(t))^2 * (1/no)
// (max instead of logsumexp for maxlog).
#pragma once
#include "sb_math.h"
#include "sb_math2.cuh"

// METHOD 0 = app, 1 = maxlog; H = bits per dimension; lr / li: the 2^H real / imaginary levels (registers);
// inv_n0 = 1 / max(no, tiny); out[2 * u + d] for dimension d (0 = re, 1 = im) and bit u of that dimension.
template <int METHOD, int H>
__device__ __forceinline__ void demap_qam_symbol(float2 yy, float inv_n0, const float* lr, const float* li, int hard_out,
                                                 float* out) {
    constexpr int L = 1 << H;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const float yd = d ? yy.y : yy.x;
        float e[L];
#pragma unroll
        for (int t = 0; t < L; ++t) {
            float dd = __fsub_rn(yd, d ? li[t] : lr[t]);
            e[t] = __fmul_rn(-__fmul_rn(dd, dd), inv_n0);
        }
        // "app": exponentials relative to the LARGEST exponent of the dimension, evaluated once (L exps) and shared by the
        // H bits of the dimension: LLR = log(sum_{bit=1} w_t) - log(sum_{bit=0} w_t), the maximum cancels. The reference's
        // per-group logsumexp (mapping.py:915-918) is the same number up to rounding; it is kept as the fallback for a
        // group whose members all underflow (e_t - max < -87.3), where the shared form would return log(0).
        float wgt[L];
        if (METHOD == 0 && H > 1) {
            float mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < L; ++t) mx = fmaxf(mx, e[t]);
            mx = (mx > -INFINITY && mx < INFINITY) ? mx : 0.f;
#pragma unroll
            for (int t = 0; t < L; t += 2) {
                float a0 = __fsub_rn(e[t], mx), a1 = __fsub_rn(e[t + 1], mx);
                float2 r = sb_expf2_inrange(make_float2(fmaxf(a0, -87.3f), fmaxf(a1, -87.3f)));
                wgt[t] = a0 < -87.3f ? 0.f : r.x;
                wgt[t + 1] = a1 < -87.3f ? 0.f : r.y;
            }
        }
#pragma unroll
        for (int u = 0; u < H; ++u) {
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int t = 0; t < L; ++t) {
                if ((t >> (H - 1 - u)) & 1) mx1 = fmaxf(mx1, e[t]);
                else mx0 = fmaxf(mx0, e[t]);
            }
            float l;
            // H == 1 (QPSK / BPSK per dimension): each group has ONE member, logsumexp of one value is the value itself
            // (sb_expf(0) == 1 and sb_logf(1) == 0 exactly), so "app" equals "maxlog" bit for bit without exp / log
            if (METHOD == 1 || H == 1) {
                l = __fsub_rn(mx1, mx0);
            } else {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int t = 0; t < L; ++t) {                       // ascending t inside each group
                    if ((t >> (H - 1 - u)) & 1) s1 = __fadd_rn(s1, wgt[t]);
                    else s0 = __fadd_rn(s0, wgt[t]);
                }
                if (s0 > 0.f && s1 > 0.f) {
                    const float2 lg = sb_logf2(make_float2(s0, s1));
                    l = __fsub_rn(lg.y, lg.x);
                } else {                                            // a whole group underflowed: per-group maxima
                    mx0 = (mx0 > -INFINITY && mx0 < INFINITY) ? mx0 : 0.f;
                    mx1 = (mx1 > -INFINITY && mx1 < INFINITY) ? mx1 : 0.f;
                    s0 = 0.f; s1 = 0.f;
                    for (int t = 0; t < L; ++t) {
                        if ((t >> (H - 1 - u)) & 1) { float a = __fsub_rn(e[t], mx1); s1 = __fadd_rn(s1, a < -87.3f ? 0.f : sb_expf(a)); }
                        else { float a = __fsub_rn(e[t], mx0); s0 = __fadd_rn(s0, a < -87.3f ? 0.f : sb_expf(a)); }
                    }
                    float b1 = __fadd_rn(s1 > 0.f ? sb_logf(s1) : -INFINITY, mx1);
                    float b0 = __fadd_rn(s0 > 0.f ? sb_logf(s0) : -INFINITY, mx0);
                    l = __fsub_rn(b1, b0);
                }
            }
            out[2 * u + d] = hard_out ? (l > 0.f ? 1.f : 0.f) : l;
        }
    }
}
