// sb_math2.cuh -- two-at-a-time versions of sb_expf / sb_logf / sb_phif for sm_100a using the packed fp32x2
// instructions FFMA2 / FADD2 / FMUL2 (PTX fma.rn.f32x2 ...; IEEE-754 RN per element). Every element goes through
// exactly the operation sequence of the scalar functions in sb_math.h, so results are bit-identical to them (and to
// the CPU oracle); only the number of issued instructions halves for the floating-point part.
#pragma once
#include "sb_math.h"

// Bit casts of the halves of a packed result go through an explicit PTX mov: with nvcc 12.9 / sm_100a a plain
// __float_as_int() on the .x half of an FADD2 result feeding integer arithmetic was observed to read a wrong register
// (tests/test_ldpc_decoder_gpu.py::test_device_phi_scalar_and_packed_equal_oracle guards this).
__device__ __forceinline__ int f2i_mov(float x) { int r; asm("mov.b32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ float i2f_mov(int x) { float r; asm("mov.b32 %0, %1;" : "=f"(r) : "r"(x)); return r; }
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }

// e^x for x in [-87.3, 88.7] (callers guarantee the range: no special cases)
__device__ __forceinline__ float2 sb_expf2_inrange(float2 x) {
    float2 t = __ffma2_rn(x, f2s(1.44269504088896341f), f2s(12582912.0f));
    float2 nf = __fadd2_rn(t, f2s(-12582912.0f));
    int n0 = f2i_mov(t.x) - 0x4B400000, n1 = f2i_mov(t.y) - 0x4B400000;
    float2 r = __ffma2_rn(nf, f2s(-0.693145751953125f), x);
    r = __ffma2_rn(nf, f2s(-1.42860677e-06f), r);
    float2 g = __ffma2_rn(f2s(0x1.a124e4p-13f), r, f2s(0x1.6d4316p-10f));
    g = __ffma2_rn(g, r, f2s(0x1.1110e0p-7f));
    g = __ffma2_rn(g, r, f2s(0x1.5554eap-5f));
    g = __ffma2_rn(g, r, f2s(0x1.555556p-3f));
    g = __ffma2_rn(g, r, f2s(0.5f));
    float2 r2 = __fmul2_rn(r, r);
    float2 s = __ffma2_rn(r2, g, r);
    float2 p = __fadd2_rn(f2s(1.0f), s);
    return f2(i2f_mov(f2i_mov(p.x) + (n0 << 23)), i2f_mov(f2i_mov(p.y) + (n1 << 23)));
}

// log(y) for positive normal y
__device__ __forceinline__ float2 sb_logf2(float2 y) {
    int ix0 = f2i_mov(y.x), ix1 = f2i_mov(y.y);
    int e0 = (ix0 - 0x3f3504f3) >> 23, e1 = (ix1 - 0x3f3504f3) >> 23;
    float2 m = f2(i2f_mov(ix0 - (e0 << 23)), i2f_mov(ix1 - (e1 << 23)));
    float2 ef = f2((float)e0, (float)e1);
    float2 r = __fadd2_rn(m, f2s(-1.0f));
    float2 p = __ffma2_rn(f2s(0x1.1d8ea6p-4f), r, f2s(-0x1.d635bcp-4f));
    p = __ffma2_rn(p, r, f2s(0x1.dea282p-4f));
    p = __ffma2_rn(p, r, f2s(-0x1.fcf4c6p-4f));
    p = __ffma2_rn(p, r, f2s(0x1.23d21ap-3f));
    p = __ffma2_rn(p, r, f2s(-0x1.555b4ap-3f));
    p = __ffma2_rn(p, r, f2s(0x1.999d5ap-3f));
    p = __ffma2_rn(p, r, f2s(-0x1.fffffcp-3f));
    p = __ffma2_rn(p, r, f2s(0x1.555554p-2f));
    float2 r2 = __fmul2_rn(r, r);
    float2 r3 = __fmul2_rn(r2, r);
    float2 nh = __fmul2_rn(f2s(-0.5f), r2);                  // == -(0.5 * r2) exactly
    float2 tl = __ffma2_rn(r3, p, nh);
    float2 lo = __ffma2_rn(ef, f2s(1.42860677e-06f), tl);
    float2 t2 = __fadd2_rn(r, lo);
    return __ffma2_rn(ef, f2s(0.693145751953125f), t2);
}

// Shared-memory copy of the sb_logf_tab table (sb_math.h), split into an inv_c array and a log c array so that the
// packed code loads each operand straight into its register pair. Entry i of this lane's copy sits at offset
// ((bits(y) >> rs) & mask) | lane_off of either array: the host picks the replication factor -- 32 copies (one per bank,
// conflict-free LDS: rs 10, mask 0x1f80, lane_off 4*lane) when the table fits next to the messages, else a single copy
// (rs 15, mask 0xfc, lane_off 0).
template <int REP>                                        // REP = 32, 8 or 1 copies: compile-time shift, mask, distance
struct LogTab {
    uint32_t inv, lane_off;                               // `inv` is CTA-uniform; the log c array follows the inv_c array
    static constexpr int log_stride = REP == 32 ? 7 : (REP == 8 ? 5 : 2);   // log2(REP * 4 bytes)
    static constexpr int rs = SB_LOGTAB_SHIFT - log_stride;
    static constexpr int mask = (SB_LOGTAB_N - 1) << log_stride;
    static constexpr int dist = SB_LOGTAB_N * REP * 4;
};

// (a & MASK) | c in one LOP3 (c in a register; nvcc otherwise emits two LOP3 when both constants are immediates)
template <int MASK>
__device__ __forceinline__ int and_or(int a, int c) {
    int d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "n"(MASK), "r"(c));
    return d;
}

__device__ __forceinline__ float lds_ro(uint32_t a) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));     // read-only after the prologue barrier
    return v;
}

// sb_logf_tab, two at a time (same operation sequence per element)
template <class LT>
__device__ __forceinline__ float2 sb_logf2_tab(float2 y, const LT& lt) {
    int ix0 = f2i_mov(y.x), ix1 = f2i_mov(y.y);
    uint32_t o0 = lt.inv + (((ix0 >> LT::rs) & LT::mask) | lt.lane_off);
    uint32_t o1 = lt.inv + (((ix1 >> LT::rs) & LT::mask) | lt.lane_off);
    float2 inv_c = f2(lds_ro(o0), lds_ro(o1));
    float2 logc = f2(lds_ro(o0 + LT::dist), lds_ro(o1 + LT::dist));
    float2 m = f2(i2f_mov(and_or<0x007fffff>(ix0, 0x3f800000)), i2f_mov(and_or<0x007fffff>(ix1, 0x3f800000)));
    float2 F = f2(i2f_mov((int)((unsigned)ix0 >> 23) | 0x4B400000), i2f_mov((int)((unsigned)ix1 >> 23) | 0x4B400000));
    float2 ef = __fadd2_rn(F, f2s(-12583039.0f));
    float2 r = __ffma2_rn(m, inv_c, f2s(-1.0f));
    float2 q = __ffma2_rn(r, f2s(-0.25f), f2s(0x1.555556p-2f));
    q = __ffma2_rn(q, r, f2s(-0.5f));
    float2 r2 = __fmul2_rn(r, r);
    float2 s = __ffma2_rn(r2, q, r);
    float2 lo = __ffma2_rn(ef, f2s(1.42860677e-06f), s);
    float2 t2 = __fadd2_rn(logc, lo);
    return __ffma2_rn(ef, f2s(0.693145751953125f), t2);
}

template <class LT>
__device__ __forceinline__ float sb_logf_tab_s(float y, const LT& lt) {
    int ix = __float_as_int(y);
    uint32_t o = lt.inv + (((ix >> LT::rs) & LT::mask) | lt.lane_off);
    return sb_logf_tab_core(ix, lds_ro(o), lds_ro(o + LT::dist));
}

// phi(x) = log(e^x + 1) - log(e^x - 1) with the reference's fp32 clipping (see sb_phif)
template <class LT>
__device__ __forceinline__ float2 sb_phif2(float2 x, const LT& lt) {
    x.x = fminf(fmaxf(x.x, 8.5e-8f), 16.635532f);
    x.y = fminf(fmaxf(x.y, 8.5e-8f), 16.635532f);
    float2 t = sb_expf2_inrange(x);
    float2 la = sb_logf2_tab(__fadd2_rn(t, f2s(1.0f)), lt);
    float2 lb = sb_logf2_tab(__fadd2_rn(t, f2s(-1.0f)), lt);
    return __ffma2_rn(lb, f2s(-1.0f), la);                   // la - lb, one rounding
}

template <class LT>
__device__ __forceinline__ float sb_phif_s(float x, const LT& lt) {
    x = fminf(fmaxf(x, 8.5e-8f), 16.635532f);
    float t = sb_expf(x);
    return __fsub_rn(sb_logf_tab_s(__fadd_rn(t, 1.0f), lt), sb_logf_tab_s(__fadd_rn(t, -1.0f), lt));
}
