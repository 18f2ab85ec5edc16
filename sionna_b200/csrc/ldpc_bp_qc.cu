// ldpc_bp_qc.cu -- belief-propagation fast path for quasi-cyclic (5G NR) decoding graphs on sm_100a.
//
// Same algorithm, arithmetic and results as ldpc_bp.cu (one CTA per codeword, messages resident in shared memory,
// in-place flooding; reference: /root/reference/src/sionna/phy/fec/ldpc/decoding.py:416-637, 681-1166) but the
// quasi-cyclic structure of the lifted base graph (encoding.py:322-352: every base entry (r, c, s) is a ZxZ identity
// shifted by s) replaces every index table by arithmetic:
//   * message slot of base entry `be` and check offset i (CN = r*Z + i):   be*Z + i
//     - CN (r, i) walks its edges with a constant stride of Z words: no loads of indices at all;
//     - VN (c, j) reaches the edge of entry (r, c, s) at be*Z + ((j - s) mod Z): one broadcast LDS of a packed
//       table word + 4 integer ops; 32 consecutive VNs hit 32 consecutive words (mod the wrap): conflict free.
//   * a warp owns 32 consecutive checks (or variables) of ONE base row (column): degree and table entries are
//     warp-uniform, so the min-sum update keeps the whole row in registers (fully unrolled degree buckets, one
//     shared-memory read and one write per edge) and the VN update keeps addresses + messages in registers.
//   * rows / columns are processed in order of decreasing degree, dealt cyclically to the warps (load balance).
// Partial trailing blocks (pruned graphs whose size is not a multiple of Z) are handled with per-entry limits.
// Summation orders (ascending VN inside a CN, ascending CN inside a VN) are those of ldpc_bp.cu, so both kernels
// and the CPU oracle (math_mode 1, order "kernel") agree bit for bit.
#include <algorithm>
#include <numeric>
#include <vector>
#include "sb_common.h"
#include "sb_math.h"
#include "sb_math2.cuh"
#include "ldpc_graph.h"

namespace {

constexpr int kRowClasses = 5;    // CN degree classes, heaviest first: >20 (loop), <=20, <=12, <=8, <=4
constexpr int kColClasses = 11;   // VN classes, see col_class() on the host side

struct QcParams {
    int Z, n_rows, n_cols, nnz, N, E, E_alloc, n_in, n_out;
    int row_cls_end[kRowClasses];   // processing order: rows of class k are [row_cls_end[k-1], row_cls_end[k])
    int col_cls_end[kColClasses];
    int row_cls_mod[kRowClasses];   // (first index of class k) mod G, G = warp groups of this launch
    int col_cls_mod[kColClasses];
    const int4* row_info;    // {first base entry, deg, zrow, fused VN base (c*Z | s << 16... see host) or -1}
    const int4* col_info;    // {first col-edge, deg, zcol, c*Z}
    const int2* col_edge;    // {be*Z*4, s*4 | (zrow*4) << 16}, ascending base row inside a column
    const int* in_idx;       // [N] natural VN order
    const int* out_pos;      // [N]
    const int* slot_of_edge; // [E] reference edge -> slot
    const float* llr;
    float* out;
    float* state_out;
    long long B;
    int num_iter, hard_out, use_tma;
    int early;               // opt-in early termination by the syndrome of the hard decisions (see the kernel)
    int* iters_out;          // [B] iterations actually run per codeword, or nullptr
    const int2* row_edge;    // [nnz] per base entry (processing order): {column * Z, shift}  (syndrome pass only)
    int tab_rep;             // copies of the phi log table in shared memory (32, 8 or 1; 0: rule does not use it)
    float offset, llr_max;
};

// Message invariant of this kernel: a v2c message is never -0.0f. The initial v2c is canonicalised (llr + 0.0f) and
// the VN update clip(x_tot - c2v) cannot yield -0.0f because x_tot = (0 + sum c2v) + llr is never -0.0f. Hence in the
// CN phase  sign bit set <=> v2c < 0, which is the reference's sign() with sign(0) := +1 (decoding.py:800-804, :1129).
// (c2v messages may be -0.0f; the VN arithmetic does not depend on the sign of a zero.)

// ---- check-node updates on the edges pm[0], pm[Z], pm[2Z], ... of one check ------------------------------------
// boxplus-phi (decoding.py:1126-1166), two edges per step on the packed fp32x2 pipe (sb_math2.cuh).
// Exact strength reductions (outputs are bit-identical, only instructions are saved), all decided warp-uniformly:
//   (1) |x| >= 16.635532 (the phi clipping bound, :1113)  =>  phi(|x|) == 0 exactly              [saturated inputs]
//   (2) p_e == 0  =>  P - p_e == P exactly  =>  phi(P - p_e) == phi(P), evaluated once per check
//   (3) P - p_e <= 8.5e-8 (lower clipping bound)  =>  phi(P - p_e) == phi(8.5e-8) == phi_max
// Once a codeword has converged every VN->CN message except those of degree-1 VNs sits at +-llr_max >= 16.64, and a
// check costs ~3 phi evaluations instead of 2*deg; the per-iteration cost therefore depends on the channel SNR.
#ifndef SB_PHI_UNROLL
// edge pairs per trip of the phi loops. A/B on B200 (ms per 4096 codewords at 2 dB / 0 dB): 1 -> 12.33 / 15.15,
// 2 -> 12.41 / 14.66, 4 -> 14.16 / 14.51 (built with -DSB_PHI_UNROLL=n, selected with SIONNA_B200_LIB)
#define SB_PHI_UNROLL 2
#endif
#define SB_PRAGMA_(x) _Pragma(#x)
#define SB_UNROLL(n) SB_PRAGMA_(unroll n)
#define SB_PHI_HI 16.635532f
#define SB_PHI_LO 8.5e-8f
// SC = false: plain evaluation; one vote per check on its first edge pair probes for saturation and raises *sat_flag,
// which makes the CTA use the SC = true variant (votes on every pair) from the next iteration on.
// Out of line on purpose: the five degree classes then share ONE copy of each variant's loops (the kernel is bound by
// instruction fetch as much as by issue: 12.35 -> 11.85 ms per 4096 codewords at 2 dB; making phi itself a call costs more
// than it saves, 12.8 ms).
template <bool SC, class LT>
__device__ __noinline__ void cn_phi_qc(float* pm, int Z, int deg, float clip, float phi_max, int* sat_flag,
                                       const LT& lt) {
    const unsigned am = __activemask();                   // lanes of this warp working on the same block row
    float P = 0.f;
    unsigned par = 0;
    int l = 0;
SB_UNROLL(SB_PHI_UNROLL)
    for (; l + 1 < deg; l += 2) {
        float* q0 = pm + l * Z;
        float* q1 = q0 + Z;
        unsigned b0 = __float_as_uint(*q0), b1 = __float_as_uint(*q1);
        par ^= b0 ^ b1;
        float a0 = __uint_as_float(b0 & 0x7fffffffu), a1 = __uint_as_float(b1 & 0x7fffffffu);
        float2 p = make_float2(0.f, 0.f);
        if (SC) {
            if (!__all_sync(am, a0 >= SB_PHI_HI && a1 >= SB_PHI_HI)) p = sb_phif2(make_float2(a0, a1), lt);   // (1)
        } else {
            p = sb_phif2(make_float2(a0, a1), lt);
            if (l == 0 && __all_sync(am, a0 >= SB_PHI_HI && a1 >= SB_PHI_HI)) *sat_flag = 1;   // probe (benign race)
        }
        P = __fadd_rn(P, p.x);                            // :1150 sequential sum, ascending VN
        P = __fadd_rn(P, p.y);
        *q0 = __uint_as_float(__float_as_uint(p.x) | (b0 & 0x80000000u));   // phi >= 0: sign bit carries sign(x)
        *q1 = __uint_as_float(__float_as_uint(p.y) | (b1 & 0x80000000u));
    }
    if (l < deg) {
        float* q0 = pm + l * Z;
        unsigned b0 = __float_as_uint(*q0);
        par ^= b0;
        float a0 = __uint_as_float(b0 & 0x7fffffffu);
        float p = 0.f;
        if (!SC || !__all_sync(am, a0 >= SB_PHI_HI)) p = sb_phif_s(a0, lt);
        P = __fadd_rn(P, p);
        *q0 = __uint_as_float(__float_as_uint(p) | (b0 & 0x80000000u));
    }
    par &= 0x80000000u;
    float yP = 0.f;                                       // phi(P), evaluated lazily (2)
    bool have_yP = false;
    l = 0;
SB_UNROLL(SB_PHI_UNROLL)
    for (; l + 1 < deg; l += 2) {
        float* q0 = pm + l * Z;
        float* q1 = q0 + Z;
        unsigned b0 = __float_as_uint(*q0), b1 = __float_as_uint(*q1);
        float2 y;
        if (SC && __all_sync(am, ((b0 | b1) & 0x7fffffffu) == 0u)) {                                 // (2)
            if (!have_yP) { yP = sb_phif_s(P, lt); have_yP = true; }
            y = make_float2(yP, yP);
        } else {
            float2 m = __fadd2_rn(make_float2(__uint_as_float(b0 | 0x80000000u), __uint_as_float(b1 | 0x80000000u)),
                                  make_float2(P, P));     // (-p) + P  (:1155)
            if (SC && __all_sync(am, m.x <= SB_PHI_LO && m.y <= SB_PHI_LO)) y = make_float2(phi_max, phi_max);   // (3)
            else y = sb_phif2(m, lt);
        }
        *q0 = __uint_as_float(__float_as_uint(fminf(y.x, clip)) | ((b0 ^ par) & 0x80000000u));   // :1161-1163
        *q1 = __uint_as_float(__float_as_uint(fminf(y.y, clip)) | ((b1 ^ par) & 0x80000000u));
    }
    if (l < deg) {
        float* q0 = pm + l * Z;
        unsigned b0 = __float_as_uint(*q0);
        float y;
        if (SC && __all_sync(am, (b0 & 0x7fffffffu) == 0u)) {
            if (!have_yP) { yP = sb_phif_s(P, lt); have_yP = true; }
            y = yP;
        } else {
            float m = __fadd_rn(__uint_as_float(b0 | 0x80000000u), P);
            if (SC && __all_sync(am, m <= SB_PHI_LO)) y = phi_max;
            else y = sb_phif_s(m, lt);
        }
        *q0 = __uint_as_float(__float_as_uint(fminf(y, clip)) | ((b0 ^ par) & 0x80000000u));
    }
}

// Exact-degree variant (deg == D, D <= SB_PHI_REG_MAXD): phi(|x|) of the whole row stays in registers between the two
// passes (no STS in pass 1, no LDS / address arithmetic in pass 2), both passes fully unrolled. Same operation sequence
// per element as cn_phi_qc, hence bit-identical. A/B on B200 (ms per 4096 codewords, 2 dB / 0 dB):
//   loops only 12.35 / 14.60;  registers for deg <= 8, plain variant only 12.32 / 13.79 (the default);  deg <= 10
//   12.52 / 13.57;  deg <= 19 12.90 / 13.28;  registers also in the voting variant (deg <= 8) 13.38 / 13.79 -- skipped
//   straight-line code still has to be fetched, a skipped loop body does not, and the kernel is instruction-cache bound.
#ifndef SB_PHI_REG_MAXD
#define SB_PHI_REG_MAXD 8
#endif
template <int D, bool SC, class LT>
__device__ __forceinline__ void cn_phi_qc_reg(float* pm, int Z, float clip, float phi_max, int* sat_flag, const LT& lt) {
    const unsigned am = __activemask();
    unsigned w[D];                                        // phi(|x|) bits | sign(x)
    float P = 0.f;
    unsigned par = 0;
#pragma unroll
    for (int l = 0; l + 1 < D; l += 2) {
        unsigned b0 = __float_as_uint(pm[l * Z]), b1 = __float_as_uint(pm[(l + 1) * Z]);
        par ^= b0 ^ b1;
        float a0 = __uint_as_float(b0 & 0x7fffffffu), a1 = __uint_as_float(b1 & 0x7fffffffu);
        float2 q = make_float2(0.f, 0.f);
        if (SC) {
            if (!__all_sync(am, a0 >= SB_PHI_HI && a1 >= SB_PHI_HI)) q = sb_phif2(make_float2(a0, a1), lt);
        } else {
            q = sb_phif2(make_float2(a0, a1), lt);
            if (l == 0 && __all_sync(am, a0 >= SB_PHI_HI && a1 >= SB_PHI_HI)) *sat_flag = 1;
        }
        P = __fadd_rn(P, q.x);
        P = __fadd_rn(P, q.y);
        w[l] = __float_as_uint(q.x) | (b0 & 0x80000000u);
        w[l + 1] = __float_as_uint(q.y) | (b1 & 0x80000000u);
    }
    if (D & 1) {
        unsigned b0 = __float_as_uint(pm[(D - 1) * Z]);
        par ^= b0;
        float a0 = __uint_as_float(b0 & 0x7fffffffu);
        float q = 0.f;
        if (!SC || !__all_sync(am, a0 >= SB_PHI_HI)) q = sb_phif_s(a0, lt);
        P = __fadd_rn(P, q);
        w[D - 1] = __float_as_uint(q) | (b0 & 0x80000000u);
    }
    par &= 0x80000000u;
    float yP = 0.f;
    bool have_yP = false;
#pragma unroll
    for (int l = 0; l + 1 < D; l += 2) {
        const unsigned b0 = w[l], b1 = w[l + 1];
        float2 y;
        if (SC && __all_sync(am, ((b0 | b1) & 0x7fffffffu) == 0u)) {
            if (!have_yP) { yP = sb_phif_s(P, lt); have_yP = true; }
            y = make_float2(yP, yP);
        } else {
            float2 m = __fadd2_rn(make_float2(__uint_as_float(b0 | 0x80000000u), __uint_as_float(b1 | 0x80000000u)),
                                  make_float2(P, P));
            if (SC && __all_sync(am, m.x <= SB_PHI_LO && m.y <= SB_PHI_LO)) y = make_float2(phi_max, phi_max);
            else y = sb_phif2(m, lt);
        }
        pm[l * Z] = __uint_as_float(__float_as_uint(fminf(y.x, clip)) | ((b0 ^ par) & 0x80000000u));
        pm[(l + 1) * Z] = __uint_as_float(__float_as_uint(fminf(y.y, clip)) | ((b1 ^ par) & 0x80000000u));
    }
    if (D & 1) {
        const unsigned b0 = w[D - 1];
        float y;
        if (SC && __all_sync(am, (b0 & 0x7fffffffu) == 0u)) {
            if (!have_yP) { yP = sb_phif_s(P, lt); have_yP = true; }
            y = yP;
        } else {
            float m = __fadd_rn(__uint_as_float(b0 | 0x80000000u), P);
            if (SC && __all_sync(am, m <= SB_PHI_LO)) y = phi_max;
            else y = sb_phif_s(m, lt);
        }
        pm[(D - 1) * Z] = __uint_as_float(__float_as_uint(fminf(y, clip)) | ((b0 ^ par) & 0x80000000u));
    }
}

template <bool SC, int CLS, class LT>
__device__ __forceinline__ void cn_phi_dispatch(float* pm, int Z, int deg, float clip, float phi_max, int* sat_flag,
                                                const LT& lt) {
#ifndef SB_PHI_REG_SC
#define SB_PHI_REG_SC 0                                   // 0: register rows only in the plain (non-voting) variant
#endif
#if SB_PHI_REG_MAXD > 0
#define SB_PHI_CASE(D) if ((SB_PHI_REG_SC || !SC) && D <= SB_PHI_REG_MAXD && deg == D) { cn_phi_qc_reg<D, SC, LT>(pm, Z, clip, phi_max, sat_flag, lt); return; }
    if (CLS == 4) { SB_PHI_CASE(3) SB_PHI_CASE(4) }
    if (CLS == 3) { SB_PHI_CASE(5) SB_PHI_CASE(6) SB_PHI_CASE(7) SB_PHI_CASE(8) }
    if (CLS == 2) { SB_PHI_CASE(9) SB_PHI_CASE(10) }
    if (CLS == 1) { SB_PHI_CASE(19) }
#undef SB_PHI_CASE
#endif
    cn_phi_qc<SC, LT>(pm, Z, deg, clip, phi_max, sat_flag, lt);
}

__device__ __forceinline__ void cn_tanh_qc(float* pm, int Z, int deg, float clip) {
    const float atanh_clip = (float)(1 - 1e-7);
    float prod = 1.f;
#pragma unroll 2
    for (int l = 0; l < deg; ++l) {
        float* q = pm + l * Z;
        float t = sb_tanhf(__fmul_rn(*q, 0.5f));
        if (t == 0.f) t = 1e-12f;
        prod = __fmul_rn(prod, t);
        *q = t;
    }
#pragma unroll 2
    for (int l = 0; l < deg; ++l) {
        float* q = pm + l * Z;
        float e = __fmul_rn(__fdiv_rn(1.f, *q), prod);
        if (fabsf(e) < 1e-7f) e = 0.f;
        e = clipf(e, atanh_clip);
        *q = clipf(__fmul_rn(2.f, sb_atanhf(e)), clip);
    }
}

// (offset-)min-sum with the row in registers. Preconditions checked on the host: |message| <= llr_max and
// (deg-1)*llr_max < 99000, so the reference's 1e5 sentinel logic (decoding.py:849-887) reduces exactly to
//   unique minimum -> that edge gets fl(fl(m2 - m1) + m1), all others m1;  repeated minimum -> all edges m1.
// Offset, max(.,0) and clipping act on only two distinct magnitudes and are hoisted out of the edge loop.
template <int DMAX, bool EXACT>                           // EXACT: deg == DMAX, no per-edge guards
__device__ __forceinline__ void cn_minsum_qc(float* pm, int Z, int deg, float clip, float offset) {
    float x[DMAX];                                        // every element is assigned unconditionally (registers)
    float m1 = INFINITY, m2 = INFINITY;
    unsigned par = 0;
#pragma unroll
    for (int l = 0; l < DMAX; ++l) {
        float v = INFINITY;                               // neutral: never the minimum, sign +
        if (EXACT || l < deg) v = pm[l * Z];              // warp-uniform predicate
        x[l] = v;
        float a = fabsf(v);
        m2 = fminf(m2, fmaxf(m1, a));
        m1 = fminf(m1, a);
        par ^= __float_as_uint(v);
    }
    par &= 0x80000000u;
    float min_e = (m2 == m1) ? m1 : __fadd_rn(__fsub_rn(m2, m1), m1);
    if (deg == 1) min_e = __fadd_rn(100000.f, m1);
    const float o1 = fminf(fmaxf(__fsub_rn(m1, offset), 0.f), clip);
    const float oe = fminf(fmaxf(__fsub_rn(min_e, offset), 0.f), clip);
#pragma unroll
    for (int l = 0; l < DMAX; ++l) {
        float v = x[l];
        float mag = (fabsf(v) == m1) ? oe : o1;
        float y = __uint_as_float(__float_as_uint(mag) | ((__float_as_uint(v) ^ par) & 0x80000000u));
        if (EXACT || l < deg) pm[l * Z] = y;
    }
}

// generic-degree fallback (re-reads shared memory instead of holding the row in registers)
__device__ __forceinline__ void cn_minsum_qc_loop(float* pm, int Z, int deg, float clip, float offset) {
    float m1 = INFINITY, m2 = INFINITY;
    unsigned par = 0;
    for (int l = 0; l < deg; ++l) {
        float v = pm[l * Z];
        float a = fabsf(v);
        m2 = fminf(m2, fmaxf(m1, a));
        m1 = fminf(m1, a);
        par ^= __float_as_uint(v);
    }
    par &= 0x80000000u;
    float min_e = (m2 == m1) ? m1 : __fadd_rn(__fsub_rn(m2, m1), m1);
    if (deg == 1) min_e = __fadd_rn(100000.f, m1);
    const float o1 = fminf(fmaxf(__fsub_rn(m1, offset), 0.f), clip);
    const float oe = fminf(fmaxf(__fsub_rn(min_e, offset), 0.f), clip);
    for (int l = 0; l < deg; ++l) {
        float v = pm[l * Z];
        float mag = (fabsf(v) == m1) ? oe : o1;
        pm[l * Z] = __uint_as_float(__float_as_uint(mag) | ((__float_as_uint(v) ^ par) & 0x80000000u));
    }
}

template <int RULE, int CLS, class LT>
__device__ __forceinline__ void cn_qc(float* pm, int Z, int deg, float clip, float offset, float phi_max, bool sc,
                                      int* sat_flag, const LT& lt) {
    if (RULE == SB_CN_BOXPLUS_PHI) {
        if (sc) cn_phi_dispatch<true, CLS, LT>(pm, Z, deg, clip, phi_max, sat_flag, lt);
        else cn_phi_dispatch<false, CLS, LT>(pm, Z, deg, clip, phi_max, sat_flag, lt);
    }
    else if (RULE == SB_CN_BOXPLUS) cn_tanh_qc(pm, Z, deg, clip);
    else {
        const float off = (RULE == SB_CN_MINSUM) ? 0.f : offset;
        // exact-degree code for the degrees of the 5G base graphs, guarded buckets otherwise (deg is warp-uniform)
        if (CLS == 4) {
            if (deg == 3) cn_minsum_qc<3, true>(pm, Z, deg, clip, off);
            else if (deg == 4) cn_minsum_qc<4, true>(pm, Z, deg, clip, off);
            else cn_minsum_qc<4, false>(pm, Z, deg, clip, off);
        } else if (CLS == 3) {
            if (deg == 5) cn_minsum_qc<5, true>(pm, Z, deg, clip, off);
            else if (deg == 6) cn_minsum_qc<6, true>(pm, Z, deg, clip, off);
            else if (deg == 7) cn_minsum_qc<7, true>(pm, Z, deg, clip, off);
            else cn_minsum_qc<8, true>(pm, Z, deg, clip, off);
        } else if (CLS == 2) {
            if (deg == 9) cn_minsum_qc<9, true>(pm, Z, deg, clip, off);
            else if (deg == 10) cn_minsum_qc<10, true>(pm, Z, deg, clip, off);
            else cn_minsum_qc<12, false>(pm, Z, deg, clip, off);
        } else if (CLS == 1) {
            if (deg == 19) cn_minsum_qc<19, true>(pm, Z, deg, clip, off);
            else cn_minsum_qc<20, false>(pm, Z, deg, clip, off);
        } else cn_minsum_qc_loop(pm, Z, deg, clip, off);
    }
}

// explicit shared-window accesses with 32-bit addresses (no generic-address arithmetic in the hot loops)
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ int2 lds_i2(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

// ---- variable-node update (decoding.py:714-732) for VN (c, j) -------------------------------------------------
// msg_s / ce_s are 32-bit shared-window addresses of the message array and of the column's first table entry.
// Returns the unclipped x_tot. MODE 0: normal update; MODE 1: initialisation v2c = llr (decoding.py:571).
// Branch free: table entries beyond the column's degree are not read (predicate), their slot address points at the
// VN's own first edge and the accumulate / store are predicated off.
template <int DMAX, bool CHECK, bool KEEPM, int MODE, bool EXACT = false>   // EXACT: deg == DMAX and !CHECK: no guards
__device__ __forceinline__ float vn_qc(uint32_t msg_s, uint32_t ce_s, int deg, int j4, int Z4, float llr, float clip) {
    uint32_t addr[DMAX];
    float m[KEEPM ? DMAX : 1];
    bool on[DMAX];
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        int2 e = make_int2(0, 0);
        if (EXACT || k < deg) e = lds_i2(ce_s + 8 * k);   // warp-uniform predicate
        int t = j4 - (e.y & 0xffff);
        t += (t >> 31) & Z4;                              // (j - s) mod Z, in bytes
        bool ok = EXACT || ((k < deg) && (!CHECK || t < (int)((unsigned)e.y >> 16)));   // edge exists
        on[k] = ok;
        addr[k] = msg_s + e.x + t;
        float v = 0.f;
        if (MODE == 0) {
            if (ok) v = lds_f32(addr[k]);
            if (ok) acc = __fadd_rn(acc, v);              // :715 sequential, ascending CN
        }
        if (KEEPM) m[KEEPM ? k : 0] = v;
    }
    float x_tot = __fadd_rn(acc, llr);                    // :716
    const float init = __fadd_rn(llr, 0.f);               // canonical +0.0 for punctured bits (llr = -0.0)
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        if (MODE == 1) {
            if (on[k]) sts_f32(addr[k], init);
        } else {
            float v = 0.f;
            if (KEEPM) v = m[KEEPM ? k : 0];
            else if (on[k]) v = lds_f32(addr[k]);
            float y = clipf(__fadd_rn(-v, x_tot), clip);  // :724-729
            if (on[k]) sts_f32(addr[k], y);
        }
    }
    return x_tot;
}

template <int MODE>
__device__ __forceinline__ float vn_qc_loop(uint32_t msg_s, uint32_t ce_s, int deg, int j4, int Z4, float llr, float clip) {
    float acc = 0.f;
    if (MODE == 0)
        for (int k = 0; k < deg; ++k) {
            int2 e = lds_i2(ce_s + 8 * k);
            int t = j4 - (e.y & 0xffff);
            t += (t >> 31) & Z4;
            if (t < (int)((unsigned)e.y >> 16)) acc = __fadd_rn(acc, lds_f32(msg_s + e.x + t));
        }
    float x_tot = __fadd_rn(acc, llr);
    for (int k = 0; k < deg; ++k) {
        int2 e = lds_i2(ce_s + 8 * k);
        int t = j4 - (e.y & 0xffff);
        t += (t >> 31) & Z4;
        if (t < (int)((unsigned)e.y >> 16)) {
            uint32_t a = msg_s + e.x + t;
            sts_f32(a, (MODE == 1) ? __fadd_rn(llr, 0.f) : clipf(__fadd_rn(-lds_f32(a), x_tot), clip));
        }
    }
    return x_tot;
}

// VN classes (host: col_class()): 0 loop (deg > 32), 1 <=32, 2 <=20, 3 <=12, 4 <=8, 5 <=4, 6 <=2; with edges into a
// pruning-cut row: 7 loop, 8 <=12, 9 <=4; 10: degree-1 columns whose update is fused into the CN phase.
// EX: exact-degree variants (min-sum kernels; the phi kernels sit at the register cap and keep the guarded buckets)
template <int MODE, int CLS, bool EX>
__device__ __forceinline__ float vn_cls(uint32_t msgb, uint32_t ce, int deg, int j4, int Z4, float llr, float clip) {
    if (CLS == 1) return vn_qc<32, false, false, MODE>(msgb, ce, deg, j4, Z4, llr, clip);
    if (CLS == 2) return vn_qc<20, false, false, MODE>(msgb, ce, deg, j4, Z4, llr, clip);
    // exact-degree code (no guards) for every degree up to 12; deg is warp-uniform
    if (CLS == 3) {
        if (EX && deg == 9) return vn_qc<9, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        if (EX && deg == 10) return vn_qc<10, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        if (EX && deg == 11) return vn_qc<11, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        return vn_qc<12, false, true, MODE, EX>(msgb, ce, deg, j4, Z4, llr, clip);
    }
    if (CLS == 4) {
        if (EX && deg == 5) return vn_qc<5, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        if (EX && deg == 6) return vn_qc<6, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        if (EX && deg == 7) return vn_qc<7, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        return vn_qc<8, false, true, MODE, EX>(msgb, ce, deg, j4, Z4, llr, clip);
    }
    if (CLS == 5) {
        if (EX && deg == 3) return vn_qc<3, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        return vn_qc<4, false, true, MODE, EX>(msgb, ce, deg, j4, Z4, llr, clip);
    }
    if (CLS == 6) {
        if (EX && deg == 1) return vn_qc<1, false, true, MODE, true>(msgb, ce, deg, j4, Z4, llr, clip);
        return vn_qc<2, false, true, MODE, EX>(msgb, ce, deg, j4, Z4, llr, clip);
    }
    if (CLS == 8) return vn_qc<12, true, true, MODE>(msgb, ce, deg, j4, Z4, llr, clip);
    if (CLS == 9) return vn_qc<4, true, true, MODE>(msgb, ce, deg, j4, Z4, llr, clip);
    if (CLS == 10) return vn_qc<1, true, true, MODE>(msgb, ce, deg, j4, Z4, llr, clip);
    return vn_qc_loop<MODE>(msgb, ce, deg, j4, Z4, llr, clip);
}

struct WarpCtx {
    int G, grp, lane_i;
};

// first index >= start that is congruent to grp modulo G (rows/columns are dealt cyclically over ALL classes);
// start_mod = start mod G comes from the host
__device__ __forceinline__ int first_of(int start, int start_mod, const WarpCtx& w) {
    return start + (w.grp - start_mod + (w.grp < start_mod ? w.G : 0));
}

template <int RULE, int CLS, class LT>
__device__ __forceinline__ void cn_class(const QcParams& p, const WarpCtx& w, float* msg, const float* llr_s,
                                         const int4* s_row, int start, int end, float clip, bool fuse,
                                         float phi_max, bool sc, int* sat_flag, const LT& lt, unsigned char* hd) {
    for (int rr = first_of(start, p.row_cls_mod[CLS], w); rr < end; rr += w.G) {
        int4 ri = s_row[rr];
        if (w.lane_i < ri.z) {
            float* pm = msg + ri.x * p.Z + w.lane_i;
            cn_qc<RULE, CLS, LT>(pm, p.Z, ri.y, clip, p.offset, phi_max, sc, sat_flag, lt);
            if (fuse && ri.w >= 0) {
                // the row's last edge goes to a degree-1 VN: apply that VN's update right here (decoding.py:714-729
                // with a single incoming message) so the VN phase can skip the column
                int s = ri.w >> 16, vb = ri.w & 0xffff;     // shift, column index
                int j = w.lane_i + s;
                j -= (j >= p.Z) ? p.Z : 0;
                float* q = pm + (ri.y - 1) * p.Z;
                float c2v = *q;
                float x_tot = __fadd_rn(__fadd_rn(0.f, c2v), llr_s[vb * p.Z + j]);
                *q = clipf(__fadd_rn(-c2v, x_tot), clip);
                if (hd) hd[vb * p.Z + j] = 0.f >= x_tot ? 1 : 0;
            }
        }
    }
}

// Syndrome of the current hard decisions hd[] (one byte per VN): every warp walks its block rows like the CN phase and
// XORs the decisions of the row's variable nodes, VN of base entry (c, s) and check offset i being c * Z + (i + s) mod Z.
// Raises *unsat if any check is violated.
__device__ __forceinline__ void syndrome_pass(const QcParams& p, const WarpCtx& w, const int4* s_row,
                                              const unsigned char* hd, int* unsat) {
    for (int rr = w.grp; rr < p.n_rows; rr += w.G) {
        const int4 ri = s_row[rr];
        if (w.lane_i < ri.z) {
            unsigned par = 0;
            for (int l = 0; l < ri.y; ++l) {
                const int2 e = __ldg(p.row_edge + ri.x + l);
                int j = w.lane_i + e.y;
                j -= (j >= p.Z) ? p.Z : 0;
                par ^= hd[e.x + j];
            }
            if (par) *unsat = 1;
        }
    }
}

template <int MODE, int CLS, bool EX>
__device__ __forceinline__ void vn_class(const QcParams& p, const WarpCtx& w, uint32_t msgb, const float* llr_s,
                                         const int4* s_col, uint32_t s_ce, int start, int end, float clip,
                                         bool final_pass, long long b, unsigned char* hd) {
    for (int cc = first_of(start, p.col_cls_mod[CLS], w); cc < end; cc += w.G) {
        int4 ci = s_col[cc];
        if (w.lane_i < ci.z) {
            int v = ci.w + w.lane_i;
            float x_tot = vn_cls<MODE, CLS, EX>(msgb, s_ce + 8 * ci.x, ci.y, 4 * w.lane_i, 4 * p.Z, llr_s[v], clip);
            if (MODE == 0 && hd) hd[v] = 0.f >= x_tot ? 1 : 0;                                // hard decision (:622-624)
            if (MODE == 0 && final_pass) {
                int o = p.out_pos[v];
                if (o >= 0) {
                    x_tot = clipf(x_tot, clip);                                              // :730
                    p.out[(size_t)b * p.n_out + o] = p.hard_out ? (0.f >= x_tot ? 1.f : 0.f) // :622-626
                                                                : __fmul_rn(x_tot, -1.f);
                }
            }
        }
    }
}

template <int MODE, bool EX>
__device__ __forceinline__ void vn_all(const QcParams& p, const WarpCtx& w, uint32_t msgb, const float* llr_s,
                                       const int4* s_col, uint32_t s_ce, float clip, bool final_pass,
                                       bool with_fused, long long b, unsigned char* hd = nullptr) {
    const int* ce = p.col_cls_end;
    vn_class<MODE, 0, EX>(p, w, msgb, llr_s, s_col, s_ce, 0, ce[0], clip, final_pass, b, hd);
    vn_class<MODE, 1, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[0], ce[1], clip, final_pass, b, hd);
    vn_class<MODE, 2, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[1], ce[2], clip, final_pass, b, hd);
    vn_class<MODE, 3, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[2], ce[3], clip, final_pass, b, hd);
    vn_class<MODE, 4, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[3], ce[4], clip, final_pass, b, hd);
    vn_class<MODE, 5, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[4], ce[5], clip, final_pass, b, hd);
    vn_class<MODE, 6, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[5], ce[6], clip, final_pass, b, hd);
    vn_class<MODE, 7, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[6], ce[7], clip, final_pass, b, hd);
    vn_class<MODE, 8, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[7], ce[8], clip, final_pass, b, hd);
    vn_class<MODE, 9, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[8], ce[9], clip, final_pass, b, hd);
    if (with_fused) vn_class<MODE, 10, EX>(p, w, msgb, llr_s, s_col, s_ce, ce[9], ce[10], clip, final_pass, b, hd);
}

// Threads per CTA. 24 warps (80 registers/thread) for every rule: 30 warps at 64 registers were measured for the min-sum
// kernels and lost 7 % (5.32 vs 4.97 ms / 4096 codewords: more barrier and spill time than latency hiding gained).
#ifndef SB_QC_PHI_THREADS
#define SB_QC_PHI_THREADS 768                             // A/B hook: -DSB_QC_PHI_THREADS=1024 builds the 64-register variant
#endif
__host__ __device__ constexpr int qc_max_threads(int rule) { return rule == SB_CN_BOXPLUS_PHI ? SB_QC_PHI_THREADS : 768; }

// REP: copies of the phi log table (32, 8 or 1). EARLY: the early-termination variant (hard-decision bytes + syndrome
// pass); a separate instantiation so that the default kernel carries none of it (as a run-time flag it cost 2 %).
template <int RULE, int REP, bool EARLY>
__global__ void __launch_bounds__(qc_max_threads(RULE), 1) ldpc_bp_qc_kernel(const __grid_constant__ QcParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const int Z = p.Z, N = p.N, Zb = (Z + 31) >> 5;
    // carve-up by byte offsets from the __shared__ base (keeps the shared address space visible to the compiler)
    const int off_llr = p.E_alloc * 4;
    const int off_col = (off_llr + N * 4 + 15) & ~15;
    const int off_row = off_col + p.n_cols * 16;
    const int off_ce = off_row + p.n_rows * 16;
    const int off_bar = (off_ce + p.nnz * 8 + 15) & ~15;
    float* msg = reinterpret_cast<float*>(smem_raw);
    float* llr_s = reinterpret_cast<float*>(smem_raw + off_llr);
    int4* s_col = reinterpret_cast<int4*>(smem_raw + off_col);
    int4* s_row = reinterpret_cast<int4*>(smem_raw + off_row);
    int2* s_ce_p = reinterpret_cast<int2*>(smem_raw + off_ce);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + off_bar);
    int* sat_flag = reinterpret_cast<int*>(smem_raw + off_bar + 8);
    int* unsat = reinterpret_cast<int*>(smem_raw + off_bar + 12);
    const int off_tab = off_bar + 16;                     // phi log table, tab_rep copies interleaved per entry
    // early termination: one hard-decision byte per VN behind the table
    unsigned char* hd = EARLY ? smem_raw + off_tab + (RULE == SB_CN_BOXPLUS_PHI ? REP * SB_LOGTAB_N * 8 : 0) : nullptr;
    const uint32_t msgb = smem_u32(smem_raw);             // 32-bit shared-window addresses for the hot loops
    const uint32_t s_ce = msgb + off_ce;
    // a warp keeps one 32-lane slice `ib` of every block row/column it visits; G warp groups share the rows
    WarpCtx w;
    w.G = W / Zb;
    w.grp = warp / Zb;
    w.lane_i = (warp - w.grp * Zb) * 32 + lane;

    for (int i = tid; i < p.n_cols; i += T) s_col[i] = p.col_info[i];
    for (int i = tid; i < p.n_rows; i += T) s_row[i] = p.row_info[i];
    for (int i = tid; i < p.nnz; i += T) s_ce_p[i] = p.col_edge[i];
    LogTab<REP> lt;
    lt.inv = 0; lt.lane_off = 0;
    if (RULE == SB_CN_BOXPLUS_PHI) {
        // two arrays of SB_LOGTAB_N * REP floats; entry i, copy c at (i * REP + c) * 4: with REP = 32 lane l reads bank l
        float* tab = reinterpret_cast<float*>(smem_raw + off_tab);
        for (int i = tid; i < SB_LOGTAB_N * REP; i += T) {
            tab[i] = sb_logtab_dev[2 * (i / REP)];
            tab[SB_LOGTAB_N * REP + i] = sb_logtab_dev[2 * (i / REP) + 1];
        }
        lt.inv = msgb + off_tab;
        lt.lane_off = 4 * (lane & (REP - 1));
    }
    if (p.use_tma && tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const float clip = p.llr_max;
    const float phi_max = sb_phif(0.f);                   // phi at its lower clipping bound (global-memory table)
    uint32_t tma_phase = 0;

    for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
        // ---- channel LLRs (decoding.py:552-565, 1444-1475), natural VN order -----------------------------------
        const float* row = p.llr + (size_t)b * p.n_in;
        if (p.use_tma) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, (uint32_t)p.n_in * 4u);
                tma_bulk_g2s(msg, row, (uint32_t)p.n_in * 4u, bar);
            }
            mbar_wait(bar, tma_phase);
            tma_phase ^= 1u;
            for (int v = tid; v < N; v += T) {
                int ii = p.in_idx[v];
                float l = ii >= 0 ? msg[ii] : (ii == -1 ? 0.f : -clip);
                llr_s[v] = __fmul_rn(clipf(l, clip), -1.f);
            }
        } else {
            for (int v = tid; v < N; v += T) {
                int ii = p.in_idx[v];
                float l = ii >= 0 ? __ldg(row + ii) : (ii == -1 ? 0.f : -clip);
                llr_s[v] = __fmul_rn(clipf(l, clip), -1.f);
            }
        }
        __syncthreads();
        if (tid == 0) *sat_flag = 0;
        // ---- v2c = llr of the edge's VN (decoding.py:571) ---------------------------------------------------------
        vn_all<1, true>(p, w, msgb, llr_s, s_col, s_ce, clip, false, true, b);
        __syncthreads();
        if (p.num_iter == 0) {
            for (int v = tid; v < N; v += T) {
                int o = p.out_pos[v];
                if (o >= 0) {
                    float x = llr_s[v];
                    p.out[(size_t)b * p.n_out + o] = p.hard_out ? (0.f >= x ? 1.f : 0.f) : __fmul_rn(x, -1.f);
                }
            }
        }
        // Early termination (opt-in; the reference always runs num_iter iterations, decoding.py:105-107). The VN phase
        // keeps the hard decision of every VN in hd[]; before iteration `it` (it >= 1) a syndrome pass checks H * hd = 0.
        // If it holds, iteration `it` becomes the final one (its CN phase does not fuse the degree-1 updates, its VN phase
        // writes the outputs): the outputs equal a fixed-iteration decode with num_iter = it + 1 bit for bit.
        int limit = p.num_iter;
        for (int it = 0; it < limit; ++it) {
            if (EARLY && it > 0 && it < limit - 1) {
                if (tid == 0) *unsat = 0;
                __syncthreads();
                syndrome_pass(p, w, s_row, hd, unsat);
                __syncthreads();
                if (*unsat == 0) limit = it + 1;               // CTA-uniform: read between barriers
            }
            const bool final_pass = it == limit - 1;
            const bool sc = *sat_flag != 0;                // CTA-uniform: read after the barrier that ended the last phase
            // ---- CN phase (degree-1 VN updates fused in, except in the final iteration) -------------------------
            if (final_pass && tid == 0 && p.use_tma && b + gridDim.x < p.B)   // pull the next codeword's logits into L2 early
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.llr + (size_t)(b + gridDim.x) * p.n_in),
                             "r"((uint32_t)p.n_in * 4u) : "memory");
            const int* re = p.row_cls_end;
            cn_class<RULE, 0, LogTab<REP>>(p, w, msg, llr_s, s_row, 0, re[0], clip, !final_pass, phi_max, sc, sat_flag, lt, hd);
            cn_class<RULE, 1, LogTab<REP>>(p, w, msg, llr_s, s_row, re[0], re[1], clip, !final_pass, phi_max, sc, sat_flag, lt, hd);
            cn_class<RULE, 2, LogTab<REP>>(p, w, msg, llr_s, s_row, re[1], re[2], clip, !final_pass, phi_max, sc, sat_flag, lt, hd);
            cn_class<RULE, 3, LogTab<REP>>(p, w, msg, llr_s, s_row, re[2], re[3], clip, !final_pass, phi_max, sc, sat_flag, lt, hd);
            cn_class<RULE, 4, LogTab<REP>>(p, w, msg, llr_s, s_row, re[3], re[4], clip, !final_pass, phi_max, sc, sat_flag, lt, hd);
            __syncthreads();
            // ---- VN phase ---------------------------------------------------------------------------------------
            vn_all<0, true>(p, w, msgb, llr_s, s_col, s_ce, clip, final_pass, final_pass, b, hd);
            __syncthreads();
        }
        if (EARLY && p.iters_out && tid == 0) p.iters_out[b] = limit;
        if (p.state_out) {
            float* st = p.state_out + (size_t)b * p.E;
            for (int e = tid; e < p.E; e += T) st[e] = __fmul_rn(msg[p.slot_of_edge[e]], -1.f);
            __syncthreads();
        }
    }
}

size_t qc_smem_bytes(const sb_ldpc_graph* g, int tab_rep, int early = 0) {
    return (early ? (size_t)g->N + 16 : 0) + ((size_t)g->qc_nnz * g->qc_Z + g->N) * 4 + 16 + (size_t)g->qc_cols * 16 + (size_t)g->qc_rows * 16 +
           (size_t)g->qc_nnz * 8 + 16 + 16 + (size_t)tab_rep * SB_LOGTAB_N * 8;
}

template <typename T>
int upload_ints(T** d, const std::vector<T>& h) {
    SB_CUDA(cudaMalloc((void**)d, std::max<size_t>(1, h.size()) * sizeof(T)));
    if (h.size()) SB_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SB_OK;
}

int qc_ensure_uploaded(sb_ldpc_graph* g) {
    if (g->qc_uploaded) return SB_OK;
    int rc;
    if ((rc = upload_ints(&g->d_qc_row_info, g->qc_row_info))) return rc;
    if ((rc = upload_ints(&g->d_qc_col_info, g->qc_col_info))) return rc;
    if ((rc = upload_ints(&g->d_qc_col_edge, g->qc_col_edge))) return rc;
    if ((rc = upload_ints(&g->d_qc_in_idx, g->qc_in_idx))) return rc;
    if ((rc = upload_ints(&g->d_qc_out_pos, g->qc_out_pos))) return rc;
    if ((rc = upload_ints(&g->d_qc_slot_of_edge, g->qc_slot_of_edge))) return rc;
    if ((rc = upload_ints(&g->d_qc_row_edge, g->qc_row_edge))) return rc;
    g->qc_uploaded = true;
    return SB_OK;
}

template <int RULE, bool EARLY>
int launch_qc_e(const sb_ldpc_graph* g, const QcParams& p, int threads, size_t smem, cudaStream_t stream) {
    auto kern = ldpc_bp_qc_kernel<RULE, 32, EARLY>;
    if constexpr (RULE == SB_CN_BOXPLUS_PHI) {
        if (p.tab_rep == 8) kern = ldpc_bp_qc_kernel<RULE, 8, EARLY>;
        if (p.tab_rep == 1) kern = ldpc_bp_qc_kernel<RULE, 1, EARLY>;
    }
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) { sb_set_error("sb_ldpc_decode(qc): kernel does not fit (threads %d, smem %zu)", threads, smem); return SB_EUNSUPPORTED; }
    long long grid = std::min<long long>(p.B, (long long)g->num_sms * occ);
    kern<<<(unsigned)grid, threads, smem, stream>>>(p);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

template <int RULE>
int launch_qc(const sb_ldpc_graph* g, const QcParams& p, int threads, size_t smem, cudaStream_t stream) {
    return p.early ? launch_qc_e<RULE, true>(g, p, threads, smem, stream) : launch_qc_e<RULE, false>(g, p, threads, smem, stream);
}

}  // namespace

void sb_qc_free_device(sb_ldpc_graph* g) {
    if (!g->qc_uploaded) return;
    cudaFree(g->d_qc_row_info); cudaFree(g->d_qc_col_info); cudaFree(g->d_qc_col_edge); cudaFree(g->d_qc_in_idx);
    cudaFree(g->d_qc_out_pos); cudaFree(g->d_qc_slot_of_edge); cudaFree(g->d_qc_row_edge);
    g->qc_uploaded = false;
}

// Attach the quasi-cyclic description of the graph: base entries (row, col, shift) of the lifted matrix with lifting
// size Z; entries outside ceil(C/Z) x ceil(N/Z) are ignored (pruned away). The description is verified against the
// handle's edge list; on mismatch the handle is left unchanged and SB_EINVAL is returned.
extern "C" int sb_ldpc_graph_set_qc(sb_ldpc_graph* g, int32_t Z, int32_t n_entries, const int32_t* base_row,
                                    const int32_t* base_col, const int32_t* shift) {
    SB_CHECK_ARG(g && Z > 0 && Z <= 16383 && n_entries > 0 && base_row && base_col && shift, "sb_ldpc_graph_set_qc: bad arguments");
    SB_CHECK_ARG((int)g->h_cn.size() == g->E, "sb_ldpc_graph_set_qc: handle holds no edge list");
    SB_CHECK_ARG(!g->ref_order, "sb_ldpc_graph_set_qc: the QC kernel sums in ascending neighbour order; graphs created with "
                                "sb_ldpc_graph_create_ordered stay on the generic kernel");
    const int C = g->C, N = g->N, E = g->E;
    const int n_rows = (C + Z - 1) / Z, n_cols = (N + Z - 1) / Z;
    auto zrow = [&](int r) { return std::min(Z, C - r * Z); };
    auto zcol = [&](int c) { return std::min(Z, N - c * Z); };
    struct Ent { int r, c, s; };
    std::vector<Ent> ents;
    for (int k = 0; k < n_entries; ++k) {
        if (base_row[k] < 0 || base_col[k] < 0 || shift[k] < 0) { sb_set_error("sb_ldpc_graph_set_qc: negative entry"); return SB_EINVAL; }
        if (base_row[k] >= n_rows || base_col[k] >= n_cols) continue;
        ents.push_back({base_row[k], base_col[k], shift[k] % Z});
    }
    // verify against the edge list
    std::vector<long long> keys(E);
    for (int e = 0; e < E; ++e) keys[e] = ((long long)g->h_cn[e] << 32) | (unsigned)g->h_vn[e];
    std::sort(keys.begin(), keys.end());
    long long total = 0;
    for (const Ent& en : ents)
        for (int i = 0; i < zrow(en.r); ++i) {
            int j = (i + en.s) % Z;
            long long key = ((long long)(en.r * Z + i) << 32) | (unsigned)(en.c * Z + j);
            if (j >= zcol(en.c) || !std::binary_search(keys.begin(), keys.end(), key)) {
                sb_set_error("sb_ldpc_graph_set_qc: base entry (%d,%d,%d) does not match the graph", en.r, en.c, en.s);
                return SB_EINVAL;
            }
            ++total;
        }
    if (total != E) { sb_set_error("sb_ldpc_graph_set_qc: %lld lifted edges != %d graph edges", total, E); return SB_EINVAL; }
    // classes (see the kernel): rows by degree bucket, heaviest first; columns by bucket / pruning-cut rows / fused
    std::vector<int> rdeg(n_rows, 0), cdeg(n_cols, 0);
    for (const Ent& en : ents) { ++rdeg[en.r]; ++cdeg[en.c]; }
    std::vector<std::vector<Ent>> by_row(n_rows), by_col(n_cols);
    for (const Ent& en : ents) { by_row[en.r].push_back(en); by_col[en.c].push_back(en); }
    for (auto& v : by_row) std::sort(v.begin(), v.end(), [](const Ent& a, const Ent& b) { return a.c < b.c; });
    for (auto& v : by_col) std::sort(v.begin(), v.end(), [](const Ent& a, const Ent& b) { return a.r < b.r; });
    // a degree-1 column whose single entry is the LAST entry of its row is updated inside the CN phase
    std::vector<int> fused_col_of_row(n_rows, -1);
    std::vector<char> col_fused(n_cols, 0);
    for (int r = 0; r < n_rows; ++r) {
        if (by_row[r].empty()) continue;
        const Ent& last = by_row[r].back();
        if (cdeg[last.c] == 1 && rdeg[r] >= 2 && last.c < 65536 && zcol(last.c) == zrow(r)) {
            fused_col_of_row[r] = last.c;
            col_fused[last.c] = 1;
        }
    }
    auto row_class = [&](int r) { int d = rdeg[r]; return d > 20 ? 0 : d > 12 ? 1 : d > 8 ? 2 : d > 4 ? 3 : 4; };
    auto col_class = [&](int c) {
        if (col_fused[c]) return 10;
        bool check = false;
        for (const Ent& en : by_col[c]) check = check || zrow(en.r) < Z;
        int d = cdeg[c];
        if (check) return d > 12 ? 7 : d > 4 ? 8 : 9;
        return d > 32 ? 0 : d > 20 ? 1 : d > 12 ? 2 : d > 8 ? 3 : d > 4 ? 4 : d > 2 ? 5 : 6;
    };
    std::vector<int> rorder(n_rows), corder(n_cols);
    std::iota(rorder.begin(), rorder.end(), 0);
    std::iota(corder.begin(), corder.end(), 0);
    std::stable_sort(rorder.begin(), rorder.end(), [&](int a, int b) {
        return row_class(a) != row_class(b) ? row_class(a) < row_class(b) : rdeg[a] > rdeg[b]; });
    std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) {
        return col_class(a) != col_class(b) ? col_class(a) < col_class(b) : cdeg[a] > cdeg[b]; });
    // Inside a class the order is free. Rows / columns are dealt cyclically to the G warp groups of a launch over ALL
    // classes (position i -> group i mod G), so the order decides the load balance: walk the positions in rounds of G
    // consecutive ones (a round gives every group at most one item) and hand the round's heaviest item to the group with
    // the smallest load so far. Degree order alone left the benchmark graph's four groups with 56/53/52/49 edges per lane
    // in the CN phase and 56/53/41/40 in the VN phase; this gives 54/54/53/49 and 48/48/47/47. G is fixed by Z and the CTA
    // size (qc_max_threads / 32 / ceil(Z / 32)), the same for every rule.
    {
        const int Zb_ = (Z + 31) / 32;
        const int G = std::max(1, (qc_max_threads(0) / 32) / Zb_);
        auto balance = [&](std::vector<int>& order, const std::vector<int>& deg, auto cls_of) {
            std::vector<long long> load(G, 0);
            size_t pos = 0;
            while (pos < order.size()) {
                size_t end = pos;
                const int c = cls_of(order[pos]);
                while (end < order.size() && cls_of(order[end]) == c) ++end;      // one class: positions [pos, end)
                std::vector<int> items(order.begin() + pos, order.begin() + end);
                std::stable_sort(items.begin(), items.end(), [&](int a, int b) { return deg[a] > deg[b]; });
                size_t next = 0;
                for (size_t r0 = pos; r0 < end; r0 += G) {
                    const size_t r1 = std::min(end, r0 + (size_t)G);
                    std::vector<size_t> slots;
                    for (size_t q = r0; q < r1; ++q) slots.push_back(q);
                    std::stable_sort(slots.begin(), slots.end(), [&](size_t a, size_t b) { return load[a % G] < load[b % G]; });
                    for (size_t q : slots) {
                        order[q] = items[next++];
                        load[q % G] += deg[order[q]];
                    }
                }
                pos = end;
            }
        };
        balance(rorder, rdeg, row_class);
        balance(corder, cdeg, col_class);
    }
    std::vector<int> row_cls_end(5, 0), col_cls_end(11, 0);
    for (int r = 0; r < n_rows; ++r) for (int k = row_class(r); k < 5; ++k) ++row_cls_end[k];
    for (int c = 0; c < n_cols; ++c) for (int k = col_class(c); k < 11; ++k) ++col_cls_end[k];
    // base-entry numbering: rows in processing order, ascending column inside a row
    std::vector<int> be_of((size_t)n_rows * n_cols, -1);
    std::vector<int> row_info(4 * n_rows), row_edge;
    int be = 0;
    for (int rr = 0; rr < n_rows; ++rr) {
        int r = rorder[rr];
        row_info[4 * rr] = be;
        row_info[4 * rr + 1] = rdeg[r];
        row_info[4 * rr + 2] = zrow(r);
        row_info[4 * rr + 3] = fused_col_of_row[r] >= 0 ? (fused_col_of_row[r] | (by_row[r].back().s << 16)) : -1;
        for (const Ent& en : by_row[r]) {
            if (be_of[(size_t)en.r * n_cols + en.c] != -1) { sb_set_error("sb_ldpc_graph_set_qc: duplicate base entry"); return SB_EINVAL; }
            be_of[(size_t)en.r * n_cols + en.c] = be++;
            row_edge.push_back(en.c * Z);
            row_edge.push_back(en.s);
        }
    }
    const int nnz = be;
    std::vector<int> col_info(4 * n_cols), col_edge(2 * (size_t)nnz);
    int ce = 0;
    for (int cc = 0; cc < n_cols; ++cc) {
        int c = corder[cc];
        col_info[4 * cc] = ce;
        for (const Ent& en : by_col[c]) {
            col_edge[2 * ce] = be_of[(size_t)en.r * n_cols + en.c] * Z * 4;
            col_edge[2 * ce + 1] = (en.s * 4) | ((zrow(en.r) * 4) << 16);
            ++ce;
        }
        col_info[4 * cc + 1] = cdeg[c];
        col_info[4 * cc + 2] = zcol(c);
        col_info[4 * cc + 3] = c * Z;
    }
    // natural-order I/O maps and the reference-edge -> slot map
    std::vector<int> in_nat(N), out_nat(N), slot(E);
    for (int r = 0; r < N; ++r) { in_nat[g->vn_order[r]] = g->in_idx[r]; out_nat[g->vn_order[r]] = g->out_pos[r]; }
    for (int e = 0; e < E; ++e) {
        int r = g->h_cn[e] / Z, i = g->h_cn[e] % Z, c = g->h_vn[e] / Z;
        slot[e] = be_of[(size_t)r * n_cols + c] * Z + i;
    }
    sb_qc_free_device(g);
    g->qc = true; g->qc_Z = Z; g->qc_rows = n_rows; g->qc_cols = n_cols; g->qc_nnz = nnz;
    g->qc_max_row_deg = *std::max_element(rdeg.begin(), rdeg.end());
    g->qc_max_col_deg = *std::max_element(cdeg.begin(), cdeg.end());
    g->qc_row_info.swap(row_info); g->qc_col_info.swap(col_info); g->qc_col_edge.swap(col_edge);
    g->qc_row_cls_end = row_cls_end; g->qc_col_cls_end = col_cls_end;
    g->qc_in_idx.swap(in_nat); g->qc_out_pos.swap(out_nat); g->qc_slot_of_edge.swap(slot); g->qc_row_edge.swap(row_edge);
    return SB_OK;
}

// Test hook: evaluates phi on the device with the scalar (sb_math.h) and the packed (sb_math2.cuh) implementation.
namespace {
__global__ void debug_phi_kernel(const float* x, float* o1, float* o2, long long n) {
    __shared__ float tab[2 * SB_LOGTAB_N * 32];           // the 32-copy layout of the decoder
    for (int i = threadIdx.x; i < SB_LOGTAB_N * 32; i += blockDim.x) {
        tab[i] = sb_logtab_dev[2 * (i / 32)];
        tab[SB_LOGTAB_N * 32 + i] = sb_logtab_dev[2 * (i / 32) + 1];
    }
    __syncthreads();
    LogTab<32> lt;
    lt.inv = smem_u32(tab);
    lt.lane_off = 4 * (threadIdx.x & 31);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 < n) {
        float2 r = sb_phif2(make_float2(x[2 * i], x[2 * i + 1]), lt);
        o2[2 * i] = r.x; o2[2 * i + 1] = r.y;
        o1[2 * i] = sb_phif(x[2 * i]); o1[2 * i + 1] = sb_phif(x[2 * i + 1]);
    }
}
}  // namespace
extern "C" int sb_debug_phi(const float* d_x, float* d_scalar, float* d_packed, int64_t n, void* stream) {
    if (n == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_x && d_scalar && d_packed && n >= 0 && n % 2 == 0, "sb_debug_phi: bad arguments");
    if (n == 0) return SB_OK;
    debug_phi_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_scalar, d_packed, n);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_ldpc_graph_is_qc(const sb_ldpc_graph* g) { return g && g->qc ? 1 : 0; }

int sb_qc_try_decode(sb_ldpc_graph* g, const float* d_llr, int64_t batch, int32_t num_iter, int32_t cn_rule,
                     int32_t vn_rule, float offset, float llr_max, int32_t hard_out, const float* d_state_in,
                     float* d_state_out, float* d_out, cudaStream_t stream, bool* handled, int32_t early, int32_t* d_iters) {
    *handled = false;
    if (!g->qc || !g->flooding || vn_rule != SB_VN_SUM || d_state_in || cn_rule > SB_CN_OFFSET_MINSUM) return SB_OK;
    // boxplus-phi keeps the log table of phi in shared memory: one copy per bank pair if it fits, else a single copy
    int tab_rep = cn_rule == SB_CN_BOXPLUS_PHI ? 32 : 0;
    if (tab_rep && qc_smem_bytes(g, tab_rep, early) > (size_t)g->smem_optin) tab_rep = 8;
    if (tab_rep && qc_smem_bytes(g, tab_rep, early) > (size_t)g->smem_optin) tab_rep = 1;
    const size_t smem = qc_smem_bytes(g, tab_rep, early);
    if (smem > (size_t)g->smem_optin) return SB_OK;
    if ((cn_rule == SB_CN_MINSUM || cn_rule == SB_CN_OFFSET_MINSUM) &&
        !(llr_max < 100000.f && (float)(g->qc_max_row_deg - 1) * llr_max < 99000.f))
        return SB_OK;                                      // the generic kernel has the literal 1e5-sentinel path
    int rc = qc_ensure_uploaded(g);
    if (rc) return rc;
    QcParams p{};
    p.Z = g->qc_Z; p.n_rows = g->qc_rows; p.n_cols = g->qc_cols; p.nnz = g->qc_nnz; p.N = g->N; p.E = g->E;
    p.E_alloc = g->qc_nnz * g->qc_Z; p.n_in = g->n_in; p.n_out = g->n_out;
    for (int k = 0; k < kRowClasses; ++k) p.row_cls_end[k] = g->qc_row_cls_end[k];
    for (int k = 0; k < kColClasses; ++k) p.col_cls_end[k] = g->qc_col_cls_end[k];
    p.row_info = (const int4*)g->d_qc_row_info; p.col_info = (const int4*)g->d_qc_col_info;
    p.col_edge = (const int2*)g->d_qc_col_edge; p.in_idx = g->d_qc_in_idx; p.out_pos = g->d_qc_out_pos;
    p.slot_of_edge = g->d_qc_slot_of_edge;
    p.llr = d_llr; p.out = d_out; p.state_out = d_state_out; p.B = batch; p.num_iter = num_iter; p.hard_out = hard_out;
    p.offset = offset; p.llr_max = llr_max; p.tab_rep = tab_rep;
    p.early = early; p.iters_out = d_iters; p.row_edge = (const int2*)g->d_qc_row_edge;
    p.use_tma = (g->n_in % 4 == 0) && (g->n_in <= p.E_alloc) && ((reinterpret_cast<uintptr_t>(d_llr) & 15) == 0);
    const int Zb = (g->qc_Z + 31) / 32;                    // 32-lane slices per block row (<= 12 for Z <= 384)
    const int max_warps = qc_max_threads(cn_rule) / 32;
    int groups = std::max(1, std::min(max_warps / Zb, std::max(g->qc_rows, g->qc_cols)));
    const int threads = groups * Zb * 32;                  // every warp owns one slice index for the whole launch
    for (int k = 0; k < kRowClasses; ++k) p.row_cls_mod[k] = (k ? g->qc_row_cls_end[k - 1] : 0) % groups;
    for (int k = 0; k < kColClasses; ++k) p.col_cls_mod[k] = (k ? g->qc_col_cls_end[k - 1] : 0) % groups;
    switch (cn_rule) {
        case SB_CN_BOXPLUS_PHI: rc = launch_qc<SB_CN_BOXPLUS_PHI>(g, p, threads, smem, stream); break;
        case SB_CN_BOXPLUS: rc = launch_qc<SB_CN_BOXPLUS>(g, p, threads, smem, stream); break;
        case SB_CN_MINSUM: rc = launch_qc<SB_CN_MINSUM>(g, p, threads, smem, stream); break;
        default: rc = launch_qc<SB_CN_OFFSET_MINSUM>(g, p, threads, smem, stream); break;
    }
    *handled = (rc == SB_OK);
    return rc;
}
