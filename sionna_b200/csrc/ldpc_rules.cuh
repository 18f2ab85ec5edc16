// ldpc_rules.cuh -- check-node update rules of the belief-propagation decoder, written once against an edge ACCESSOR so
// that the shared-memory kernel (ldpc_bp.cu: slot = off[l] + rank) and the unfused [num_edges, batch] kernels
// (ldpc_bp_flat.cu) run literally the same arithmetic in the same order.
//
// An accessor `A` gives the l-th edge (list order) of ONE check node:
//     float A::in(int l)            incoming v2c message
//     void  A::out(int l, float v)  store the outgoing c2v message (also used to stage per-edge intermediates)
//     float A::staged(int l)        read back what out(l, .) stored
// Reference: /root/reference/src/sionna/phy/fec/ldpc/decoding.py, cn_update_phi :1045-1166, cn_update_tanh :955-1043,
// cn_update_offset_minsum :755-909, cn_update_minsum :911-953.
#pragma once
#include "sb_math.h"
#include "ldpc_graph.h"

// boxplus-phi, decoding.py:1126-1166
template <class A>
__device__ __forceinline__ void cn_phi(const A& e, int deg, float clip) {
    float P = 0.f;
    unsigned par = 0;
#pragma unroll 4
    for (int l = 0; l < deg; ++l) {
        float x = e.in(l);
        unsigned neg = x < 0.f;                         // sign(0) := +1 (:1129)
        par ^= neg;
        float p = sb_phif(fabsf(x));                    // >= 0 for every input (tools/check_math.c)
        P = __fadd_rn(P, p);                            // :1150 sequential sum
        e.out(l, neg ? -p : p);                         // stage phi(|x|), sign bit carries sign(x)
    }
#pragma unroll 4
    for (int l = 0; l < deg; ++l) {
        unsigned bits = __float_as_uint(e.staged(l));
        unsigned neg = bits >> 31;
        float p = __uint_as_float(bits & 0x7fffffffu);
        float y = sb_phif(__fadd_rn(-p, P));            // :1155-1161
        y = (neg ^ par) ? -y : y;                       // extrinsic sign = sign(x_e) * prod(signs)
        e.out(l, clipf(y, clip));                       // :1163
    }
}

// boxplus (tanh), decoding.py:1000-1043
template <class A>
__device__ __forceinline__ void cn_tanh(const A& e, int deg, float clip) {
    const float atanh_clip = (float)(1 - 1e-7);
    float prod = 1.f;
#pragma unroll 2
    for (int l = 0; l < deg; ++l) {
        float t = sb_tanhf(__fmul_rn(e.in(l), 0.5f));   // x/2 == x*0.5 exactly
        if (t == 0.f) t = 1e-12f;
        prod = __fmul_rn(prod, t);
        e.out(l, t);
    }
#pragma unroll 2
    for (int l = 0; l < deg; ++l) {
        float q = __fmul_rn(__fdiv_rn(1.f, e.staged(l)), prod);
        if (fabsf(q) < 1e-7f) q = 0.f;
        q = clipf(q, atanh_clip);
        float y = __fmul_rn(2.f, sb_atanhf(q));
        e.out(l, clipf(y, clip));
    }
}

// (offset-)min-sum, decoding.py:796-909. The reference's "subtract min, replace zeros by 1e5, take the
// min again, detect duplicate minima through the row sum" sequence is reproduced exactly:
//   unique minimum  -> that edge gets fl(fl(m2 - m1) + m1), every other edge m1
//   repeated minimum-> every edge gets m1
// A slow path redoes the reference's row sum literally when magnitudes are large enough (>= ~1e5/deg)
// for the sum test or the "== 1e5" test to behave differently.
template <class A>
__device__ __forceinline__ void cn_minsum(const A& e, int deg, float clip, float offset) {
    const float large_val = 100000.f;
    float m1 = INFINITY, m2 = INFINITY, amax = 0.f;
    unsigned par = 0;
    int cnt = 0;
#pragma unroll 4
    for (int l = 0; l < deg; ++l) {
        float x = clipf(e.in(l), large_val);            // :808
        par ^= (unsigned)(x < 0.f);
        float a = fabsf(x);
        amax = fmaxf(amax, a);
        if (a < m1) { m2 = m1; m1 = a; cnt = 1; }
        else if (a == m1) { ++cnt; m2 = m1; }
        else if (a < m2) { m2 = a; }
    }
    float min_e;                                        // value written at the minimum position(s)
    bool literal = (float)(deg - 1) * (amax - m1) >= 99000.f;
    if (!literal) {
        min_e = (cnt >= 2) ? m1 : __fadd_rn(__fsub_rn(m2, m1), m1);   // :863, :876
        if (deg == 1) min_e = __fadd_rn(large_val, m1);               // single edge: min over {1e5}
    } else {
        float min2 = INFINITY, node_sum = 0.f;
        for (int l = 0; l < deg; ++l) {
            float a = fabsf(clipf(e.in(l), large_val));
            float d = __fsub_rn(a, m1);
            if (d == 0.f) d = large_val;
            min2 = fminf(min2, d);
            node_sum = __fadd_rn(node_sum, d);
        }
        float min_val_2 = __fadd_rn(min2, m1);
        node_sum = __fsub_rn(node_sum, 199999.f);
        float sg = node_sum > 0.f ? 1.f : (node_sum < 0.f ? -1.f : 0.f);
        float dm = __fmul_rn(0.5f, __fsub_rn(1.f, sg));
        min_e = __fadd_rn(__fmul_rn(__fsub_rn(1.f, dm), m1), __fmul_rn(dm, min_val_2));
    }
#pragma unroll 4
    for (int l = 0; l < deg; ++l) {
        float x = clipf(e.in(l), large_val);
        unsigned neg = x < 0.f;
        float a = fabsf(x);
        bool at_min = literal ? (__fsub_rn(a, m1) == 0.f || __fsub_rn(a, m1) == large_val) : (a == m1);
        float m = at_min ? min_e : m1;                  // :886
        m = fmaxf(__fsub_rn(m, offset), 0.f);           // :895
        m = (neg ^ par) ? -m : m;                       // :903
        e.out(l, clipf(m, clip));                       // :906
    }
}

template <class A>
__device__ __forceinline__ void cn_identity(const A& e, int deg) {
    for (int l = 0; l < deg; ++l) e.out(l, e.in(l));
}

template <int RULE, class A>
__device__ __forceinline__ void cn_node(const A& e, int deg, float clip, float offset) {
    if (RULE == SB_CN_BOXPLUS_PHI) cn_phi(e, deg, clip);
    else if (RULE == SB_CN_BOXPLUS) cn_tanh(e, deg, clip);
    else if (RULE == SB_CN_MINSUM) cn_minsum(e, deg, clip, 0.f);
    else if (RULE == SB_CN_OFFSET_MINSUM) cn_minsum(e, deg, clip, offset);
    else cn_identity(e, deg);
}
