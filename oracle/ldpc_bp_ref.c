/* oracle/ldpc_bp_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32, one codeword at a time) of the reference's belief-propagation
 * decoder, /root/reference/src/sionna/phy/fec/ldpc/decoding.py:
 *   LDPCBPDecoder.call          :544-637   -> sbo_bp_decode()
 *   LDPCBPDecoder._bp_iter      :416-524   -> loop body of sbo_bp_decode()
 *   vn_update_sum               :681-732   -> vn_update()
 *   cn_update_offset_minsum     :755-909   -> cn_offset_minsum()
 *   cn_update_minsum            :911-953   -> cn_offset_minsum(offset = 0)
 *   cn_update_tanh ("boxplus")  :955-1043  -> cn_tanh()
 *   cn_update_phi               :1045-1166 -> cn_phi()
 *   cn/vn identity              :644-679, :735-753
 * Same operations, same order, fp32 throughout; no "numerically nicer" substitutions.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library. The product (sionna_b200/) never does.
 *
 * The reference's arithmetic runs inside TensorFlow (not vendored under /root/reference,
 * pinned as tensorflow>=2.14,!=2.16,!=2.17 in pyproject.toml:55), which is not importable in
 * the build container: bit-level parity with TF's exp/log kernels is UNPINNED. What is pinned
 * (tests/test_oracle_*.py): the in-test NumPy node-update loops of
 * /root/reference/test/unit/fec/test_ldpc_decoding.py:397-654 (restated in the tests), the
 * duplicate-minimum, all-erasure, llr_max-bound and 0-iteration KATs of the same file.
 *
 * math_mode 0: glibc expf/logf/tanhf/atanhf (a <=1 ulp libm, like the one TF uses).
 * math_mode 1: sionna_b200/csrc/sb_math.h, the deterministic functions the CUDA kernels use;
 *              in this mode the CUDA path must agree with this file bit for bit.
 *
 * Edge bookkeeping follows decoding.py:277-345: edges are numbered in VN order (the reference's
 * argsort(vn_idx) order, supplied by the caller); vn_ptr[v]..vn_ptr[v+1] are the edges of VN v;
 * cn_edge[cn_ptr[c]..cn_ptr[c+1]] lists, in the caller's CN-view order (v2c_perm), the edge
 * numbers of CN c. Reductions over a node run sequentially in those list orders.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
/* -DSBO_PURE_LIBM builds oracle/_build/libsbo_libm.so: this file WITHOUT the product's sb_math.h (math_mode 1 is then
 * refused), i.e. an oracle that shares no line of code with sionna_b200/. The order="reference" parity tests of the
 * rules without transcendental functions (minsum, offset-minsum) and the libm-mode comparisons run against it. */
#ifndef SBO_PURE_LIBM
#include "../sionna_b200/csrc/sb_math.h"
#else
static inline float sb_expf(float x) { (void)x; return NAN; }
static inline float sb_logf(float x) { (void)x; return NAN; }
static inline float sb_logf_tab(float x) { (void)x; return NAN; }
static inline float sb_tanhf(float x) { (void)x; return NAN; }
static inline float sb_atanhf(float x) { (void)x; return NAN; }
#endif

enum { SBO_CN_PHI = 0, SBO_CN_TANH = 1, SBO_CN_MINSUM = 2, SBO_CN_OFFSET_MINSUM = 3, SBO_CN_IDENTITY = 4 };
enum { SBO_VN_SUM = 0, SBO_VN_IDENTITY = 1 };

static inline float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float signf0(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }   /* tf.sign */

static inline float m_exp(float x, int mode) { return mode ? sb_expf(x) : expf(x); }
static inline float m_log(float x, int mode) { return mode ? sb_logf(x) : logf(x); }
static inline float m_tanh(float x, int mode) { return mode ? sb_tanhf(x) : tanhf(x); }
static inline float m_atanh(float x, int mode) { return mode ? sb_atanhf(x) : atanhf(x); }

/* decoding.py:1110-1120. In kernel-math mode the two logs are the table-driven sb_logf_tab (sb_math.h), which is what
 * the CUDA kernels evaluate inside phi. */
static inline float m_log_phi(float x, int mode) { return mode ? sb_logf_tab(x) : logf(x); }
static inline float phi(float x, int mode) {
    x = clipf(x, 8.5e-8f, 16.635532f);
    float t = m_exp(x, mode);
    return m_log_phi(t + 1.f, mode) - m_log_phi(t - 1.f, mode);
}

/* decoding.py:1126-1166. x: incoming v2c of one CN (deg values), out: c2v. */
static void cn_phi(const float* x, float* out, int deg, int has_clip, float clip, int mode) {
    float sign_node = 1.f, sum = 0.f;
    for (int i = 0; i < deg; ++i) {
        float s = signf0(x[i]);
        if (s == 0.f) s = 1.f;                      /* :1129 */
        sign_node *= s;                             /* :1132 */
        out[i] = phi(fabsf(x[i]), mode);            /* :1144-1147, staged in out[] */
    }
    for (int i = 0; i < deg; ++i) sum += out[i];    /* :1150 */
    for (int i = 0; i < deg; ++i) {
        float s = signf0(x[i]);
        if (s == 0.f) s = 1.f;
        float s_ext = s * sign_node;                /* :1137 */
        float m = (-1.f * out[i]) + sum;            /* :1155 */
        float y = s_ext * phi(m, mode);             /* :1161 */
        if (has_clip) y = clipf(y, -clip, clip);    /* :1163 */
        out[i] = y;
    }
}

/* decoding.py:1000-1043 */
static void cn_tanh(const float* x, float* out, int deg, int has_clip, float clip, int mode) {
    const float atanh_clip = (float)(1 - 1e-7);     /* :1001 */
    float prod = 1.f;
    for (int i = 0; i < deg; ++i) {
        float t = m_tanh(x[i] / 2.f, mode);         /* :1003-1005 */
        if (t == 0.f) t = 1e-12f;                   /* :1008 */
        out[i] = t;
        prod *= t;                                  /* :1011 */
    }
    for (int i = 0; i < deg; ++i) {
        float e = (1.f / out[i]) * prod;            /* :1020-1022 (msg**-1 * prod) */
        if (fabsf(e) < 1e-7f) e = 0.f;              /* :1028 */
        e = clipf(e, -atanh_clip, atanh_clip);      /* :1031 */
        float y = 2.f * m_atanh(e, mode);           /* :1036 */
        if (has_clip) y = clipf(y, -clip, clip);
        out[i] = y;
    }
}

/* decoding.py:796-909 */
static void cn_offset_minsum(const float* xin, float* out, int deg, int has_clip, float clip, float offset) {
    const float large_val = 100000.f;
    float sign_node = 1.f, min_val = INFINITY;
    for (int i = 0; i < deg; ++i) {
        float x = clipf(xin[i], -large_val, large_val);      /* :808 */
        float s = signf0(x);
        if (s == 0.f) s = 1.f;                               /* :800-804 */
        sign_node *= s;                                      /* :816 */
        float a = fabsf(x);                                  /* :831 */
        if (a < min_val) min_val = a;                        /* :842 */
    }
    float min2 = INFINITY, node_sum = 0.f;
    for (int i = 0; i < deg; ++i) {
        float a = fabsf(clipf(xin[i], -large_val, large_val));
        float d = a - min_val;                               /* :849 */
        if (d == 0.f) d = large_val;                         /* :857 */
        out[i] = d;                                          /* staged */
        if (d < min2) min2 = d;
        node_sum += d;
    }
    float min_val_2 = min2 + min_val;                        /* :863 */
    node_sum = node_sum - (float)(2 * 100000. - 1.);         /* :870 */
    float double_min = 0.5f * (1.f - signf0(node_sum));      /* :872 */
    float min_val_e = (1.f - double_min) * min_val + double_min * min_val_2;   /* :876 */
    for (int i = 0; i < deg; ++i) {
        float x = clipf(xin[i], -large_val, large_val);
        float s = signf0(x);
        if (s == 0.f) s = 1.f;
        float s_ext = s * sign_node;                         /* :823 */
        float m = (out[i] == large_val) ? min_val_e : min_val;   /* :886 */
        m = fmaxf(m - offset, 0.f);                          /* :895 */
        float y = s_ext * m;                                 /* :903 */
        if (has_clip) y = clipf(y, -clip, clip);             /* :906 */
        out[i] = y;
    }
}

/* Node-level entry points (used by the KAT tests that restate the reference's NumPy loops). */
void sbo_cn_update(int rule, const float* x, float* out, int deg, int has_clip, float clip,
                   float offset, int math_mode) {
    switch (rule) {
    case SBO_CN_PHI: cn_phi(x, out, deg, has_clip, clip, math_mode); break;
    case SBO_CN_TANH: cn_tanh(x, out, deg, has_clip, clip, math_mode); break;
    case SBO_CN_MINSUM: cn_offset_minsum(x, out, deg, has_clip, clip, 0.f); break;
    case SBO_CN_OFFSET_MINSUM: cn_offset_minsum(x, out, deg, has_clip, clip, offset); break;
    default: memcpy(out, x, sizeof(float) * (size_t)deg); break;
    }
}

/* decoding.py:714-732. c2v: incoming messages of one VN; returns x_tot, writes v2c to out. */
float sbo_vn_update(int rule, const float* c2v, float* out, int deg, float llr_ch, int has_clip, float clip) {
    float x = 0.f;
    for (int i = 0; i < deg; ++i) x += c2v[i];               /* :715 */
    float x_tot = x + llr_ch;                                /* :716 */
    if (rule == SBO_VN_IDENTITY) {                           /* :677-679 */
        for (int i = 0; i < deg; ++i) out[i] = c2v[i];
        return x_tot;
    }
    for (int i = 0; i < deg; ++i) {
        float xe = (-1.f * c2v[i]) + x_tot;                  /* :724 */
        if (has_clip) xe = clipf(xe, -clip, clip);           /* :728 */
        out[i] = xe;
    }
    if (has_clip) x_tot = clipf(x_tot, -clip, clip);         /* :730 */
    return x_tot;
}

float sbo_phi(float x, int math_mode) { return phi(x, math_mode); }
int sbo_is_pure_libm(void) {
#ifdef SBO_PURE_LIBM
    return 1;
#else
    return 0;
#endif
}

/* Full decoder, decoding.py:544-637.
 *  llr_ch   [B, N]  logits as passed to LDPCBPDecoder.call
 *  x_out    [B, N]  hard bits (0/1) or soft logits
 *  state_in / state_out [E, B] (reference layout of msg_v2c), either may be NULL
 *  schedule [n_sub, n_active] CN indices per sub-iteration (flooding: 1 x C, 0..C-1)
 *  flooding != 0 selects the reference's "all nodes are updated" branch (:498-500).
 */
int sbo_bp_decode(int C, int N, int E,
                  const int32_t* vn_ptr, const int32_t* cn_ptr, const int32_t* cn_edge,
                  const int32_t* schedule, int n_sub, int n_active, int flooding,
                  const float* llr_ch, int B, int num_iter,
                  int cn_rule, int vn_rule, float offset, float llr_max,
                  int hard_out, const float* state_in, float* state_out, float* x_out,
                  int math_mode, int num_threads) {
#ifdef SBO_PURE_LIBM
    if (math_mode != 0) return -1;
#endif
    int max_deg = 0;
    for (int c = 0; c < C; ++c) if (cn_ptr[c + 1] - cn_ptr[c] > max_deg) max_deg = cn_ptr[c + 1] - cn_ptr[c];
    for (int v = 0; v < N; ++v) if (vn_ptr[v + 1] - vn_ptr[v] > max_deg) max_deg = vn_ptr[v + 1] - vn_ptr[v];
    int32_t* vn_of_edge = (int32_t*)malloc(sizeof(int32_t) * (size_t)(E > 0 ? E : 1));
    for (int v = 0; v < N; ++v) for (int e = vn_ptr[v]; e < vn_ptr[v + 1]; ++e) vn_of_edge[e] = v;
    const int has_clip = 1;          /* LDPCBPDecoder always passes self.llr_max (:482, :509) */
#if defined(_OPENMP)
    if (num_threads <= 0) num_threads = 1;
#pragma omp parallel num_threads(num_threads)
#endif
    {
        float* v2c = (float*)malloc(sizeof(float) * (size_t)(E + 1));
        float* c2v = (float*)malloc(sizeof(float) * (size_t)(E + 1));
        float* llr = (float*)malloc(sizeof(float) * (size_t)(N + 1));
        float* xh = (float*)malloc(sizeof(float) * (size_t)(N + 1));
        float* tin = (float*)malloc(sizeof(float) * (size_t)(max_deg + 1));
        float* tout = (float*)malloc(sizeof(float) * (size_t)(max_deg + 1));
#if defined(_OPENMP)
#pragma omp for schedule(dynamic, 1)
#endif
        for (int b = 0; b < B; ++b) {
            for (int v = 0; v < N; ++v) {
                float l = clipf(llr_ch[(size_t)b * N + v], -llr_max, llr_max);   /* :552 */
                llr[v] = l * -1.f;                                               /* :565 */
                xh[v] = llr[v];                                                  /* :607 (x_hat = llr_ch) */
            }
            for (int e = 0; e < E; ++e) {
                v2c[e] = state_in ? state_in[(size_t)e * B + b] * -1.f           /* :573 */
                                  : llr[vn_of_edge[e]];                          /* :571 */
                c2v[e] = 0.f;                                                    /* :581 */
            }
            for (int it = 0; it < num_iter; ++it) {
                for (int j = 0; j < n_sub; ++j) {
                    /* CN update of the active nodes (:479-500); for flooding every c2v is
                     * replaced, for a custom schedule only the active CNs' edges (:489-497). */
                    for (int a = 0; a < n_active; ++a) {
                        int c = schedule[(size_t)j * n_active + a];
                        int deg = cn_ptr[c + 1] - cn_ptr[c];
                        const int32_t* ed = cn_edge + cn_ptr[c];
                        for (int i = 0; i < deg; ++i) tin[i] = v2c[ed[i]];
                        sbo_cn_update(cn_rule, tin, tout, deg, has_clip, llr_max, offset, math_mode);
                        for (int i = 0; i < deg; ++i) c2v[ed[i]] = tout[i];
                    }
                    (void)flooding;
                    /* full VN update (:506-519) */
                    for (int v = 0; v < N; ++v) {
                        int s = vn_ptr[v], deg = vn_ptr[v + 1] - s;
                        xh[v] = sbo_vn_update(vn_rule, c2v + s, v2c + s, deg, llr[v], has_clip, llr_max);
                    }
                }
            }
            for (int v = 0; v < N; ++v) {
                float x = xh[v];
                x_out[(size_t)b * N + v] = hard_out ? ((0.f >= x) ? 1.f : 0.f)   /* :623 */
                                                    : x * -1.f;                  /* :626 */
            }
            if (state_out) for (int e = 0; e < E; ++e) state_out[(size_t)e * B + b] = v2c[e] * -1.f;   /* :636 */
        }
        free(v2c); free(c2v); free(llr); free(xh); free(tin); free(tout);
    }
    free(vn_of_edge);
    return 0;
}
