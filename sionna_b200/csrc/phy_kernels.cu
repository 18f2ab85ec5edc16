// phy_kernels.cu -- bytes-bound element-wise kernels of the link-level chain (sm_100a):
//   sb_binary_source   BinarySource.call                 mapping.py:1350-1352
//   sb_qam_map         Mapper.call                       mapping.py:497-519
//   sb_demap           Demapper.call + SymbolLogits2LLRs mapping.py:664-691, 927-967
//   sb_awgn            AWGN.call + complex_normal        channel/awgn.py:63-78, utils/misc.py:19-54
//   sb_count_errors    count_errors / count_block_errors utils/metrics.py:94-144
// (paths relative to /root/reference/src/sionna/phy). All are one pass over HBM with coalesced accesses and
// grids sized as a multiple of the SM count; transcendental functions of the demapper come from sb_math.h so the
// CPU oracle reproduces LLRs bit for bit.
#include "sb_common.h"
#include "sb_math.h"
#include "sb_math2.cuh"
#include "rng.cuh"
#include "demap_qam.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
__global__ void binary_source_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                     unsigned long long offset) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one Philox block -> 128 bits
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long blk = i; blk * 128 < n; blk += stride) {
        uint4 r = philox4x32_10(seed, offset, (unsigned long long)blk);
        unsigned w[4] = {r.x, r.y, r.z, r.w};
        long long base = blk * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                long long idx = base + k * 32 + b;
                if (idx < n) out[idx] = (float)((w[k] >> b) & 1u);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// bits [n_sym, m] (float 0/1) -> points[index], index = sum bit_k << (m-1-k)   (mapping.py:500-514)
__global__ void qam_map_kernel(const float* __restrict__ bits, const float2* __restrict__ points, int m,
                               float2* __restrict__ out, int* __restrict__ idx_out, long long n_sym) {
    extern __shared__ float2 s_pts[];
    for (int i = threadIdx.x; i < (1 << m); i += blockDim.x) s_pts[i] = points[i];
    __syncthreads();
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_sym; s += stride) {
        int v = 0;
        for (int k = 0; k < m; ++k) v = (v << 1) | ((int)bits[s * m + k] & 1);
        out[s] = s_pts[v];
        if (idx_out) idx_out[s] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// log_sigmoid(x) = -softplus(-x) with TensorFlow's softplus branches (threshold = log(eps) + 2)
__device__ __forceinline__ float log1p_pos(float u) {   // u >= 0
    float w = __fadd_rn(1.f, u);
    if (w == 1.f) return u;
    return __fmul_rn(sb_logf(w), __fdiv_rn(u, __fsub_rn(w, 1.f)));
}
__device__ __forceinline__ float softplusf(float x) {
    const float threshold = -13.942385f;   // logf(FLT_EPSILON) + 2
    if (x > -threshold) return x;
    float ex = sb_expf(x);
    if (x < threshold) return ex;
    return log1p_pos(ex);
}
__device__ __forceinline__ float log_sigmoidf(float x) { return -softplusf(-x); }

// One thread per symbol, M = bits per symbol at compile time. Exponents e_j = -|y - c_j|^2 / max(no, tiny) (+ prior
// term) are evaluated once per pass and feed all 2M groups {points with bit i = v} at the same time:
//   pass 1: group maxima (maxlog: done);  pass 2 (app): sum_j exp(e_j - max_group) per group, two groups per packed
//   FP32x2 exp (sb_math2.cuh, bit-identical to sb_expf); LLR_i = logsumexp(bit i = 1) - logsumexp(bit i = 0).
// Per group the operation order is the one of tf.reduce_logsumexp over the points in ascending label order, which is
// what the CPU oracle (oracle/mapping_ref.c) evaluates. The M LLRs of a warp's 32 symbols are staged through shared
// memory so that global stores are contiguous.
template <int M>
__device__ __forceinline__ float demap_exponent(float2 yy, float2 c, float n0, const float* ls1, const float* ls0, int j,
                                                bool with_prior) {
    float dr = __fsub_rn(yy.x, c.x), di = __fsub_rn(yy.y, c.y);
    float a = __fsqrt_rn(__fmaf_rn(dr, dr, __fmul_rn(di, di)));     // |y - c|  (tf.abs)
    float e = __fdiv_rn(-__fmul_rn(a, a), n0);                       // -|.|^2 / no
    if (with_prior) {
        float ps = 0.f;
#pragma unroll
        for (int k = 0; k < M; ++k) ps = __fadd_rn(ps, ((j >> (M - 1 - k)) & 1) ? ls1[k] : ls0[k]);
        e = __fadd_rn(ps, e);
    }
    return e;
}

template <int METHOD, int M>   // METHOD 0 = app, 1 = maxlog
__global__ void __launch_bounds__(128) demap_kernel(const float2* __restrict__ y, const float* __restrict__ no,
                                                    long long no_inner, const float2* __restrict__ points,
                                                    const float* __restrict__ prior, long long prior_inner,
                                                    float* __restrict__ llr, long long n_sym, int hard_out) {
    extern __shared__ float2 s_pts[];
    constexpr int NPTS = 1 << M;
    float* s_out = reinterpret_cast<float*>(s_pts + NPTS);          // [blockDim.x * M] staging for coalesced stores
    for (int i = threadIdx.x; i < NPTS; i += blockDim.x) s_pts[i] = points[i];
    __syncthreads();
    const float tiny = 1.17549435e-38f;   // np.finfo(float32).tiny (mapping.py:653)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n_sym; base += stride) {
        const long long s = base + threadIdx.x;
        float out[M];
        if (s < n_sym) {
            const float2 yy = y[s];
            const float n0 = fmaxf(no[s / no_inner], tiny);
            float ls1[M], ls0[M];
            const bool with_prior = prior != nullptr;
            if (with_prior) {
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    float pk = prior[(s / prior_inner) * M + k];
                    ls1[k] = log_sigmoidf(pk);                       // label bit 1 -> +prior
                    ls0[k] = log_sigmoidf(__fmul_rn(-1.f, pk));
                }
            }
            float mx0[M], mx1[M];
#pragma unroll
            for (int i = 0; i < M; ++i) { mx0[i] = -INFINITY; mx1[i] = -INFINITY; }
#pragma unroll 4
            for (int j = 0; j < NPTS; ++j) {
                const float e = demap_exponent<M>(yy, s_pts[j], n0, ls1, ls0, j, with_prior);
#pragma unroll
                for (int i = 0; i < M; ++i) {                        // label bit i, MSB first (mapping.py:894-907)
                    if ((j >> (M - 1 - i)) & 1) mx1[i] = fmaxf(mx1[i], e);
                    else mx0[i] = fmaxf(mx0[i], e);
                }
            }
            if (METHOD == 1) {
#pragma unroll
                for (int i = 0; i < M; ++i) out[i] = __fsub_rn(mx1[i], mx0[i]);
            } else {
                // tf.reduce_logsumexp: log(sum(exp(x - max))) + max, max replaced by 0 if not finite
                float sm0[M], sm1[M];
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    mx0[i] = (mx0[i] > -INFINITY && mx0[i] < INFINITY) ? mx0[i] : 0.f;
                    mx1[i] = (mx1[i] > -INFINITY && mx1[i] < INFINITY) ? mx1[i] : 0.f;
                    sm0[i] = 0.f; sm1[i] = 0.f;
                }
#pragma unroll 2
                for (int j = 0; j < NPTS; ++j) {
                    const float e = demap_exponent<M>(yy, s_pts[j], n0, ls1, ls0, j, with_prior);
                    float t[M + 1];
#pragma unroll
                    for (int i = 0; i < M; ++i) t[i] = __fsub_rn(e, ((j >> (M - 1 - i)) & 1) ? mx1[i] : mx0[i]);
                    t[M] = 0.f;
#pragma unroll
                    for (int i = 0; i < M; i += 2) {                 // exp of two groups at a time (FFMA2)
                        float2 a = make_float2(fmaxf(t[i], -87.3f), fmaxf(t[i + 1], -87.3f));
                        float2 r = sb_expf2_inrange(a);
                        if (t[i] < -87.3f) r.x = 0.f;                // sb_expf: exact 0 below -87.3
                        if (t[i + 1] < -87.3f) r.y = 0.f;
                        t[i] = r.x;
                        t[i + 1] = r.y;
                    }
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        if ((j >> (M - 1 - i)) & 1) sm1[i] = __fadd_rn(sm1[i], t[i]);
                        else sm0[i] = __fadd_rn(sm0[i], t[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const float2 lg = sb_logf2(make_float2(fmaxf(sm0[i], 1.17549435e-38f), fmaxf(sm1[i], 1.17549435e-38f)));
                    float a1 = __fadd_rn(sm1[i] > 0.f ? lg.y : -INFINITY, mx1[i]);
                    float a0 = __fadd_rn(sm0[i] > 0.f ? lg.x : -INFINITY, mx0[i]);
                    out[i] = __fsub_rn(a1, a0);
                }
            }
#pragma unroll
            for (int i = 0; i < M; ++i) out[i] = hard_out ? (out[i] > 0.f ? 1.f : 0.f) : out[i];   // utils/misc.py:270
        }
        __syncthreads();                                            // previous tile's copy-out has finished
        if (s < n_sym) {
#pragma unroll
            for (int i = 0; i < M; ++i) s_out[threadIdx.x * M + i] = out[i];
        }
        __syncthreads();
        const long long rem = n_sym - base;
        const int cnt = (int)((rem < (long long)blockDim.x ? rem : (long long)blockDim.x) * M);
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) llr[base * M + e] = s_out[e];
    }
}

// Separable constellations (every square QAM of mapping.py:104-117: even label bits select the real PAM level, odd
// label bits the imaginary one): the 2-D sums factor, exp(e_j) = exp(e_re) exp(e_im), and the factor of the other
// dimension cancels in the LLR, so bit i only needs the 2^(M/2) exponents of its own dimension:
//   LLR_(2u+d) = logsumexp_{t: bit u of t = 1} e_d(t) - logsumexp_{t: bit u = 0} e_d(t),  e_d(t) = -(y_d - a_d(t))^2 * (1/no)
// (max instead of logsumexp for maxlog). 2 * 2^(M/2) exponents instead of 2^M and M * 2^(M/2) instead of M * 2^M
// exp() per symbol. Same value as the generic kernel up to fp32 rounding (tests: rtol 1e-4 against the oracle).
template <int METHOD, int H>   // H = M / 2 bits per dimension
__global__ void __launch_bounds__(128) demap_qam_kernel(const float2* __restrict__ y, const float* __restrict__ no,
                                                        long long no_inner, const float* __restrict__ lev_re,
                                                        const float* __restrict__ lev_im, float* __restrict__ llr,
                                                        long long n_sym, int hard_out) {
    constexpr int L = 1 << H, M = 2 * H;
    extern __shared__ float s_out_q[];                              // [blockDim.x * M]
    float lr[L], li[L];
#pragma unroll
    for (int t = 0; t < L; ++t) { lr[t] = lev_re[t]; li[t] = lev_im[t]; }
    const float tiny = 1.17549435e-38f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n_sym; base += stride) {
        const long long s = base + threadIdx.x;
        float out[M];
        if (s < n_sym) {
            const float2 yy = y[s];
            const float inv_n0 = __fdiv_rn(1.0f, fmaxf(no[s / no_inner], tiny));   // one division per symbol
            demap_qam_symbol<METHOD, H>(yy, inv_n0, lr, li, lev_re, lev_im, hard_out, out);
        }
        // stage the warp's 32 x M LLRs through its own shared-memory tile: contiguous global stores, no CTA barrier
        {
            const int lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * 32 * M;
            __syncwarp();                                           // previous tile's copy-out has finished
            if (s < n_sym) {
#pragma unroll
                for (int i = 0; i < M; ++i) s_out_q[wbase + lane * M + i] = out[i];
            }
            __syncwarp();
            const long long wfirst = base + (threadIdx.x & ~31);
            const long long rem = n_sym - wfirst;
            const int cnt = rem <= 0 ? 0 : (int)((rem < 32 ? rem : 32) * M);
            for (int q = lane; q < cnt; q += 32) llr[wfirst * M + q] = s_out_q[wbase + q];
        }
    }
}

template <int METHOD>
void launch_demap_qam(int h, int grid, cudaStream_t st, const float2* y, const float* no, long long no_inner,
                      const float* lev_re, const float* lev_im, float* llr, long long n_sym, int hard_out) {
    const size_t smem = sizeof(float) * 128 * 2 * h;
#define SB_QAM_CASE(HH) case HH: demap_qam_kernel<METHOD, HH><<<grid, 128, smem, st>>>(y, no, no_inner, lev_re, lev_im, llr, n_sym, hard_out); break;
    switch (h) { SB_QAM_CASE(1) SB_QAM_CASE(2) SB_QAM_CASE(3) SB_QAM_CASE(4) SB_QAM_CASE(5) }
#undef SB_QAM_CASE
}

template <int METHOD>
void launch_demap(int m, int grid, size_t smem, cudaStream_t st, const float2* y, const float* no, long long no_inner,
                  const float2* pts, const float* prior, long long prior_inner, float* llr, long long n_sym, int hard_out) {
#define SB_DEMAP_CASE(MM) case MM: demap_kernel<METHOD, MM><<<grid, 128, smem, st>>>(y, no, no_inner, pts, prior, prior_inner, llr, n_sym, hard_out); break;
    switch (m) {
        SB_DEMAP_CASE(1) SB_DEMAP_CASE(2) SB_DEMAP_CASE(3) SB_DEMAP_CASE(4) SB_DEMAP_CASE(5) SB_DEMAP_CASE(6)
        SB_DEMAP_CASE(7) SB_DEMAP_CASE(8) SB_DEMAP_CASE(9) SB_DEMAP_CASE(10) SB_DEMAP_CASE(11) SB_DEMAP_CASE(12)
    }
#undef SB_DEMAP_CASE
}

// ---------------------------------------------------------------------------------------------------------
// y = x + sqrt(no) * n,  n ~ CN(0, 1): re, im ~ N(0, 1/2) (utils/misc.py:46-52, channel/awgn.py:66-78).
// One Philox block (4 uniforms -> 2 Box-Muller pairs) serves two complex samples.
__global__ void awgn_kernel(const float2* x, const float* __restrict__ no, long long no_inner,
                            float2* y, long long n, unsigned long long seed, unsigned long long offset) {
    long long stride = (long long)gridDim.x * blockDim.x;
    const float stddev = 0.70710678118654752f;   // sqrt(var/2), var = 1
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; 2 * p < n; p += stride) {
        uint4 r = philox4x32_10(seed, offset, (unsigned long long)p);
        float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
        long long i0 = 2 * p, i1 = 2 * p + 1;
        float s0 = sqrtf(no[i0 / no_inner]);
        float2 a = x[i0];
        y[i0] = make_float2(a.x + (g0.x * stddev) * s0, a.y + (g0.y * stddev) * s0);
        if (i1 < n) {
            float s1 = sqrtf(no[i1 / no_inner]);
            float2 b = x[i1];
            y[i1] = make_float2(b.x + (g1.x * stddev) * s1, b.y + (g1.y * stddev) * s1);
        }
    }
}

// real-valued variant used for LLR-domain test sources (GaussianPriorSource): out = mean + std * N(0,1)
__global__ void normal_kernel(float* __restrict__ out, long long n, float mean, float stdv, unsigned long long seed,
                              unsigned long long offset) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; 4 * p < n; p += stride) {
        uint4 r = philox4x32_10(seed, offset, (unsigned long long)p);
        float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
        float v[4] = {g0.x, g0.y, g1.x, g1.y};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * p + k < n) out[4 * p + k] = mean + stdv * v[k];
    }
}

// uniform variant: out = lo + (hi - lo) * u, u in [0, 1) with 24 random bits (TDL Doppler / angle / phase draws)
__global__ void uniform_kernel(float* __restrict__ out, long long n, float lo, float hi, unsigned long long seed,
                               unsigned long long offset) {
    long long stride = (long long)gridDim.x * blockDim.x;
    const float w = hi - lo;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; 4 * p < n; p += stride) {
        uint4 r = philox4x32_10(seed, offset, (unsigned long long)p);
        unsigned v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * p + k < n) out[4 * p + k] = lo + w * ((float)(v[k] >> 8) * 5.9604644775390625e-08f);   // 2^-24
    }
}

// ---------------------------------------------------------------------------------------------------------
// counters[0] += #(b != b_hat), counters[1] += #rows with any difference, counters[2] += B*k, counters[3] += B.
// One warp per row; block-level reduction, one atomic per block and counter.
__global__ void count_errors_kernel(const float* __restrict__ b, const float* __restrict__ bh, long long rows, int k,
                                    unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long s_bit[32], s_blk[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned long long bit_e = 0, blk_e = 0;
    for (long long r = (long long)blockIdx.x * nwarps + warp; r < rows; r += (long long)gridDim.x * nwarps) {
        const float* pb = b + r * k;
        const float* ph = bh + r * k;
        unsigned cnt = 0;
        for (int i = lane; i < k; i += 32) cnt += (pb[i] != ph[i]);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        bit_e += cnt;
        blk_e += (cnt != 0);
    }
    if (lane == 0) { s_bit[warp] = bit_e; s_blk[warp] = blk_e; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tb = 0, tk = 0;
        for (int w = 0; w < nwarps; ++w) { tb += s_bit[w]; tk += s_blk[w]; }
        if (tb) atomicAdd(&counters[0], tb);
        if (tk) atomicAdd(&counters[1], tk);
        if (blockIdx.x == 0) {
            atomicAdd(&counters[2], (unsigned long long)rows * (unsigned long long)k);
            atomicAdd(&counters[3], (unsigned long long)rows);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// CRCEncoder.call (fec/crc.py:175-215): the reference multiplies the bit row with a dense [k, L] generator matrix and
// reduces mod 2; the CRC is linear, so the parity word is the XOR of the (bit-packed) matrix rows selected by the set
// bits. One warp per row: lanes stride over the k bits, XOR-reduce, then write [bits | parity] (MSB = first parity bit).
__global__ void crc_encode_kernel(const float* __restrict__ bits, const unsigned* __restrict__ gtab, int k, int L,
                                  float* __restrict__ out, long long rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (long long r = (long long)blockIdx.x * nwarps + warp; r < rows; r += (long long)gridDim.x * nwarps) {
        const float* b = bits + r * k;
        float* o = out + r * (long long)(k + L);
        unsigned acc = 0;
        for (int i = lane; i < k; i += 32) {
            float v = b[i];
            o[i] = v;
            if (((int)v) & 1) acc ^= gtab[i];
        }
        acc = __reduce_xor_sync(0xffffffffu, acc);
        if (lane < L) o[k + lane] = (float)((acc >> (L - 1 - lane)) & 1u);
    }
}

// CRCDecoder.call (fec/crc.py:300-327): the whole word [info | parity] (n bits) is run through the encoder again and the
// check passes iff the new parity is all zero. One warp per row; also copies the n - L information bits.
__global__ void crc_check_kernel(const float* __restrict__ x, const unsigned* __restrict__ gtab, int n, int L,
                                 float* __restrict__ info, unsigned char* __restrict__ valid, long long rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int k = n - L;
    for (long long r = (long long)blockIdx.x * nwarps + warp; r < rows; r += (long long)gridDim.x * nwarps) {
        const float* b = x + r * (long long)n;
        unsigned acc = 0;
        for (int i = lane; i < n; i += 32) {
            float v = b[i];
            if (info && i < k) info[r * (long long)k + i] = v;
            if (((int)v) & 1) acc ^= gtab[i];
        }
        acc = __reduce_xor_sync(0xffffffffu, acc);
        if (lane == 0) valid[r] = acc == 0u ? 1 : 0;
    }
}

// TB5GScrambler.call (fec/scrambling.py:442-468): binary: |x - c|; soft values: x * (1 - 2c). seq has seq_rows rows
// (one per stream) of length n; row r of x uses seq row (r mod seq_rows).
__global__ void scramble_kernel(const float* __restrict__ x, const float* __restrict__ seq, int binary,
                                float* __restrict__ out, long long rows, int n, int seq_rows) {
    const long long total = rows * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % n);
        long long r = i / n;
        float c = seq[(r % seq_rows) * n + j];
        float v = x[i];
        out[i] = binary ? fabsf(v - c) : v * (-2.f * c + 1.f);
    }
}

inline int grid_for(long long work_items, int threads) {
    int sms = sb_num_sms();
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;                       // grid-stride beyond 8 CTAs per SM
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" int sb_binary_source(float* d_out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    if (n == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_out && n >= 0, "sb_binary_source: bad arguments");
    if (n == 0) return SB_OK;
    binary_source_kernel<<<grid_for((n + 127) / 128, 128), 128, 0, (cudaStream_t)stream>>>(d_out, n, seed, offset);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_qam_map(const float* d_bits, const float* d_points, int32_t m, float* d_out, int32_t* d_idx_out,
                          int64_t n_sym, void* stream) {
    if (n_sym == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_bits && d_points && d_out && m >= 1 && m <= 12 && n_sym >= 0, "sb_qam_map: bad arguments");
    if (n_sym == 0) return SB_OK;
    size_t smem = sizeof(float2) << m;
    qam_map_kernel<<<grid_for(n_sym, 256), 256, smem, (cudaStream_t)stream>>>(
        d_bits, (const float2*)d_points, m, (float2*)d_out, d_idx_out, n_sym);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_demap(const float* d_y, const float* d_no, int64_t no_inner, const float* d_points, int32_t m,
                        int32_t method, const float* d_prior, int64_t prior_inner, float* d_llr, int64_t n_sym,
                        int32_t hard_out, void* stream) {
    if (n_sym == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_y && d_no && d_points && d_llr && m >= 1 && m <= 12 && n_sym >= 0 && no_inner >= 1,
                 "sb_demap: bad arguments");
    SB_CHECK_ARG(method == 0 || method == 1, "sb_demap: method must be 0 (app) or 1 (maxlog)");
    SB_CHECK_ARG(!d_prior || prior_inner >= 1, "sb_demap: prior_inner must be >= 1");
    if (n_sym == 0) return SB_OK;
    size_t smem = (sizeof(float2) << m) + sizeof(float) * 128 * m;
    int grid = grid_for(n_sym, 128);
    if (method == 0)
        launch_demap<0>(m, grid, smem, (cudaStream_t)stream, (const float2*)d_y, d_no, no_inner, (const float2*)d_points,
                        d_prior, prior_inner, d_llr, n_sym, hard_out);
    else
        launch_demap<1>(m, grid, smem, (cudaStream_t)stream, (const float2*)d_y, d_no, no_inner, (const float2*)d_points,
                        d_prior, prior_inner, d_llr, n_sym, hard_out);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_demap_qam(const float* d_y, const float* d_no, int64_t no_inner, const float* d_levels_re,
                            const float* d_levels_im, int32_t m, int32_t method, float* d_llr, int64_t n_sym,
                            int32_t hard_out, void* stream) {
    if (n_sym == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_y && d_no && d_levels_re && d_levels_im && d_llr && m >= 2 && m <= 10 && m % 2 == 0 && no_inner >= 1 &&
                     (method == 0 || method == 1), "sb_demap_qam: bad arguments (m even, 2..10; method 0 | 1)");
    int grid = grid_for(n_sym, 128);
    if (method == 0)
        launch_demap_qam<0>(m / 2, grid, (cudaStream_t)stream, (const float2*)d_y, d_no, no_inner, d_levels_re, d_levels_im,
                            d_llr, n_sym, hard_out);
    else
        launch_demap_qam<1>(m / 2, grid, (cudaStream_t)stream, (const float2*)d_y, d_no, no_inner, d_levels_re, d_levels_im,
                            d_llr, n_sym, hard_out);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_awgn(const float* d_x, const float* d_no, int64_t no_inner, float* d_y, int64_t n, uint64_t seed,
                       uint64_t offset, void* stream) {
    if (n == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_x && d_no && d_y && n >= 0 && no_inner >= 1, "sb_awgn: bad arguments");
    if (n == 0) return SB_OK;
    awgn_kernel<<<grid_for((n + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)d_x, d_no, no_inner,
                                                                              (float2*)d_y, n, seed, offset);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_normal(float* d_out, int64_t n, float mean, float stddev, uint64_t seed, uint64_t offset, void* stream) {
    if (n == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_out && n >= 0, "sb_normal: bad arguments");
    if (n == 0) return SB_OK;
    normal_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(d_out, n, mean, stddev, seed, offset);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_uniform(float* d_out, int64_t n, float lo, float hi, uint64_t seed, uint64_t offset, void* stream) {
    if (n == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_out && n >= 0 && hi >= lo, "sb_uniform: bad arguments");
    uniform_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(d_out, n, lo, hi, seed, offset);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_count_errors(const float* d_b, const float* d_b_hat, int64_t rows, int32_t k, int64_t* d_counters,
                               void* stream) {
    if (rows == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_b && d_b_hat && d_counters && rows >= 0 && k >= 1, "sb_count_errors: bad arguments");
    if (rows == 0) return SB_OK;
    count_errors_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        d_b, d_b_hat, rows, k, (unsigned long long*)d_counters);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_crc_encode(const float* d_bits, const uint32_t* d_gen_rows, int32_t k, int32_t crc_length, float* d_out,
                             int64_t rows, void* stream) {
    if (rows == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_bits && d_gen_rows && d_out && k >= 1 && crc_length >= 1 && crc_length <= 32 && rows >= 0,
                 "sb_crc_encode: bad arguments");
    if (rows == 0) return SB_OK;
    crc_encode_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(d_bits, d_gen_rows, k, crc_length, d_out, rows);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_crc_check(const float* d_x, const uint32_t* d_gen_rows, int32_t n, int32_t crc_length, float* d_info,
                            uint8_t* d_valid, int64_t rows, void* stream) {
    if (rows == 0) return SB_OK;
    SB_CHECK_ARG(d_x && d_gen_rows && d_valid && crc_length >= 1 && crc_length <= 32 && n >= crc_length && rows >= 0,
                 "sb_crc_check: bad arguments");
    crc_check_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_gen_rows, n, crc_length, d_info, d_valid, rows);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_scramble(const float* d_x, const float* d_seq, int32_t binary, float* d_out, int64_t rows, int32_t n,
                           int32_t seq_rows, void* stream) {
    if (rows == 0) return SB_OK;                         // empty batch: nothing to do, pointers may be null
    SB_CHECK_ARG(d_x && d_seq && d_out && rows >= 0 && n >= 1 && seq_rows >= 1, "sb_scramble: bad arguments");
    if (rows == 0) return SB_OK;
    scramble_kernel<<<grid_for(rows * n, 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_seq, binary, d_out, rows, n, seq_rows);
    SB_LAUNCH_CHECK();
    return SB_OK;
}
