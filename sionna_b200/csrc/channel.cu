// channel.cu -- on-device channel generation helpers for sm_100a, second part: everything
// the CIR -> channel conversion needs so that no step of it runs as an eager tensor expression. Replaces (paths under
// /root/reference/src/sionna/phy/):
//   sb_phase_table     exp(-j 2 pi f tau) of cir_to_ofdm_channel   channel/utils.py:232-244
//                      sinc(l - tau W)    of cir_to_time_channel   channel/utils.py:318-338
//   sb_cir_gram        (no counterpart: Gram matrix of the table, used to normalise without a second pass over h)
//   sb_cir_link_scale  normalisation factor of channel/utils.py:246-251 (OFDM) and :341-348 (time)
//   sb_cir_apply       h = sum_p a_p e_p (channel/utils.py:240-244, 336-338), per-link tables and scaling folded in
//   sb_spatial_corr    TDL._apply_correlation (channel/tr38901/tdl.py:466-490): v' = L v per (batch, path, time step)
//
// Normalisation without touching h twice. The reference computes c = sqrt(mean |h|^2) over (rx ant, tx ant, time,
// frequency) per link from the finished tensor and divides. Because h[f] = sum_p a_p e[p, f],
//     sum_f |h[f]|^2 = sum_{p,q} a_p conj(a_q) G[p, q],   G[p, q] = sum_f e[p, f] conj(e[q, f]),
// so the link energy follows from the P path gains and the P x P Gram matrix of the table (P <= 24 for every TDL
// model): sb_cir_link_scale evaluates that quadratic form (P^2 MACs per (antenna pair, time step) instead of reading
// F * 8 bytes) and sb_cir_apply writes the already scaled h exactly once. Every reduction runs in a fixed order
// (no atomics): results are reproducible bit for bit from run to run.
#include <algorithm>
#include "sb_common.h"

namespace {

__device__ __forceinline__ float2 cmul_(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

inline int grid_cap(long long blocks) {
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)sb_num_sms() * 16));
}

// e[tab, p, j]: mode 0: exp(-j 2 pi x_j tau[tab, p]); mode 1: sinc(x_j - tau[tab, p] * scale) (+ 0 j).
// The phase is reduced in double precision (f tau reaches a few turns) and evaluated with sincospi.
__global__ void phase_table_kernel(const float* __restrict__ tau, const float* __restrict__ x, float2* __restrict__ e,
                                   long long n_tab, int P, int F, float scale, int mode) {
    const long long total = n_tab * P * (long long)F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % F);
        const long long tp = i / F;
        const double t = (double)tau[tp];
        if (mode == 0) {
            double turns = (double)x[j] * t;                     // f * tau
            turns -= rint(turns);
            double s, c;
            sincospi(-2.0 * turns, &s, &c);
            e[i] = make_float2((float)c, (float)s);
        } else {
            const double u = (double)x[j] - t * (double)scale;   // l - tau W
            double v = 1.0;
            if (u != 0.0) v = sinpi(u) / (3.141592653589793 * u);
            e[i] = make_float2((float)v, 0.f);
        }
    }
}

// G[tab, p, q] = sum_j e[tab, p, j] conj(e[tab, q, j]); one warp per (tab, p, q), lanes stride j, fixed shuffle tree.
__global__ void cir_gram_kernel(const float2* __restrict__ e, float2* __restrict__ g, long long n_tab, int P, int F) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = n_tab * P * (long long)P;
    for (long long w = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < total; w += warps) {
        const int q = (int)(w % P);
        const int p = (int)((w / P) % P);
        const long long tab = w / ((long long)P * P);
        const float2* ep = e + (tab * P + p) * (long long)F;
        const float2* eq = e + (tab * P + q) * (long long)F;
        float re = 0.f, im = 0.f;
        for (int j = lane; j < F; j += 32) {
            const float2 a = ep[j], b = eq[j];
            re += a.x * b.x + a.y * b.y;
            im += a.y * b.x - a.x * b.y;
        }
        for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_down_sync(0xffffffffu, re, o);
            im += __shfl_down_sync(0xffffffffu, im, o);
        }
        if (lane == 0) g[w] = make_float2(re, im);
    }
}

// Tap layout a[b, rx, ra, tx, ta, p, t]; a "link" is (b, rx, tx), its rows are the RA * TA antenna pairs.
struct CirDims { long long B; int RX, RA, TX, TA, P, T; };

// scale[link] = 1 / sqrt( (sum over the link's antenna pairs and time steps of a^H G a) / (RA * TA * T * denom) ), 0 if
// the energy is 0. One CTA per link: thread k takes (pair, t) items k, k + blockDim, ... in order; fixed tree afterwards.
__global__ void cir_link_scale_kernel(const float2* __restrict__ a, const float2* __restrict__ g, long long g_link_stride,
                                      float* __restrict__ scale, CirDims d, float denom) {
    extern __shared__ float2 s_g[];                              // P x P
    __shared__ float s_part[32];
    const long long links = d.B * d.RX * d.TX;
    for (long long link = blockIdx.x; link < links; link += gridDim.x) {
        const int tx = (int)(link % d.TX);
        const int rx = (int)((link / d.TX) % d.RX);
        const long long b = link / ((long long)d.TX * d.RX);
        const float2* gp = g + link * g_link_stride;
        __syncthreads();
        for (int i = threadIdx.x; i < d.P * d.P; i += blockDim.x) s_g[i] = gp[i];
        __syncthreads();
        const int items = d.RA * d.TA * d.T;
        float acc = 0.f;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int t = it % d.T;
            const int pair = it / d.T;
            const int ta = pair % d.TA, ra = pair / d.TA;
            const long long row = (((b * d.RX + rx) * d.RA + ra) * d.TX + tx) * d.TA + ta;
            const float2* ap = a + row * (long long)d.P * d.T + t;
            float e = 0.f;
            for (int p = 0; p < d.P; ++p) {
                const float2 x = ap[(size_t)p * d.T];
                // diagonal term + 2 Re of the strictly lower triangle (G is Hermitian)
                e += (x.x * x.x + x.y * x.y) * s_g[p * d.P + p].x;
                float2 s = make_float2(0.f, 0.f);
                for (int q = 0; q < p; ++q) {
                    const float2 y = ap[(size_t)q * d.T];
                    const float2 gq = s_g[p * d.P + q];             // sum_f e_p conj(e_q)
                    // terms (p, q) and (q, p) together: 2 Re( a_p conj(a_q) G[p, q] )
                    const float2 cy = cmul_(x, make_float2(y.x, -y.y));
                    s.x += cy.x * gq.x - cy.y * gq.y;
                }
                e += 2.f * s.x;
            }
            acc += e;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
            const float mean = tot / ((float)items * denom);
            scale[link] = mean > 0.f ? 1.0f / sqrtf(mean) : 0.f;
        }
    }
}

// h[row, t, j] = scale[link(row)] * sum_p a[row, p, t] e[tab(row), p, j]
// A CTA owns a TILE_T x TILE_F tile of one antenna-pair row: the P x TILE_T taps and the P x TILE_F table entries are
// staged in shared memory once, every thread accumulates RT x RJ outputs in registers (per path: RT + RJ shared loads
// feed RT * RJ complex multiply-adds), so the kernel is bound by the FP32 pipe and by the single write of h instead of
// by 2 * P loads per output. Two shapes: wide rows (OFDM: F = 76 ... 4096 subcarriers, few time steps) and narrow rows
// (time channel: l_tot ~ 20 taps, thousands of time steps).
template <int TILE_T, int RT, int TILE_F, int RJ>
__global__ void __launch_bounds__((TILE_T / RT) * (TILE_F / RJ))
cir_apply_kernel(const float2* __restrict__ a, const float2* __restrict__ e, long long e_link_stride,
                 const float* __restrict__ scale, float2* __restrict__ h, CirDims d, int F, int tiles_t, int tiles_f) {
    extern __shared__ float2 s_cir[];
    constexpr int NTJ = TILE_F / RJ;                             // threads along the columns
    float2* s_a = s_cir;                                          // [P][TILE_T]
    float2* s_e = s_cir + (size_t)d.P * TILE_T;                   // [P][TILE_F]
    const int tj = threadIdx.x % NTJ, tt = threadIdx.x / NTJ;
    const long long R = d.B * d.RX * d.RA * d.TX * d.TA;
    const long long n_tiles = R * tiles_t * tiles_f;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int ft = (int)(tile % tiles_f);
        const int tq = (int)((tile / tiles_f) % tiles_t);
        const long long r = tile / ((long long)tiles_f * tiles_t);
        const int tx = (int)((r / d.TA) % d.TX);
        const long long link = (r / ((long long)d.TA * d.TX * d.RA)) * d.TX + tx;   // (b * RX + rx) * TX + tx
        const float sc = scale ? scale[link] : 1.f;
        const float2* ap = a + r * (long long)d.P * d.T;
        const float2* ep = e + link * e_link_stride;
        const int t0 = tq * TILE_T, j0 = ft * TILE_F;
        __syncthreads();
        for (int i = threadIdx.x; i < d.P * TILE_T; i += blockDim.x) {
            const int p = i / TILE_T, t = t0 + i % TILE_T;
            s_a[i] = t < d.T ? ap[(size_t)p * d.T + t] : make_float2(0.f, 0.f);
        }
        for (int i = threadIdx.x; i < d.P * TILE_F; i += blockDim.x) {
            const int p = i / TILE_F, j = j0 + i % TILE_F;
            s_e[i] = j < F ? ep[(size_t)p * F + j] : make_float2(0.f, 0.f);
        }
        __syncthreads();
        float2 acc[RT][RJ];
#pragma unroll
        for (int u = 0; u < RT; ++u)
#pragma unroll
            for (int v = 0; v < RJ; ++v) acc[u][v] = make_float2(0.f, 0.f);
        for (int p = 0; p < d.P; ++p) {
            float2 av[RT], ev[RJ];
#pragma unroll
            for (int u = 0; u < RT; ++u) av[u] = s_a[p * TILE_T + tt * RT + u];
#pragma unroll
            for (int v = 0; v < RJ; ++v) ev[v] = s_e[p * TILE_F + tj + v * NTJ];
#pragma unroll
            for (int u = 0; u < RT; ++u)
#pragma unroll
                for (int v = 0; v < RJ; ++v) {
                    // explicit FMAs (the library is built with -fmad=false for the bit-exact decoder kernels)
                    acc[u][v].x = fmaf(av[u].x, ev[v].x, fmaf(-av[u].y, ev[v].y, acc[u][v].x));
                    acc[u][v].y = fmaf(av[u].x, ev[v].y, fmaf(av[u].y, ev[v].x, acc[u][v].y));
                }
        }
#pragma unroll
        for (int u = 0; u < RT; ++u) {
            const int t = t0 + tt * RT + u;
            if (t >= d.T) continue;
            float2* hp = h + (r * d.T + t) * (long long)F;
#pragma unroll
            for (int v = 0; v < RJ; ++v) {
                const int j = j0 + tj + v * NTJ;
                if (j < F) hp[j] = make_float2(acc[u][v].x * sc, acc[u][v].y * sc);
            }
        }
    }
}

template <int TILE_T, int RT, int TILE_F, int RJ>
int launch_cir_apply(const float2* a, const float2* e, long long e_link_stride, const float* scale, float2* h, const CirDims& d,
                     int F, cudaStream_t stream) {
    auto kern = cir_apply_kernel<TILE_T, RT, TILE_F, RJ>;
    const size_t smem = sizeof(float2) * (size_t)d.P * (TILE_T + TILE_F);
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_t = (d.T + TILE_T - 1) / TILE_T, tiles_f = (F + TILE_F - 1) / TILE_F;
    const long long R = d.B * d.RX * d.RA * d.TX * d.TA;
    kern<<<grid_cap(R * tiles_t * tiles_f), (TILE_T / RT) * (TILE_F / RJ), smem, stream>>>(a, e, e_link_stride, scale, h, d, F,
                                                                                         tiles_t, tiles_f);
    return SB_OK;
}

// out[b, i, c] = sum_j L[i, j] in[b, j, c]: i, j over the n antenna pairs (rx ant major), c over the (path, time) columns.
__global__ void spatial_corr_kernel(const float2* __restrict__ in, const float2* __restrict__ l, float2* __restrict__ out,
                                    long long B, int n, long long cols) {
    extern __shared__ float2 s_l[];                              // n x n
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) s_l[i] = l[i];
    __syncthreads();
    const long long total = B * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / cols, c = i - b * cols;
        const float2* ip = in + b * n * cols + c;
        float2* op = out + b * n * cols + c;
        for (int r = 0; r < n; ++r) {
            float2 acc = make_float2(0.f, 0.f);
            for (int j = 0; j <= r; ++j) {                         // L is lower triangular (Cholesky factor)
                const float2 v = cmul_(s_l[r * n + j], ip[(size_t)j * cols]);
                acc.x += v.x;
                acc.y += v.y;
            }
            op[(size_t)r * cols] = acc;
        }
    }
}

}  // namespace

extern "C" int sb_phase_table(const float* d_tau, const float* d_x, float scale, int32_t mode, float* d_e, int64_t n_tab,
                              int32_t num_paths, int32_t num_cols, void* stream) {
    if (n_tab == 0) return SB_OK;
    SB_CHECK_ARG(d_tau && d_x && d_e && n_tab > 0 && num_paths > 0 && num_cols > 0 && (mode == 0 || mode == 1),
                 "sb_phase_table: bad arguments");
    const long long total = n_tab * num_paths * (long long)num_cols;
    phase_table_kernel<<<grid_cap((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_tau, d_x, (float2*)d_e, n_tab,
                                                                                       num_paths, num_cols, scale, mode);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_cir_gram(const float* d_e, float* d_g, int64_t n_tab, int32_t num_paths, int32_t num_cols, void* stream) {
    if (n_tab == 0) return SB_OK;
    SB_CHECK_ARG(d_e && d_g && n_tab > 0 && num_paths > 0 && num_cols > 0, "sb_cir_gram: bad arguments");
    const long long warps = n_tab * num_paths * (long long)num_paths;
    cir_gram_kernel<<<grid_cap((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const float2*)d_e, (float2*)d_g, n_tab,
                                                                                num_paths, num_cols);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

static int check_dims(int64_t batch, int32_t rx, int32_t ra, int32_t tx, int32_t ta, int32_t p, int32_t t) {
    return batch > 0 && rx > 0 && ra > 0 && tx > 0 && ta > 0 && p > 0 && t > 0;
}

extern "C" int sb_cir_link_scale(const float* d_a, const float* d_g, int64_t g_link_stride, float* d_scale, int64_t batch,
                                 int32_t num_rx, int32_t num_rx_ant, int32_t num_tx, int32_t num_tx_ant, int32_t num_paths,
                                 int32_t num_time_steps, float denom, void* stream) {
    if (batch == 0) return SB_OK;
    SB_CHECK_ARG(d_a && d_g && d_scale && check_dims(batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_paths, num_time_steps) &&
                     g_link_stride >= 0 && denom > 0.f && num_paths <= 64, "sb_cir_link_scale: bad arguments");
    CirDims d{batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_paths, num_time_steps};
    const long long links = batch * num_rx * (long long)num_tx;
    const size_t smem = sizeof(float2) * (size_t)num_paths * num_paths;
    cir_link_scale_kernel<<<grid_cap(links), 256, smem, (cudaStream_t)stream>>>((const float2*)d_a, (const float2*)d_g,
                                                                                g_link_stride, d_scale, d, denom);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_cir_apply(const float* d_a, const float* d_e, int64_t e_link_stride, const float* d_scale, float* d_h,
                            int64_t batch, int32_t num_rx, int32_t num_rx_ant, int32_t num_tx, int32_t num_tx_ant,
                            int32_t num_paths, int32_t num_time_steps, int32_t num_cols, void* stream) {
    if (batch == 0) return SB_OK;
    SB_CHECK_ARG(d_a && d_e && d_h && check_dims(batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_paths, num_time_steps) &&
                     num_cols > 0 && e_link_stride >= 0, "sb_cir_apply: bad arguments");
    CirDims d{batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_paths, num_time_steps};
    SB_CHECK_ARG(num_paths <= 96, "sb_cir_apply: more than 96 paths are not supported");
    int rc;
    if (num_cols > 48)          // OFDM-like: wide rows
        rc = launch_cir_apply<16, 2, 256, 8>((const float2*)d_a, (const float2*)d_e, e_link_stride, d_scale, (float2*)d_h, d,
                                             num_cols, (cudaStream_t)stream);
    else                        // time-channel-like: a few taps, many time steps
        rc = launch_cir_apply<128, 8, 32, 2>((const float2*)d_a, (const float2*)d_e, e_link_stride, d_scale, (float2*)d_h, d,
                                             num_cols, (cudaStream_t)stream);
    if (rc) return rc;
    SB_LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int sb_spatial_corr(const float* d_in, const float* d_l, float* d_out, int64_t batch, int32_t n, int64_t cols,
                               void* stream) {
    if (batch == 0) return SB_OK;
    SB_CHECK_ARG(d_in && d_l && d_out && d_in != d_out && batch > 0 && n > 0 && n <= 128 && cols > 0,
                 "sb_spatial_corr: bad arguments (n <= 128, out of place)");
    const long long total = batch * cols;
    const size_t smem = sizeof(float2) * (size_t)n * n;
    if (smem > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(spatial_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spatial_corr_kernel<<<grid_cap((total + 127) / 128), 128, smem, (cudaStream_t)stream>>>((const float2*)d_in, (const float2*)d_l,
                                                                                          (float2*)d_out, batch, n, cols);
    SB_LAUNCH_CHECK();
    return SB_OK;
}
