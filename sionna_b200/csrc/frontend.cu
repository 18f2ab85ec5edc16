// frontend.cu -- fused OFDM receive front-end for sm_100a: from the received resource grid
// to bit LLRs in ONE launch. Replaces the chain (paths under /root/reference/src/sionna/phy/)
//   RemoveNulledSubcarriers                     ofdm/resource_grid.py:551
//   LSChannelEstimator / PUSCHLSChannelEstimator ofdm/channel_estimation.py:138-173, 257-285, nr/pusch_channel_estimation.py:117-169
//   Nearest-neighbour / linear interpolation    ofdm/channel_estimation.py:364-435, 479-734
//   OFDMEqualizer glue + lmmse_equalizer        ofdm/equalization.py:126-275, mimo/equalization.py:101-233
//   Demapper (square QAM, app / maxlog)         mapping.py:664-691, 927-967
// for receivers without interfering streams (diagonal noise-plus-estimation-error covariance) and up to 4 streams.
//
// Why this fuses cleanly. LS estimation (y_p / p), CDM de-spreading, nearest-neighbour and (time-averaged) linear
// interpolation are all LINEAR in the received pilot symbols, and the error variance they propagate is the noise power
// times a constant. So for stream q and resource element r
//     h_hat[ant, q](r) = sum_{i < NT} W[q, r, i] * y[ant, P[q, r, i]],      err_var[ant, q](r) = no[ant] * E[q, r]
// with host-built tables (P: pilot RE of the full grid, W: complex weight = interpolation weight / pilot, E >= 0). The
// kernel never materialises h_hat (512 B per RE for 4 x 16) or err_var: a thread owns one RE, walks the antennas once
// (y coalesced over the subcarriers; the pilot REs it gathers are shared by the whole frame and stay in L1), builds
// B = H_w^H H_w and z = H_w^H y_w in registers, solves (lmmse_diag.cuh), demaps (demap_qam.cuh) and writes the LLRs in
// the layout the decoder reads. HBM traffic per RE: y (8 B per antenna) + LLRs - against 692 B of SURVEY 8(d).
#include <algorithm>
#include "sb_common.h"
#include "lmmse_diag.cuh"
#include "demap_qam.cuh"

namespace {

struct FrontParams {
    const float2* y;          // [B, RX, ANT, GRID] full grid (OFDM symbols x fft_size)
    const float* no;          // addressed with no_stride over (b, rx, ant)
    long long no_stride[3];
    const int* des;           // [RX, K] tx-stream index of every desired stream
    const int* out_ts;        // [RX, K] output row (tx * streams_per_tx + stream) after the stream re-ordering
    const int* data_pos;      // [TXS, SF] index among that stream's data symbols or -1
    const int* re_full;       // [SF] position of the effective RE in the full grid
    const int* t_idx;         // [TXS, SF, NT] pilot position in the full grid, -1 = unused term
    const float2* t_w;        // [TXS, SF, NT]
    const float* e_sum;       // [SF] sum over ALL streams of the (floored) error-variance factor
    const float* lev_re;      // [2^H] real PAM levels by label
    const float* lev_im;
    float* llr;               // [B, TXS, ND * 2H]
    float2* xh;               // optional [B, TXS, ND]
    float* ne;                // optional [B, TXS, ND]
    long long B;
    int RX, ANT, TXS, SF, ND, NT, GRID, method, hard_out;
};

constexpr int kAntChunk = 4;

template <int K, int H>
__global__ void __launch_bounds__(128) ofdm_frontend_kernel(const FrontParams p) {
    using namespace sb_lmmse;
    constexpr int L = 1 << H, M = 2 * H;
    float lr[L], li[L];
#pragma unroll
    for (int t = 0; t < L; ++t) { lr[t] = p.lev_re[t]; li[t] = p.lev_im[t]; }
    const long long SF = p.SF;
    const long long total = p.B * p.RX * SF;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int re = (int)(i % SF);
        const int rx = (int)((i / SF) % p.RX);
        const long long b = i / (SF * p.RX);
        int ts[K], dp[K], des[K];
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            ts[k] = p.out_ts[rx * K + k];
            des[k] = p.des[rx * K + k];
            dp[k] = p.data_pos[(size_t)ts[k] * SF + re];
            any = any || dp[k] >= 0;
        }
        if (!any) continue;
        const int rf = p.re_full[re];
        const float es = p.e_sum[re];
        float2 Bm[K * (K + 1) / 2], z[K];
#pragma unroll
        for (int e = 0; e < K * (K + 1) / 2; ++e) Bm[e] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = make_float2(0.f, 0.f);
        for (int m0 = 0; m0 < p.ANT; m0 += kAntChunk) {
            const float2* yp[kAntChunk];
            float w[kAntChunk];
            float2 h[kAntChunk][K];
#pragma unroll
            for (int c = 0; c < kAntChunk; ++c) {
                const int m = min(m0 + c, p.ANT - 1);                        // tail lanes repeat the last antenna (w = 0)
                yp[c] = p.y + ((b * p.RX + rx) * p.ANT + m) * (long long)p.GRID;
                const float nn = p.no[b * p.no_stride[0] + rx * p.no_stride[1] + m * p.no_stride[2]];
                // whitening by 1 / sqrt(no + sum_q err_var_q), err_var_q = no * E_q  (ofdm/equalization.py:205-218)
                w[c] = (m0 + c < p.ANT) ? 1.0f / sqrtf(nn + nn * es) : 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) h[c][k] = make_float2(0.f, 0.f);
            }
            // channel estimates of this antenna chunk: the table words are read once per chunk
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const long long tb = ((long long)des[k] * SF + re) * p.NT;
                for (int t = 0; t < p.NT; ++t) {
                    const int pi = p.t_idx[tb + t];
                    if (pi < 0) break;
                    const float2 wt = p.t_w[tb + t];
#pragma unroll
                    for (int c = 0; c < kAntChunk; ++c) {
                        const float2 yv = yp[c][pi];
                        h[c][k].x = fmaf(wt.x, yv.x, fmaf(-wt.y, yv.y, h[c][k].x));
                        h[c][k].y = fmaf(wt.x, yv.y, fmaf(wt.y, yv.x, h[c][k].y));
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kAntChunk; ++c) {
                float2 yw = yp[c][rf];
                yw = make_float2(yw.x * w[c], yw.y * w[c]);
                float2 hw[K];
#pragma unroll
                for (int k = 0; k < K; ++k) hw[k] = make_float2(h[c][k].x * w[c], h[c][k].y * w[c]);
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    z[a] = cadd(z[a], cmulc(yw, hw[a]));                       // conj(H_w[m, a]) * y_w[m]
#pragma unroll
                    for (int q = 0; q <= a; ++q) Bm[a * (a + 1) / 2 + q] = cadd(Bm[a * (a + 1) / 2 + q], cmulc(hw[q], hw[a]));
                }
            }
        }
        float2 xo[K];
        float no_e[K];
        lmmse_diag_solve<K>(Bm, z, xo, no_e);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dp[k] < 0) continue;
            const long long o = (b * p.TXS + ts[k]) * (long long)p.ND + dp[k];
            if (p.xh) { p.xh[o] = xo[k]; p.ne[o] = no_e[k]; }
            if (p.llr) {
                const float inv_n0 = __fdiv_rn(1.0f, fmaxf(no_e[k], 1.17549435e-38f));   // mapping.py:653, 672-681
                float out[M];
                if (p.method == 1) demap_qam_symbol<1, H>(xo[k], inv_n0, lr, li, p.hard_out, out);
                else demap_qam_symbol<0, H>(xo[k], inv_n0, lr, li, p.hard_out, out);
                float* lp = p.llr + o * M;
#pragma unroll
                for (int j = 0; j < M; ++j) lp[j] = out[j];
            }
        }
    }
}

template <int K>
int launch_front(const FrontParams& p, int h, long long total, cudaStream_t st) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 127) / 128, (long long)sb_num_sms() * 16));
#define SB_FRONT_CASE(HH) case HH: ofdm_frontend_kernel<K, HH><<<grid, 128, 0, st>>>(p); break;
    switch (h) { SB_FRONT_CASE(1) SB_FRONT_CASE(2) SB_FRONT_CASE(3) SB_FRONT_CASE(4) SB_FRONT_CASE(5) default: return SB_EUNSUPPORTED; }
#undef SB_FRONT_CASE
    return SB_OK;
}

}  // namespace

extern "C" int sb_ofdm_frontend(const float* d_y, const float* d_no, const int64_t* h_no_stride, const int32_t* d_desired,
                                const int32_t* d_out_stream, const int32_t* d_data_pos, const int32_t* d_re_full,
                                const int32_t* d_t_idx, const float* d_t_w, const float* d_e_sum, const float* d_lev_re,
                                const float* d_lev_im, float* d_llr, float* d_x_hat, float* d_no_eff, int64_t batch,
                                int32_t num_rx, int32_t num_rx_ant, int32_t num_tx_streams, int32_t num_re,
                                int32_t grid_size, int32_t streams_per_rx, int32_t num_terms, int32_t num_data,
                                int32_t bits_per_dim, int32_t method, int32_t hard_out, void* stream) {
    if (batch == 0) return SB_OK;
    SB_CHECK_ARG(d_y && d_no && h_no_stride && d_desired && d_out_stream && d_data_pos && d_re_full && d_t_idx && d_t_w &&
                     d_e_sum && (d_llr || (d_x_hat && d_no_eff)) && (!d_llr || (d_lev_re && d_lev_im)),
                 "sb_ofdm_frontend: null pointer");
    SB_CHECK_ARG(batch > 0 && num_rx > 0 && num_rx_ant > 0 && num_tx_streams > 0 && num_re > 0 && grid_size >= num_re &&
                     num_terms >= 1 && num_data > 0 && (method == 0 || method == 1), "sb_ofdm_frontend: bad sizes");
    SB_CHECK_ARG(streams_per_rx >= 1 && streams_per_rx <= 4, "sb_ofdm_frontend: 1..4 streams per receiver");
    SB_CHECK_ARG(!d_llr || (bits_per_dim >= 1 && bits_per_dim <= 5), "sb_ofdm_frontend: square QAM up to 1024 points");
    SB_CHECK_ARG((d_x_hat == nullptr) == (d_no_eff == nullptr), "sb_ofdm_frontend: x_hat and no_eff go together");
    FrontParams p{};
    p.y = (const float2*)d_y; p.no = d_no;
    for (int i = 0; i < 3; ++i) p.no_stride[i] = h_no_stride[i];
    p.des = d_desired; p.out_ts = d_out_stream; p.data_pos = d_data_pos; p.re_full = d_re_full; p.t_idx = d_t_idx;
    p.t_w = (const float2*)d_t_w; p.e_sum = d_e_sum; p.lev_re = d_lev_re; p.lev_im = d_lev_im; p.llr = d_llr;
    p.xh = (float2*)d_x_hat; p.ne = d_no_eff; p.B = batch; p.RX = num_rx; p.ANT = num_rx_ant; p.TXS = num_tx_streams;
    p.SF = num_re; p.ND = num_data; p.NT = num_terms; p.GRID = grid_size; p.method = method; p.hard_out = hard_out;
    const long long total = batch * num_rx * (long long)num_re;
    const int h = d_llr ? bits_per_dim : 1;
    int rc;
    // 3 streams run the 4-stream code with an all-zero fourth column? No: K is the exact stream count (1, 2, 3 or 4)
    switch (streams_per_rx) {
        case 1: rc = launch_front<1>(p, h, total, (cudaStream_t)stream); break;
        case 2: rc = launch_front<2>(p, h, total, (cudaStream_t)stream); break;
        case 3: rc = launch_front<3>(p, h, total, (cudaStream_t)stream); break;
        default: rc = launch_front<4>(p, h, total, (cudaStream_t)stream); break;
    }
    if (rc) { sb_set_error("sb_ofdm_frontend: unsupported configuration"); return rc; }
    SB_LAUNCH_CHECK();
    return SB_OK;
}
