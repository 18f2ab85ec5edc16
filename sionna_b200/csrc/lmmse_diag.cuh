// lmmse_diag.cuh -- LMMSE equalisation of one resource element whose interference-plus-noise covariance is diagonal
// (no interfering streams): everything lives in registers. Shared by the OFDM equaliser kernel (ofdm_mimo.cu) and the
// fused receive front-end (frontend.cu), so both run the same arithmetic.
//   input : B = H_w^H H_w (K x K Hermitian, lower triangle, row a holds (a, 0..a)) and z = H_w^H y_w of the WHITENED
//           channel H_w = S^-1/2 H, y_w = S^-1/2 y
//   A = B + I = C C^H, A^-1 = C^-H C^-1;  G y_w = A^-1 z;  diag(G H_w)_k = sum_j (A^-1)_kj B_jk
//   output: x_hat_k = (G y_w)_k / diag_k, no_eff_k = Re(1 / diag_k - 1)     (mimo/equalization.py:217-231)
#pragma once
#include <cuda_runtime.h>

namespace sb_lmmse {
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {
    float d = b.x * b.x + b.y * b.y;
    return make_float2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

template <int K>
__device__ __forceinline__ void lmmse_diag_solve(const float2* Bm, const float2* z, float2* xh, float* ne) {
    // A = B + I = C C^H (lower, in registers)
    float2 C[K * (K + 1) / 2];
#pragma unroll
    for (int e = 0; e < K * (K + 1) / 2; ++e) C[e] = Bm[e];
#pragma unroll
    for (int a = 0; a < K; ++a) C[a * (a + 1) / 2 + a].x += 1.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        float dj = C[j * (j + 1) / 2 + j].x;
#pragma unroll
        for (int k = 0; k < j; ++k) { float2 l = C[j * (j + 1) / 2 + k]; dj -= l.x * l.x + l.y * l.y; }
        dj = sqrtf(dj);
        C[j * (j + 1) / 2 + j] = make_float2(dj, 0.f);
#pragma unroll
        for (int r = j + 1; r < K; ++r) {
            float2 v = C[r * (r + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = csub(v, cmulc(C[r * (r + 1) / 2 + k], C[j * (j + 1) / 2 + k]));
            C[r * (r + 1) / 2 + j] = make_float2(v.x / dj, v.y / dj);
        }
    }
    // Ci = C^-1 (lower), column by column
    float2 Ci[K * (K + 1) / 2];
#pragma unroll
    for (int c = 0; c < K; ++c) {
#pragma unroll
        for (int r = c; r < K; ++r) {
            float2 v = make_float2(r == c ? 1.f : 0.f, 0.f);
#pragma unroll
            for (int k = c; k < r; ++k) v = csub(v, cmul(C[r * (r + 1) / 2 + k], Ci[k * (k + 1) / 2 + c]));
            float dr = C[r * (r + 1) / 2 + r].x;
            Ci[r * (r + 1) / 2 + c] = make_float2(v.x / dr, v.y / dr);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // row k of A^-1 = C^-H C^-1: (A^-1)_kj = sum_{r >= max(k, j)} conj(Ci[r, k]) Ci[r, j]
        float2 gy = make_float2(0.f, 0.f), dd = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            float2 ainv = make_float2(0.f, 0.f);
#pragma unroll
            for (int r = (k > j ? k : j); r < K; ++r)
                ainv = cadd(ainv, cmulc(Ci[r * (r + 1) / 2 + j], Ci[r * (r + 1) / 2 + k]));
            gy = cadd(gy, cmul(ainv, z[j]));
            // B_jk: stored lower triangle, B_jk = conj(B_kj)
            float2 bjk = j >= k ? Bm[j * (j + 1) / 2 + k] : make_float2(Bm[k * (k + 1) / 2 + j].x, -Bm[k * (k + 1) / 2 + j].y);
            dd = cadd(dd, cmul(ainv, bjk));
        }
        float2 inv = cdiv(make_float2(1.f, 0.f), dd);
        xh[k] = cdiv(gy, dd);
        ne[k] = inv.x - 1.f;
    }
}
}  // namespace sb_lmmse
