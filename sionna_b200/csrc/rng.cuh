// rng.cuh -- counter-based Philox4x32-10 (Salmon et al., SC'11) and Box-Muller, device side.
// Replaces tf.random.Generator (Philox) used by BinarySource / complex_normal
// (/root/reference/src/sionna/phy/config.py:109, mapping.py:1346-1352, utils/misc.py:46-52). The streams cannot be
// bit-identical to TensorFlow's (different key schedule / op ordering), so parity tests feed identical tensors to
// both sides instead of identical seeds ("RNG parity").
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint4 philox4x32_10(unsigned long long seed, unsigned long long offset,
                                               unsigned long long ctr) {
    // key = seed; counter = (ctr_lo, ctr_hi, offset_lo, offset_hi)
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = (unsigned)offset, c3 = (unsigned)(offset >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// two independent N(0,1) from two 32-bit words; u1 in (0,1], u2 in [0,1)
__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
    float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(rad * c, rad * s);
}
