// Philox4x32-10 (Salmon et al., SC'11) -- device + host, bit-identical to oracle/philox.py.
// Counter layout: (c0, c1, c2, tag); key = (seed lo, seed hi).
#pragma once
#include <stdint.h>

namespace gl {

enum : uint32_t {
    GL_TAG_INIT = 0,   // initial population: c0,c1 = global row lo/hi, c2 = component
    GL_TAG_WIDTH = 1,  // initial proposal widths: c0 = draw (<100), c2 = component
    GL_TAG_PICK = 2,   // survivor pick: c0 = dead slot, c1 = step, c2 = attempt
    GL_TAG_MOVE = 3,   // trial move: c0 = dead slot, c1 = step, c2 = sweep*d + component
    GL_TAG_USER = 4    // gl_prior_draw
};

struct U2 {
    double a, b;
};

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t m, uint32_t x, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
    lo = m * x;
    hi = __umulhi(m, x);
#else
    uint64_t p = (uint64_t)m * x;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two 32-bit words -> double strictly inside (0,1): (2m+1)*2^-53, m = 52 random bits
__host__ __device__ __forceinline__ double u52(uint32_t a, uint32_t b) {
    unsigned long long m = ((unsigned long long)(a >> 6) << 26) | (unsigned long long)(b >> 6);
    return (2.0 * (double)m + 1.0) * 1.1102230246251565404e-16;   // 2^-53
}

__host__ __device__ __forceinline__ U2 philox_uniforms(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t tag) {
    uint32_t w[4];
    philox4x32_10(c0, c1, c2, tag, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    U2 r;
    r.a = u52(w[0], w[1]);
    r.b = u52(w[2], w[3]);
    return r;
}

}  // namespace gl
