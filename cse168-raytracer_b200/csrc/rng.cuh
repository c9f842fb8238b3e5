// rng.cuh -- counter-based random numbers (Philox4x32-10) shared by device ray generation and the host
// helper mirogpu_rng_uniforms.  The reference draws from rand() (Utility.h:14-17), unseeded and racy under
// OpenMP, so its streams cannot be matched; instead every random decision is a pure function of
// (seed, element index, sample, dimension), and parity is established by handing the oracle the same
// uniforms (or the generated rays themselves).
#ifndef MIROGPU_RNG_CUH
#define MIROGPU_RNG_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace mirogpu {

#ifndef MIRO_HD
#define MIRO_HD __host__ __device__ __forceinline__
#endif

MIRO_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// Philox4x32 with 10 rounds (Salmon et al., SC'11).  ctr = (index, sample, dimension, 0), key = (seed, tag).
MIRO_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two uniforms in [0, 1) with 24 random bits each (every value is exactly representable in binary32).
MIRO_HD void uniform2(uint32_t seed, uint32_t index, uint32_t sample, uint32_t dimension, float& u1, float& u2)
{
    uint32_t r[4];
    philox4x32_10(index, sample, dimension, 0u, seed, 0x4D49524Fu /* "MIRO" */, r);
    u1 = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
    u2 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
}

enum { RNG_DIM_PIXEL = 0, RNG_DIM_BOUNCE = 1, RNG_DIM_LIGHT = 2 };

}  // namespace mirogpu
#endif
