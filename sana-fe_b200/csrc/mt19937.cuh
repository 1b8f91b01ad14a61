// mt19937.cuh — MT19937 (Matsumoto & Nishimura 1998) and the 53-bit canonical double that libstdc++'s
// std::uniform_real_distribution<double>{0,1} builds from two of its 32-bit outputs (low word first):
// U = (lo + hi * 2^32) / 2^64, nudged below 1.0. This is the generator behind the reference's Poisson inputs
// (InputModel, src/models.hpp:347-366, src/models.cpp:880-885).
//
// Written once for host and device: the host compilation is checked draw for draw against std::mt19937 +
// std::uniform_real_distribution (tests/test_poisson_inputs.py), the device compilation is the same code.
// The state is addressed through a stride so that the device can interleave the states of many generators
// (word i of generator g at mt[i * stride + g]: neighbouring threads touch neighbouring words).
#ifndef SFE_MT19937_CUH_
#define SFE_MT19937_CUH_

#include <cstdint>

#ifdef __CUDACC__
#define SFE_MT_HD __host__ __device__ __forceinline__
#else
#define SFE_MT_HD inline
#endif

namespace sfe
{
constexpr uint32_t kMtWords = 624;

SFE_MT_HD void mt_seed(uint32_t *mt, const size_t stride, uint32_t *idx, const uint32_t seed)
{
    uint32_t prev = seed;
    mt[0] = prev;
    for (uint32_t i = 1; i < kMtWords; ++i)
    {
        prev = 1812433253u * (prev ^ (prev >> 30)) + i;
        mt[i * stride] = prev;
    }
    *idx = kMtWords;
}

SFE_MT_HD uint32_t mt_next(uint32_t *mt, const size_t stride, uint32_t *idx)
{
    if (*idx >= kMtWords)
    {
        for (uint32_t i = 0; i < kMtWords; ++i)
        {
            const uint32_t y = (mt[i * stride] & 0x80000000u) | (mt[((i + 1) % kMtWords) * stride] & 0x7fffffffu);
            mt[i * stride] = mt[((i + 397) % kMtWords) * stride] ^ (y >> 1) ^ ((y & 1u) != 0u ? 0x9908b0dfu : 0u);
        }
        *idx = 0;
    }
    uint32_t y = mt[(*idx)++ * stride];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

SFE_MT_HD double mt_canonical(uint32_t *mt, const size_t stride, uint32_t *idx)
{
    const double lo = static_cast<double>(mt_next(mt, stride, idx));
    const double hi = static_cast<double>(mt_next(mt, stride, idx));
    const double u = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    return u >= 1.0 ? 0.99999999999999988897769753748 : u; // nextafter(1.0, 0.0)
}
} // namespace sfe
#endif
