// rng.h — counter-based Philox4x32-10 sampling.
//
// The reference draws from a default-seeded thread_local std::mt19937 (src/util.cpp:81-99),
// whose stream depends on bucket-to-thread scheduling, so sample-exact parity is impossible;
// what must match is the DISTRIBUTION of each sampler:
//   rand_double()        uniform [0,1)                         src/util.cpp:95-99
//   rand_int(a,b)        uniform integer in [a,b]              src/util.cpp:83-87
//   unit_disk_sample()   rejection in [-1,1]^2                 src/util.cpp:101-107
//   hemisphere_sample()  uniform sphere flipped to the normal  src/util.cpp:109-129
// Keys: (seed). Counters: (pixel, sample, stream, block) — a function of WHAT is being
// sampled, never of which GPU/thread does it, so N-GPU renders are shard-invariant.
#pragma once
#include "hd.h"

namespace hxr {

struct Rng {
    uint32_t key0, key1;
    uint32_t c0, c1, c2, c3;
    uint32_t out[4];
    int have;

    HXR_HD void init(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stream)
    {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        c0 = pixel; c1 = sample; c2 = stream; c3 = 0;
        have = 0;
    }
    HXR_HD void refill()
    {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, k0 = key0, k1 = key1;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int r = 0; r < 10; r++) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * x0;
            const uint64_t p1 = (uint64_t)0xCD9E8D57u * x2;
            const uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ k0;
            const uint32_t y1 = (uint32_t)p1;
            const uint32_t y2 = (uint32_t)(p0 >> 32) ^ x3 ^ k1;
            const uint32_t y3 = (uint32_t)p0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
        have = 4;
        c3++;
    }
    HXR_HD uint32_t next_u32()
    {
        if (have == 0) refill();
        // consume from the end so `have` doubles as the index
        uint32_t v = have == 4 ? out[0] : (have == 3 ? out[1] : (have == 2 ? out[2] : out[3]));
        have--;
        return v;
    }
    HXR_HD double rand_double()
    {
        const uint32_t a = next_u32() >> 5, b = next_u32() >> 6;  // 27 + 26 = 53 bits
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    HXR_HD int rand_int(int a, int b)
    {
        const uint32_t span = (uint32_t)(b - a) + 1u;
        return a + (int)(((uint64_t)next_u32() * span) >> 32);
    }
};

HXR_HD uint32_t hash_u32(uint32_t a, uint32_t b)
{
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x85EBCA6Bu + (a << 6) + (a >> 2));
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

HXR_HD void unit_disk_sample(Rng& rng, double& x, double& y)
{
    do {
        x = rng.rand_double() * 2 - 1;
        y = rng.rand_double() * 2 - 1;
    } while (x * x + y * y > 1);
}

HXR_HD d3 hemisphere_sample(Rng& rng, const d3& normal)
{
    double u = rng.rand_double();
    double v = rng.rand_double();
    double theta = 2 * HXR_PI * u;
    double cosPhi = 2 * v - 1;
    double sinPhi = sqrt(1 - cosPhi * cosPhi);
    double st, ct;
    sincos(theta, &st, &ct);  // one argument reduction for both
    d3 vec = mk3(ct * sinPhi, cosPhi, st * sinPhi);
    if (dot(vec, normal) < 0) vec = -vec;
    return vec;
}

}  // namespace hxr
