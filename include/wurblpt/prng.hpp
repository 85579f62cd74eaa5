/* prng.hpp -- host-side Prng, used by applications to place objects (mirror of reference
 * libwurblpt/prng.hpp:47-101): xoshiro128+ (Blackman & Vigna, public domain) seeded with
 * splitmix64(pixelIndex + 42). The device carries its own copy (csrc/device_math.cuh). */
#pragma once

#include <cstdint>

#include "gvm.hpp"

namespace WurblPT {

class Prng
{
private:
    uint32_t s[4];

    static uint64_t splitmix64(uint64_t x)
    {
        uint64_t z = (x += 0x9e3779b97f4a7c15);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9;
        z = (z ^ (z >> 27)) * 0x94d049bb133111eb;
        return z ^ (z >> 31);
    }

    uint32_t next()
    {
        const uint32_t result = s[0] + s[3];
        const uint32_t t = s[1] << 9;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = (s[3] << 11) | (s[3] >> 21);
        return result;
    }

public:
    Prng(unsigned int pixelIndex)
    {
        uint64_t seed = pixelIndex;
        seed += 42;
        uint64_t s01 = splitmix64(seed);
        uint64_t s23 = splitmix64(s01);
        s[0] = s01 >> 32ULL;
        s[1] = s01 & 0xffffffffULL;
        s[2] = s23 >> 32ULL;
        s[3] = s23 & 0xffffffffULL;
    }

    float in01()
    {
        uint32_t x = next();
        return (x >> 8) * 0x1.0p-24;
    }

    vec2 in01x2() { return vec2(in01(), in01()); }
};

}
