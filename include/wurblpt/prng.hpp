/* prng.hpp -- the random number stream of the WurblPT API on the host (reference libwurblpt/prng.hpp:47-101):
 * the xoshiro128+ generator of Blackman and Vigna, its four state words filled from two rounds of the
 * splitmix64 mixer started at pixelIndex + 42. Applications use it to place objects; the device carries its own
 * copy (csrc/device_math.cuh) and tests/ compare both with the reference's known answers. */
#pragma once

#include <cstdint>

#include "gvm.hpp"

namespace WurblPT {

class Prng
{
private:
    uint32_t s[4];

    static uint32_t rotateLeft(uint32_t word, int bits) { return (word << bits) | (word >> (32 - bits)); }

    /* one step of the splitmix64 sequence: advances `state`, returns the mixed output */
    static uint64_t splitmix64(uint64_t& state)
    {
        state += 0x9e3779b97f4a7c15ULL;
        uint64_t mixed = state;
        mixed = (mixed ^ (mixed >> 30)) * 0xbf58476d1ce4e5b9ULL;
        mixed = (mixed ^ (mixed >> 27)) * 0x94d049bb133111ebULL;
        return mixed ^ (mixed >> 31);
    }

    uint32_t next()
    {
        const uint32_t output = s[0] + s[3];
        const uint32_t shifted = s[1] << 9;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= shifted;
        s[3] = rotateLeft(s[3], 11);
        return output;
    }

public:
    Prng(unsigned int pixelIndex)
    {
        /* the second mixer round starts from the OUTPUT of the first (not from its advanced state) */
        uint64_t first = uint64_t(pixelIndex) + 42;
        const uint64_t low = splitmix64(first);
        uint64_t second = low;
        const uint64_t high = splitmix64(second);
        s[0] = uint32_t(low >> 32);
        s[1] = uint32_t(low);
        s[2] = uint32_t(high >> 32);
        s[3] = uint32_t(high);
    }

    /* uniform in [0,1): the upper 24 bits of a 32-bit output */
    float in01() { return (next() >> 8) * 0x1.0p-24; }

    /* two draws; note that the compiler is free to evaluate the two calls in either order (g++ takes the
     * second argument first), and that the reference has the same expression */
    vec2 in01x2() { return vec2(in01(), in01()); }
};

}
