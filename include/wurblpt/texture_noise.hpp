/* texture_noise.hpp -- procedural noise textures (mirror of reference libwurblpt/texture_noise.hpp:37-250:
 * TextureValueNoise, TextureGradientNoise, TextureWorleyNoise, TexturePerlinNoise).
 *
 * Same constructors, the same PRNG draws in the same order and the same value() arithmetic, so a texture built
 * from the same Prng is the same function as in the reference. value() runs on the host; for the device a noise
 * texture (or a class derived from one) is sampled into an image when the scene is prepared, like every
 * Texture that has no description of its own (Flattener::texture, texture.hpp). */
#pragma once

#include <cassert>
#include <utility>
#include <vector>

#include "texture.hpp"
#include "prng.hpp"
#include "sampler.hpp"

namespace WurblPT {

namespace detail {

/* A w x h lattice of random values that repeats in both directions, and the position of a texture coordinate
 * in it: cell (ix, iy), offset (fx, fy) inside the cell. */
template<typename V>
class NoiseLattice
{
public:
    int w, h;
    std::vector<V> cells;

    NoiseLattice(int width, int height) : w(width), h(height), cells(size_t(width) * height) {}

    const V& at(int x, int y) const { return cells[size_t(y % h) * w + (x % w)]; }

    void locate(const vec2& tc, int& ix, int& iy, float& fx, float& fy) const
    {
        ix = floor(w * tc.x());
        iy = floor(h * tc.y());
        fx = fract(w * tc.x());
        fy = fract(h * tc.y());
    }
};

inline float smoothstep01(float x) { return x * x * (3.0f - 2.0f * x); }

}

/* bilinear interpolation of random values in [-1,1) (texture_noise.hpp:37-69) */
class TextureValueNoise : public Texture
{
private:
    detail::NoiseLattice<float> _lattice;

public:
    TextureValueNoise(int w, int h, Prng& prng) : _lattice(w, h)
    {
        for (float& v : _lattice.cells)
            v = prng.in01() * 2.0f - 1.0f;
    }

    virtual vec4 value(const vec2& texcoord, float /* t */) const override
    {
        int ix, iy;
        float fx, fy;
        _lattice.locate(texcoord, ix, iy, fx, fy);
        float lower = mix(_lattice.at(ix, iy), _lattice.at(ix + 1, iy), fx);
        float upper = mix(_lattice.at(ix, iy + 1), _lattice.at(ix + 1, iy + 1), fx);
        return vec4(mix(lower, upper, fy));
    }
};

/* random gradients on the lattice corners, smoothstep blending (texture_noise.hpp:71-107) */
class TextureGradientNoise : public Texture
{
private:
    detail::NoiseLattice<vec2> _lattice;

    float corner(int ix, int iy, int dx, int dy, float fx, float fy) const
    {
        return dot(_lattice.at(ix + dx, iy + dy), vec2(fx - float(dx), fy - float(dy)));
    }

public:
    TextureGradientNoise(int w, int h, Prng& prng) : _lattice(w, h)
    {
        for (vec2& g : _lattice.cells)
            g = Sampler::onUnitDisk(prng.in01());
    }

    virtual vec4 value(const vec2& texcoord, float /* t */) const override
    {
        int ix, iy;
        float fx, fy;
        _lattice.locate(texcoord, ix, iy, fx, fy);
        const float sx = detail::smoothstep01(fx);
        const float sy = detail::smoothstep01(fy);
        float lower = mix(corner(ix, iy, 0, 0, fx, fy), corner(ix, iy, 1, 0, fx, fy), sx);
        float upper = mix(corner(ix, iy, 0, 1, fx, fy), corner(ix, iy, 1, 1, fx, fy), sx);
        return vec4(mix(lower, upper, sy));
    }
};

/* distances to the four nearest of n random points on the unit torus (texture_noise.hpp:110-162) */
class TextureWorleyNoise : public Texture
{
private:
    std::vector<vec2> _points;

    static float wrappedDistance(const vec2& p, const vec2& q)
    {
        float dx = abs(q.x() - p.x());
        float dy = abs(q.y() - p.y());
        if (dx > 0.5f)
            dx = 1.0f - dx;
        if (dy > 0.5f)
            dy = 1.0f - dy;
        return sqrt(dx * dx + dy * dy);
    }

public:
    TextureWorleyNoise(int n, Prng& prng) : _points(n)
    {
        assert(n >= 4);
        for (vec2& p : _points)
            p = prng.in01x2();
    }

    virtual vec4 value(const vec2& texcoord, float /* t */) const override
    {
        const vec2 uv = fract(texcoord);
        float nearest[4] = { maxval, maxval, maxval, maxval };      /* ascending */
        for (const vec2& p : _points) {
            const float d = wrappedDistance(uv, p);
            int k = 4;
            while (k > 0 && d < nearest[k - 1])
                k--;
            if (k < 4) {
                for (int m = 3; m > k; m--)
                    nearest[m] = nearest[m - 1];
                nearest[k] = d;
            }
        }
        return vec4(nearest[0], nearest[1], nearest[2], nearest[3]);
    }
};

/* Perlin noise after Shirley's ray tracing books: 256 random unit vectors, three permutations, the z = 128 slice
 * of the 3D noise (texture_noise.hpp:165-250) */
class TexturePerlinNoise : public Texture
{
private:
    static constexpr int Size = 256;
    std::vector<vec3> _gradients;
    std::vector<int> _permX, _permY, _permZ;

    static float blend(int corner, float s) { return corner * s + (1.0f - corner) * (1.0f - s); }

public:
    TexturePerlinNoise(Prng& prng) : _gradients(Size), _permX(Size), _permY(Size), _permZ(Size)
    {
        for (int i = 0; i < Size; i++) {
            _gradients[i] = Sampler::onUnitSphere(prng.in01x2());
            _permX[i] = _permY[i] = _permZ[i] = i;
        }
        /* three interleaved Fisher-Yates shuffles: one draw for x, y and z per step */
        for (int i = Size - 1; i > 0; i--) {
            int sx = prng.in01() * (i + 1);
            std::swap(_permX[i], _permX[sx]);
            int sy = prng.in01() * (i + 1);
            std::swap(_permY[i], _permY[sy]);
            int sz = prng.in01() * (i + 1);
            std::swap(_permZ[i], _permZ[sz]);
        }
    }

    virtual vec4 value(const vec2& texcoords, float /* t */) const override
    {
        const float x = fract(texcoords.x()) * Size;
        const float y = fract(texcoords.y()) * Size;
        const float z = 0.5f * Size;
        const float u = fract(x), v = fract(y), w = fract(z);
        const int i = x, j = y, k = z;
        const float su = detail::smoothstep01(u), sv = detail::smoothstep01(v), sw = detail::smoothstep01(w);
        float sum = 0.0f;
        for (int di = 0; di < 2; di++) {
            for (int dj = 0; dj < 2; dj++) {
                for (int dk = 0; dk < 2; dk++) {
                    const vec3& g = _gradients[_permX[(i + di) & (Size - 1)] ^ _permY[(j + dj) & (Size - 1)]
                        ^ _permZ[(k + dk) & (Size - 1)]];
                    sum += blend(di, su) * blend(dj, sv) * blend(dk, sw) * dot(g, vec3(u - di, v - dj, w - dk));
                }
            }
        }
        return vec4(sum, sum, sum, 1.0f);
    }

    vec4 turbulence(const vec2& texcoord, float t, int freqs = 7, float minFreq = 1.0f) const
    {
        float sum = 0.0f;
        float amplitude = 1.0f;
        float frequency = minFreq;
        for (int octave = 0; octave < freqs; octave++) {
            sum += amplitude * TexturePerlinNoise::value(texcoord * frequency, t).r();
            frequency *= 2.0;
            amplitude /= 2.0;
        }
        return abs(vec4(sum, sum, sum, 0.0f));
    }
};

}
