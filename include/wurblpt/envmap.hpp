/* envmap.hpp -- environment maps with parameterisation-independent importance sampling (mirror of
 * reference libwurblpt/envmap.hpp:44-285; method: M. Lambers, arXiv:2208.10815).
 * initializeImportanceSampling() runs on the host exactly like the reference (it evaluates L() on an
 * N x N equal-area grid, sorts the bins, builds the CDF); the tables are then uploaded and p()/d()
 * run on the device. */
#pragma once

#include <algorithm>
#include <vector>

#include "gvm.hpp"
#include "prng.hpp"
#include "texture.hpp"
#include "flatten.hpp"

namespace WurblPT {

class EnvironmentMap
{
private:
    int N;
    std::vector<float> M;       // importance per bin, normalised
    std::vector<int> Ms;        // bins by descending importance
    std::vector<float> Mcs;     // cumulative importance along Ms

    /* square [0,1]^2 -> unit sphere, equal area (Shirley's square-to-disk, then Lambert) */
    static vec3 invm(const vec2& uv)
    {
        float u = 2.0f * uv.x() - 1.0f;
        float v = 2.0f * uv.y() - 1.0f;
        float r, alpha;
        if (u * u > v * v) {
            r = u;
            alpha = pi_4 * v / u;
        } else {
            r = v;
            if (abs(v) > 0.0f)
                alpha = pi_2 - pi_4 * u / v;
            else
                alpha = 0.0f;
        }
        float lat = pi_2 - 2.0f * asin(r);
        float lon = alpha + pi_2;
        vec3 d = vec3(-cos(lat) * sin(lon), sin(lat), cos(lat) * cos(lon));
        return normalize(d);
    }

    /* unit sphere -> square, inverse of invm() */
    static vec2 m(const vec3& d)
    {
        float lat = asin(clamp(d.y(), -1.0f, +1.0f));
        float lon = atan2(-d.x(), d.z());
        float r = sin(0.5f * (pi_2 - lat));
        float alpha = lon - pi_2;
        float u, v;
        if (alpha < -pi_4)
            alpha += 2.0f * pi;
        if (alpha < pi_4) {
            u = r;
            v = alpha * u / pi_4;
        } else if (alpha < pi_2 + pi_4) {
            v = r;
            u = -(alpha - pi_2) * v / pi_4;
        } else if (alpha < pi + pi_4) {
            u = -r;
            v = (alpha - pi) * u / pi_4;
        } else {
            v = -r;
            u = -(alpha - (pi + pi_2)) * v / pi_4;
        }
        return vec2(0.5f * (u + 1.0f), 0.5f * (v + 1.0f));
    }

public:
    EnvironmentMap() : N(0) {}
    virtual ~EnvironmentMap() {}

    void initializeImportanceSampling(int n)
    {
        N = n;
        M.resize(N * N);
        Ms.resize(N * N);
        Mcs.resize(N * N);
        float totalImportance = 0.0f;
        for (int y = 0; y < N; y++) {
            float v = (y + 0.5f) / N;
            for (int x = 0; x < N; x++) {
                float u = (x + 0.5f) / N;
                vec4 radiance = L(invm(vec2(u, v)));
                float importance = radiance.x() + radiance.y() + radiance.z() + radiance.w();
                totalImportance += importance;
                M[y * N + x] = importance;
            }
        }
        for (int i = 0; i < N * N; i++)
            M[i] /= totalImportance;
        for (int i = 0; i < N * N; i++)
            Ms[i] = i;
        std::sort(Ms.begin(), Ms.end(), [this](unsigned int i, unsigned int j) { return this->M[i] > this->M[j]; });
        float sum = 0.0f;
        for (int i = 0; i < N * N; i++) {
            sum += M[Ms[i]];
            Mcs[i] = sum;
        }
    }

    bool supportsImportanceSampling() const { return N > 0; }

    virtual vec4 L(const vec3& direction, float t = 0.0f) const = 0;

    float p(const vec3& direction) const
    {
        vec2 uv = m(direction);
        int x = uv.x() * N;
        int y = uv.y() * N;
        if (x >= N)
            x = N - 1;
        if (y >= N)
            y = N - 1;
        float invBinSizeOnSphere = (N * N) * 0.25f * inv_pi;
        return M[y * N + x] * invBinSizeOnSphere;
    }

    vec3 d(Prng& prng) const
    {
        float r = prng.in01();
        int a = 0;
        int b = N * N - 1;
        while (b > a + 1) {
            int c = (a + b) / 2;
            if (Mcs[c] < r)
                a = c;
            else
                b = c;
        }
        int bin = Ms[Mcs[a] >= r ? a : b];
        float u = ((bin % N) + prng.in01()) / N;
        float v = ((bin / N) + prng.in01()) / N;
        return invm(vec2(u, v));
    }

    /* type-specific part of the description; returns false for user-defined subclasses */
    virtual bool flattenType(Flattener& /* F */, wpt_envmap& /* out */) const { return false; }

    bool flatten(Flattener& F) const
    {
        if (!flattenType(F, F.envmap))
            return false;
        F.envmap.is_n = N;
        F.envM = M;
        F.envMs.assign(Ms.begin(), Ms.end());
        F.envMcs = Mcs;
        return true;
    }
};

class EnvironmentMapEquiRect final : public EnvironmentMap
{
public:
    enum Compatibility {
        CompatibilityMitsuba,
        CompatibilitySurroundVideo
    };

private:
    const Compatibility _compatibility;
    const Texture* _tex;

public:
    EnvironmentMapEquiRect(const Texture* tex, Compatibility compat = CompatibilityMitsuba) : _compatibility(compat), _tex(tex) {}

    virtual vec4 L(const vec3& direction, float t = 0.0f) const override
    {
        float y = asin(clamp(direction.y(), -1.0f, 1.0f));
        float x = atan(-direction.x(), direction.z());
        if (_compatibility == CompatibilityMitsuba) {
            x -= pi;
            if (x < 0.0f)
                x += 2.0f * pi;
        }
        x *= 0.5f * inv_pi;
        y = y * inv_pi + 0.5f;
        return _tex->value(vec2(x, y), t);
    }

    virtual bool flattenType(Flattener& F, wpt_envmap& out) const override
    {
        out.type = WPT_ENV_EQUIRECT;
        out.compat = (_compatibility == CompatibilityMitsuba ? WPT_ENV_COMPAT_MITSUBA : WPT_ENV_COMPAT_SURROUND_VIDEO);
        out.tex[0] = F.texture(_tex);
        return out.tex[0] >= 0;
    }
};

class EnvironmentMapCube final : public EnvironmentMap
{
private:
    const Texture* _cubesides[6];   // posx negx posy negy posz negz

public:
    EnvironmentMapCube(const Texture* posx, const Texture* negx, const Texture* posy, const Texture* negy,
            const Texture* posz, const Texture* negz) : _cubesides { posx, negx, posy, negy, posz, negz }
    {
    }

    virtual vec4 L(const vec3& direction, float t = 0.0f) const override
    {
        float ax = abs(direction.x());
        float ay = abs(direction.y());
        float az = abs(direction.z());
        int cubeside;
        float u, v;
        if (ax > ay && ax > az) {
            u = 0.5f * (direction.z() / -direction.x() + 1.0f);
            v = 0.5f * (direction.y() / ax + 1.0f);
            cubeside = 0 + signbit(direction.x());
        } else if (ay > az) {
            u = 0.5f * (direction.x() / ay + 1.0f);
            v = 0.5f * (direction.z() / -direction.y() + 1.0f);
            cubeside = 2 + signbit(direction.y());
        } else {
            u = 0.5f * (direction.x() / direction.z() + 1.0f);
            v = 0.5f * (direction.y() / az + 1.0f);
            cubeside = 4 + signbit(direction.z());
        }
        return _cubesides[cubeside]->value(vec2(u, v), t);
    }

    virtual bool flattenType(Flattener& F, wpt_envmap& out) const override
    {
        out.type = WPT_ENV_CUBE;
        for (int i = 0; i < 6; i++) {
            out.tex[i] = F.texture(_cubesides[i]);
            if (out.tex[i] < 0)
                return false;
        }
        return true;
    }
};

}
