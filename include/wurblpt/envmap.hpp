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

namespace detail {

/* The area-preserving map between the unit square and the unit sphere that the importance tables live on
 * (Lambers 2022; reference envmap.hpp:55-118): Shirley and Chiu's concentric square-to-disk map followed by
 * Lambert's azimuthal equal-area projection, with +y as the pole direction. The statements keep the
 * reference's arithmetic, because bin boundaries decide which table entry a direction reads. */
struct SquareSphereMap
{
    /* disk radius and angle of a point of [-1,1]^2 */
    static void squareToPolar(float u, float v, float& radius, float& angle)
    {
        const bool uDominates = u * u > v * v;
        radius = uDominates ? u : v;
        if (uDominates)
            angle = pi_4 * v / u;
        else
            angle = (abs(v) > 0.0f ? pi_2 - pi_4 * u / v : 0.0f);
    }

    static vec3 toSphere(const vec2& uv)
    {
        float radius, angle;
        squareToPolar(2.0f * uv.x() - 1.0f, 2.0f * uv.y() - 1.0f, radius, angle);
        const float latitude = pi_2 - 2.0f * asin(radius);
        const float longitude = angle + pi_2;
        return normalize(vec3(-cos(latitude) * sin(longitude), sin(latitude), cos(latitude) * cos(longitude)));
    }

    static vec2 toSquare(const vec3& direction)
    {
        const float latitude = asin(clamp(direction.y(), -1.0f, +1.0f));
        const float longitude = atan2(-direction.x(), direction.z());
        const float radius = sin(0.5f * (pi_2 - latitude));
        float angle = longitude - pi_2;
        if (angle < -pi_4)
            angle += 2.0f * pi;
        /* the four triangles of the square, counter-clockwise from the one around +u */
        const int triangle = (angle < pi_4 ? 0 : angle < pi_2 + pi_4 ? 1 : angle < pi + pi_4 ? 2 : 3);
        float u = 0.0f, v = 0.0f;
        switch (triangle) {
        case 0:
            u = radius;
            v = angle * u / pi_4;
            break;
        case 1:
            v = radius;
            u = -(angle - pi_2) * v / pi_4;
            break;
        case 2:
            u = -radius;
            v = (angle - pi) * u / pi_4;
            break;
        default:
            v = -radius;
            u = -(angle - (pi + pi_2)) * v / pi_4;
            break;
        }
        return vec2(0.5f * (u + 1.0f), 0.5f * (v + 1.0f));
    }
};

}

class EnvironmentMap
{
private:
    int N;                      /* bins per side of the square; 0 = no importance sampling */
    std::vector<float> M;       /* normalised importance of bin y * N + x */
    std::vector<int> Ms;        /* bin numbers, most important first */
    std::vector<float> Mcs;     /* running sum of M along Ms */

    int binOf(const vec3& direction) const
    {
        const vec2 uv = detail::SquareSphereMap::toSquare(direction);
        int column = uv.x() * N;
        int row = uv.y() * N;
        column = (column >= N ? N - 1 : column);
        row = (row >= N ? N - 1 : row);
        return row * N + column;
    }

    /* index into Ms of the first bin whose running sum reaches r (bisection as in the reference) */
    int rankFor(float r) const
    {
        int below = 0;
        int above = N * N - 1;
        while (above > below + 1) {
            const int middle = (below + above) / 2;
            if (Mcs[middle] < r)
                below = middle;
            else
                above = middle;
        }
        return Mcs[below] >= r ? below : above;
    }

public:
    EnvironmentMap() : N(0) {}
    virtual ~EnvironmentMap() {}

    virtual vec4 L(const vec3& direction, float t = 0.0f) const = 0;

    /* Tabulates the radiance (sum of the four channels) at the bin centres, normalises, ranks the bins and
     * accumulates; single precision sums in this order, like the reference (envmap.hpp:121-158). */
    void initializeImportanceSampling(int n)
    {
        N = n;
        const int bins = N * N;
        M.assign(bins, 0.0f);
        float total = 0.0f;
        for (int bin = 0; bin < bins; bin++) {
            const float v = ((bin / N) + 0.5f) / N;
            const float u = ((bin % N) + 0.5f) / N;
            const vec4 radiance = L(detail::SquareSphereMap::toSphere(vec2(u, v)));
            const float importance = radiance.x() + radiance.y() + radiance.z() + radiance.w();
            total += importance;
            M[bin] = importance;
        }
        for (float& importance : M)
            importance /= total;
        Ms.resize(bins);
        for (int bin = 0; bin < bins; bin++)
            Ms[bin] = bin;
        std::sort(Ms.begin(), Ms.end(), [this](unsigned int i, unsigned int j) { return this->M[i] > this->M[j]; });
        Mcs.resize(bins);
        float running = 0.0f;
        for (int rank = 0; rank < bins; rank++) {
            running += M[Ms[rank]];
            Mcs[rank] = running;
        }
    }

    bool supportsImportanceSampling() const { return N > 0; }

    /* solid-angle density of d() in `direction` */
    float p(const vec3& direction) const
    {
        const float binsPerSteradian = (N * N) * 0.25f * inv_pi;
        return M[binOf(direction)] * binsPerSteradian;
    }

    /* a direction drawn proportionally to the tabulated radiance: one draw picks the bin, two place the
     * direction inside it */
    vec3 d(Prng& prng) const
    {
        const int bin = Ms[rankFor(prng.in01())];
        const float u = ((bin % N) + prng.in01()) / N;
        const float v = ((bin / N) + prng.in01()) / N;
        return detail::SquareSphereMap::toSphere(vec2(u, v));
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
