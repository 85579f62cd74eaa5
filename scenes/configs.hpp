/*
 * The benchmark / parity scene set of this repository, written against the public
 * WurblPT C++ API only (Scene::take, Mesh, MeshInstance, Sphere, Medium, generate*,
 * Material*, Light*, EnvironmentMapEquiRect, createTextureImage, Camera, Optics,
 * Parameters). Include <wurblpt/wurblpt.hpp> before this header.
 *
 * Because it only uses that API, the same file compiles
 *   - against this repository's API mirror (include/wurblpt/) -> product path, and
 *   - against the unmodified reference headers (oracle/_ref build) -> oracle path,
 * which is how both arms are guaranteed to render identical scenes.
 *
 * Scenes follow BASELINE.json's configs:
 *   cornell  C1  reference wurblpt-cornellbox/wurblpt-cornellbox.cpp:44-228 (default variant)
 *   rtiow    C2  reference wurblpt-rtiow/wurblpt-rtiow.cpp:37-82
 *   synth    C3  procedural ~1M-triangle courtyard, GGX, HDR envmap with importance sampling
 *   synth360 C4  synth scene, 360 degree top-bottom stereo camera (pattern: wurblpt-sponza.cpp:120-148)
 *   tof      C5  reference wurblpt-tof-example/wurblpt-tof-example.cpp:60-87
 *   media    C5  reference wurblpt-participating-media/wurblpt-participating-media.cpp:32-206
 *   furnace      reference wurblpt-furnace-test/wurblpt-furnace-test.cpp:34-80
 *   mis / mis_nohot  reference wurblpt-mis-test/wurblpt-mis-test.cpp:34-97
 */
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace wptcfg {

using namespace WurblPT;

/* ---- small helpers ---- */

struct QuadDef {
    float p[4][3];
    float n[4][3];
    int texcoordVariant;    // 0: (0,0)(1,0)(1,1)(0,1); 1: (0,1)(0,0)(1,0)(1,1)
};

inline Mesh* meshFromQuad(const QuadDef& q)
{
    std::vector<vec3> pos, nrm;
    for (int i = 0; i < 4; i++) {
        pos.push_back(vec3(q.p[i][0], q.p[i][1], q.p[i][2]));
        nrm.push_back(vec3(q.n[i][0], q.n[i][1], q.n[i][2]));
    }
    std::vector<vec2> tc;
    if (q.texcoordVariant == 0)
        tc = { vec2(0.0f, 0.0f), vec2(1.0f, 0.0f), vec2(1.0f, 1.0f), vec2(0.0f, 1.0f) };
    else
        tc = { vec2(0.0f, 1.0f), vec2(0.0f, 0.0f), vec2(1.0f, 0.0f), vec2(1.0f, 1.0f) };
    return new Mesh(pos, nrm, tc, { 0, 1, 2, 0, 2, 3 });
}

#define WPTCFG_N4(x, y, z) { { x, y, z }, { x, y, z }, { x, y, z }, { x, y, z } }

/* The Cornell box geometry (the classic data set; same numbers as the Mitsuba
 * cbox scene shipped next to the reference app). Order matters: it is the BVH
 * builder's input order. */
static const QuadDef cornellWalls[5] = {
    // left wall (red)
    { { { -1.01f, 0.0f, 0.99f }, { -0.99f, 0.0f, -1.04f }, { -1.02f, 1.99f, -1.04f }, { -1.02f, 1.99f, 0.99f } },
      { { 0.9999874f, 0.005025057f, 0.0f }, { 0.9998379f, 0.01507292f, 0.009850611f },
        { 0.9999874f, 0.005025057f, 0.0f }, { 0.9999874f, 0.005025057f, 0.0f } }, 0 },
    // right wall (green)
    { { { 1.0f, 0.0f, -1.04f }, { 1.0f, 0.0f, 0.99f }, { 1.0f, 1.99f, 0.99f }, { 1.0f, 1.99f, -1.04f } },
      WPTCFG_N4(-1.0f, 0.0f, 0.0f), 0 },
    // floor
    { { { -1.01f, 0.0f, 0.99f }, { 1.0f, 0.0f, 0.99f }, { 1.0f, 0.0f, -1.04f }, { -0.99f, 0.0f, -1.04f } },
      WPTCFG_N4(0.0f, 1.0f, 0.0f), 0 },
    // ceiling
    { { { -1.02f, 1.99f, 0.99f }, { -1.02f, 1.99f, -1.04f }, { 1.0f, 1.99f, -1.04f }, { 1.0f, 1.99f, 0.99f } },
      WPTCFG_N4(0.0f, -1.0f, 0.0f), 0 },
    // back wall
    { { { -0.99f, 0.0f, -1.04f }, { 1.0f, 0.0f, -1.04f }, { 1.0f, 1.99f, -1.04f }, { -1.02f, 1.99f, -1.04f } },
      WPTCFG_N4(0.0f, 0.0f, 1.0f), 0 }
};

static const QuadDef cornellShortBox[6] = {
    { { { -0.05f, 0.0f, 0.57f }, { -0.05f, 0.6f, 0.57f }, { 0.13f, 0.6f, 0.0f }, { 0.13f, 0.0f, 0.0f } },
      WPTCFG_N4(-0.9535826f, 0.0f, -0.3011314f), 0 },
    { { { 0.7f, 0.0f, 0.17f }, { 0.7f, 0.6f, 0.17f }, { 0.53f, 0.6f, 0.75f }, { 0.53f, 0.0f, 0.75f } },
      WPTCFG_N4(0.9596285f, 0.0f, 0.2812705f), 0 },
    { { { 0.53f, 0.0f, 0.75f }, { 0.7f, 0.0f, 0.17f }, { 0.13f, 0.0f, 0.0f }, { -0.05f, 0.0f, 0.57f } },
      WPTCFG_N4(0.0f, -1.0f, 0.0f), 0 },
    { { { 0.53f, 0.6f, 0.75f }, { 0.7f, 0.6f, 0.17f }, { 0.13f, 0.6f, 0.0f }, { -0.05f, 0.6f, 0.57f } },
      WPTCFG_N4(0.0f, 1.0f, 0.0f), 0 },
    { { { 0.13f, 0.0f, 0.0f }, { 0.13f, 0.6f, 0.0f }, { 0.7f, 0.6f, 0.17f }, { 0.7f, 0.0f, 0.17f } },
      WPTCFG_N4(0.2858051f, 0.0f, -0.9582878f), 0 },
    { { { 0.53f, 0.0f, 0.75f }, { 0.53f, 0.6f, 0.75f }, { -0.05f, 0.6f, 0.57f }, { -0.05f, 0.0f, 0.57f } },
      WPTCFG_N4(-0.2963993f, 0.0f, 0.9550642f), 0 }
};

static const QuadDef cornellTallBox[6] = {
    { { { -0.53f, 0.0f, 0.09f }, { -0.53f, 1.2f, 0.09f }, { -0.71f, 1.2f, -0.49f }, { -0.71f, 0.0f, -0.49f } },
      WPTCFG_N4(-0.9550642f, 0.0f, 0.2963992f), 0 },
    { { { -0.14f, 0.0f, -0.67f }, { -0.14f, 1.2f, -0.67f }, { 0.04f, 1.2f, -0.09f }, { 0.04f, 0.0f, -0.09f } },
      WPTCFG_N4(0.9550642f, 0.0f, -0.2963992f), 0 },
    { { { -0.53f, 0.0f, 0.09f }, { 0.04f, 0.0f, -0.09f }, { -0.14f, 0.0f, -0.67f }, { -0.71f, 0.0f, -0.49f } },
      WPTCFG_N4(0.0f, -1.0f, 0.0f), 0 },
    { { { -0.53f, 1.2f, 0.09f }, { 0.04f, 1.2f, -0.09f }, { -0.14f, 1.2f, -0.67f }, { -0.71f, 1.2f, -0.49f } },
      WPTCFG_N4(0.0f, 1.0f, 0.0f), 0 },
    { { { -0.71f, 0.0f, -0.49f }, { -0.71f, 1.2f, -0.49f }, { -0.14f, 1.2f, -0.67f }, { -0.14f, 0.0f, -0.67f } },
      WPTCFG_N4(-0.3011314f, 0.0f, -0.9535826f), 0 },
    { { { 0.04f, 0.0f, -0.09f }, { 0.04f, 1.2f, -0.09f }, { -0.53f, 1.2f, 0.09f }, { -0.53f, 0.0f, 0.09f } },
      WPTCFG_N4(0.3011314f, 0.0f, 0.9535826f), 0 }
};

static const QuadDef cornellLight =
    { { { -0.24f, 1.98f, 0.16f }, { -0.24f, 1.98f, -0.22f }, { 0.23f, 1.98f, -0.22f }, { 0.23f, 1.98f, 0.16f } },
      WPTCFG_N4(0.0f, -1.0f, 0.0f), 1 };

/* ---- C1: Cornell box ---- */

inline void buildCornell(Scene& scene)
{
    Material* white = scene.take(new MaterialLambertian(vec3(0.725f, 0.71f, 0.68f)));
    Material* red = scene.take(new MaterialLambertian(vec3(0.63f, 0.065f, 0.05f)));
    Material* green = scene.take(new MaterialLambertian(vec3(0.14f, 0.45f, 0.091f)));
    Material* light = scene.take(new LightDiffuse(vec3(4.0f)));
    // The reference app also registers these two (unused in the default variant);
    // keep them so material indices agree.
    scene.take(new MaterialGGX(vec3(1.0f), vec2(0.04f)));
    scene.take(new MaterialGlass(vec3(0.2f), 1.5f));

    Material* wallMaterials[5] = { red, green, white, white, white };
    for (int i = 0; i < 5; i++)
        scene.take(new MeshInstance(scene.take(meshFromQuad(cornellWalls[i])), wallMaterials[i]));
    for (int i = 0; i < 6; i++)
        scene.take(new MeshInstance(scene.take(meshFromQuad(cornellShortBox[i])), white, -1));
    for (int i = 0; i < 6; i++)
        scene.take(new MeshInstance(scene.take(meshFromQuad(cornellTallBox[i])), white));
    scene.take(new MeshInstance(scene.take(meshFromQuad(cornellLight)), light), HotSpot);
}

/* ---- C2: "Ray Tracing in One Weekend" random spheres ---- */

inline void buildRtiow(Scene& scene)
{
    Prng prng(17);
    Material* material;

    material = scene.take(new MaterialLambertian(vec3(0.5f)));
    scene.take(new Sphere(vec3(0.0f, -1000.0f, 0.0f), 1000.0f, material));

    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            // Note: each draw is a separate statement so the sequence does not depend on
            // the compiler's argument evaluation order. The reference app draws inside
            // constructor argument lists; with g++ those are evaluated right to left
            // (SURVEY.md section 8c), which the statements below reproduce.
            float chooseMaterial = prng.in01();
            float cz = prng.in01();
            float cx = prng.in01();
            vec3 center(a + 0.9f * cx, 0.2f, b + 0.9f * cz);
            if (length(center - vec3(4.0f, 0.2f, 0.0f)) > 0.9f) {
                if (chooseMaterial < 0.8f) {
                    float b1 = prng.in01(), b0 = prng.in01();
                    float g1 = prng.in01(), g0 = prng.in01();
                    float r1 = prng.in01(), r0 = prng.in01();
                    material = scene.take(new MaterialLambertian(vec3(r0 * r1, g0 * g1, b0 * b1)));
                    scene.take(new Sphere(center, 0.2f, material));
                } else if (chooseMaterial < 0.95f) {
                    float rough = prng.in01();
                    float ab = prng.in01();
                    float ag = prng.in01();
                    float ar = prng.in01();
                    material = scene.take(new MaterialGGX(
                                vec3(0.5f * (1.0f + ar), 0.5f * (1.0f + ag), 0.5f * (1.0f + ab)),
                                vec2(0.001f + 0.499f * rough)));
                    scene.take(new Sphere(center, 0.2f, material));
                } else {
                    material = scene.take(new MaterialGlass(MaterialGlass::transparentColorToAbsorption(vec3(1.0f)), 1.5f));
                    scene.take(new Sphere(center, 0.2f, material));
                }
            }
        }
    }

    material = scene.take(new MaterialGlass(MaterialGlass::transparentColorToAbsorption(vec3(1.0f)), 1.5f));
    scene.take(new Sphere(vec3(0.0f, 1.0f, 0.0f), 1.0f, material));
    material = scene.take(new MaterialLambertian(vec3(0.4f, 0.2f, 0.1f)));
    scene.take(new Sphere(vec3(-4.0f, 1.0f, 0.0f), 1.0f, material));
    material = scene.take(new MaterialGGX(vec3(0.7f, 0.6f, 0.5f), vec2(0.001f)));
    scene.take(new Sphere(vec3(4.0f, 1.0f, 0.0f), 1.0f, material));

    // sky: emissive sphere seen from inside (the front material is never used)
    material = scene.take(new LightDiffuse(vec3(1.0f)));
    material = scene.take(new MaterialTwoSided(nullptr, material));
    scene.take(new Sphere(vec3(0.0f), 2000.0f, material));
}

/* ---- C3 / C4: procedural ~1M-triangle courtyard ---- */

/* HDR sky: vertical gradient plus a small sun disk whose peak is about 5e3, so that
 * importance sampling matters. Row 0 is the bottom of the image (v = 0 <-> latitude -90). */
inline TGD::Array<float> makeSkyImage(unsigned int w, unsigned int h)
{
    TGD::Array<float> img({ size_t(w), size_t(h) }, 3);
    const float sunU = 0.62f, sunV = 0.70f;     // sun position in texture space
    const float sunRadius = 0.012f;
    for (unsigned int y = 0; y < h; y++) {
        float v = (y + 0.5f) / h;
        for (unsigned int x = 0; x < w; x++) {
            float u = (x + 0.5f) / w;
            float sky = v < 0.5f ? 0.05f + 0.3f * v : 0.2f + 1.6f * (v - 0.5f);
            float r = 0.6f * sky, g = 0.75f * sky, b = 1.0f * sky;
            if (v < 0.5f) {     // below the horizon: dim brownish ground glow
                r = 0.12f * sky / 0.2f; g = 0.1f * sky / 0.2f; b = 0.08f * sky / 0.2f;
            }
            float du = (u - sunU) * 2.0f, dv = (v - sunV);
            float d2 = (du * du + dv * dv) / (sunRadius * sunRadius);
            if (d2 < 9.0f) {
                float s = 5000.0f * std::exp(-2.0f * d2);
                r += s; g += 0.92f * s; b += 0.8f * s;
            }
            float* p = img[{ size_t(x), size_t(y) }];
            p[0] = r; p[1] = g; p[2] = b;
        }
    }
    return img;
}

/* `scale` multiplies the tessellation (1.0 ~ 1.0M triangles, 0.1 ~ 10k). */
inline void buildSynth(Scene& scene, float scale = 1.0f, int envWidth = 2048, int isN = 512)
{
    Prng prng(2024);
    auto tess = [scale](int n) { int r = int(n * std::sqrt(scale) + 0.5f); return r < 4 ? 4 : r; };
    auto ggx = [&scene, &prng]() -> Material* {
        float r = prng.in01(), g = prng.in01(), b = prng.in01();
        float rough = 0.05f + 0.45f * prng.in01();
        return scene.take(new MaterialGGX(vec3(0.3f + 0.65f * r, 0.3f + 0.65f * g, 0.3f + 0.65f * b), vec2(rough)));
    };
    const quat upright = toQuat(radians(-90.0f), vec3(1.0f, 0.0f, 0.0f));

    // ground
    {
        Transformation T(vec3(0.0f, 0.0f, 0.0f), upright, vec3(14.0f, 14.0f, 1.0f));
        scene.take(new MeshInstance(scene.take(generateQuad(T, tess(8))), ggx()));
    }
    // two colonnades of 10 columns each, with arches (tori) between neighbours
    for (int row = 0; row < 2; row++) {
        float x = (row == 0 ? -3.0f : 3.0f);
        for (int i = 0; i < 10; i++) {
            float z = -9.0f + 2.0f * i;
            Transformation base(vec3(x, 0.15f, z), quat::null(), vec3(0.42f, 0.15f, 0.42f));
            scene.take(new MeshInstance(scene.take(generateCube(base, tess(14))), ggx()));
            // column shafts are stretched spheres: evenly shaped triangles. (Very thin slivers such
            // as a 1000-slice cylinder with one stack drive the reference's SAH sweep into trees
            // deeper than its 128-level limit, bvh.hpp:230,260-263.)
            Transformation shaft(vec3(x, 1.8f, z), quat::null(), vec3(0.28f, 1.5f, 0.28f));
            scene.take(new MeshInstance(scene.take(generateSphere(shaft, tess(100), tess(60))), ggx()));
            Transformation cap(vec3(x, 3.4f, z), quat::null(), vec3(0.4f, 0.1f, 0.4f));
            scene.take(new MeshInstance(scene.take(generateCube(cap, tess(10))), ggx()));
            if (i < 9) {
                Transformation arch(vec3(x, 3.5f, z + 1.0f),
                        toQuat(radians(90.0f), vec3(0.0f, 1.0f, 0.0f)), vec3(1.0f, 1.0f, 0.35f));
                scene.take(new MeshInstance(scene.take(generateTorus(arch, 0.72f, tess(90), tess(90))), ggx()));
            }
        }
    }
    // statues: spheres on pedestals along the middle
    for (int i = 0; i < 8; i++) {
        float z = -7.0f + 2.0f * i;
        float x = (i % 2 == 0 ? -0.9f : 0.9f);
        Transformation ped(vec3(x, 0.3f, z), toQuat(radians(20.0f * i), vec3(0.0f, 1.0f, 0.0f)), vec3(0.3f));
        scene.take(new MeshInstance(scene.take(generateCube(ped, tess(12))), ggx()));
        Transformation ball(vec3(x, 0.95f, z), quat::null(), vec3(0.35f));
        scene.take(new MeshInstance(scene.take(generateSphere(ball, tess(170), tess(85))), ggx()));
    }
    // centre piece: a big lying torus and a cone
    {
        Transformation T(vec3(0.0f, 0.45f, 0.5f), upright, vec3(1.6f, 1.6f, 1.2f));
        scene.take(new MeshInstance(scene.take(generateTorus(T, 0.55f, tess(240), tess(240))), ggx()));
        Transformation C(vec3(0.0f, 1.6f, 0.5f), toQuat(radians(35.0f), vec3(0.0f, 0.0f, 1.0f)), vec3(0.45f, 0.45f, 0.3f));
        scene.take(new MeshInstance(scene.take(generateTorus(C, 0.3f, tess(150), tess(150))), ggx()));
    }
    // back wall
    {
        Transformation T(vec3(0.0f, 2.5f, -11.0f), quat::null(), vec3(6.0f, 2.5f, 1.0f));
        scene.take(new MeshInstance(scene.take(generateQuad(T, tess(8))), ggx()));
    }

    // HDR environment with importance sampling; no hot spots, so the envmap NEE branch runs
    TGD::Array<float> sky = makeSkyImage(envWidth, envWidth / 2);
    Texture* skyTex = scene.take(createTextureImage(sky));
    EnvironmentMap* env = scene.take(new EnvironmentMapEquiRect(skyTex));
    if (isN > 0)
        env->initializeImportanceSampling(isN);
}

/* ---- C5a: AMCW time-of-flight example ---- */

inline void buildTof(Scene& scene)
{
    Material* bgMaterial = scene.take(new MaterialLambertian(vec4(1.0f)));
    Transformation bgT(vec3(0.0f, 0.0f, -2.0f), quat::null(), vec3(5.0f));
    scene.take(new MeshInstance(scene.take(generateQuad(bgT)), bgMaterial));

    Material* quadMaterial = scene.take(new MaterialModPhong(vec3(0.7f), vec3(0.3f), 100.0f));
    Transformation quadT(vec3(0.0f), quat::null(), vec3(0.2f));
    int quadAnim = scene.take(new AnimationKeyframes(
                0.0f, Transformation(vec3(-1.0f, 0.5f, -1.5f)),
                5.0f, Transformation(vec3(+1.0f, 0.5f, -1.5f))));
    scene.take(new MeshInstance(scene.take(generateQuad(quadT)), quadMaterial, quadAnim));

    Material* objectMaterial = scene.take(new MaterialModPhong(vec3(0.5f), vec3(0.5f), 100.0f));
    Transformation objectT(vec3(0.0f), quat::null(), vec3(0.33f));
    AnimationKeyframes* objectAnimation = new AnimationKeyframes();
    for (int k = 0; k < 5; k++)
        objectAnimation->addKeyframe(1.25f * k, Transformation(vec3(0.0f, -0.3f, -1.0f),
                    toQuat(radians(60.0f * k), vec3(0.0f, 1.0f, 0.5f))));
    int objectAnim = scene.take(objectAnimation);
    scene.take(new MeshInstance(scene.take(generateIcosahedron(objectT)), objectMaterial, objectAnim));

    Material* lightFront = scene.take(new LightTof(40.0f / (4.0f * pi), radians(120.0f)));
    Material* lightBack = scene.take(new MaterialLambertian(vec4(0.0f)));
    Material* lightMaterial = scene.take(new MaterialTwoSided(lightFront, lightBack));
    Transformation lightT(vec3(0.0f), toQuat(radians(180.0f), vec3(1.0f, 0.0f, 0.0f)), vec3(0.10f, 0.05f, 1.0f));
    scene.take(new MeshInstance(scene.take(generateQuad(lightT)), lightMaterial), HotSpot);
}

/* ---- C5b: participating media in a Cornell box ---- */

inline void buildMedia(Scene& scene)
{
    Material* white = scene.take(new MaterialLambertian(vec3(0.725f, 0.71f, 0.68f)));
    Material* red = scene.take(new MaterialLambertian(vec3(0.63f, 0.065f, 0.05f)));
    Material* green = scene.take(new MaterialLambertian(vec3(0.14f, 0.45f, 0.091f)));
    Material* light = scene.take(new MaterialTwoSided(
                scene.take(new LightDiffuse(vec3(4.0f))),
                scene.take(new MaterialLambertian(vec3(0.0f)))));

    Material* wallMaterials[5] = { red, green, white, white, white };
    for (int i = 0; i < 5; i++)
        scene.take(new MeshInstance(scene.take(meshFromQuad(cornellWalls[i])), wallMaterials[i]));

    // short object: glass sphere filled with a thin blue medium
    std::vector<const Hitable*> shortBoundary = scene.take(new Sphere(vec3(0.33f, 0.3f, 0.37f), 0.3f,
                scene.take(new MaterialGlass(MaterialGlass::transparentColorToAbsorption(vec3(1.0f, 1.0f, 1.0f)), 1.5f))));
    Material* shortPhase = scene.take(new MaterialPhaseFunctionIsotropic(vec3(0.2f, 0.4f, 0.9f)));
    scene.take(new Medium(shortBoundary, 8.0f, shortPhase));

    // tall object: dense medium bounded by an invisible icosahedron
    Transformation tallT(vec3(-0.345f, 0.8f, -0.29f), toQuat(radians(160.0f), vec3(0.0f, 1.0f, 0.0f)), vec3(0.5f));
    std::vector<const Hitable*> tallBoundary = scene.takeUnhitable(
            new MeshInstance(scene.take(generateIcosahedron(tallT)), nullptr));
    Material* tallPhase = scene.take(new MaterialPhaseFunctionIsotropic(vec3(0.8f, 0.8f, 1.0f)));
    scene.take(new Medium(tallBoundary, 250.0f, tallPhase));

    scene.take(new MeshInstance(scene.take(meshFromQuad(cornellLight)), light), HotSpot);
}

/* ---- furnace test ---- */

inline void buildFurnace(Scene& scene)
{
    Texture* tex = scene.take(new TextureConstant(vec4(1.0f)));
    scene.take(new EnvironmentMapEquiRect(tex));
    Material* mat = scene.take(new MaterialLambertian(vec3(0.42f)));
    scene.take(new Sphere(vec3(0.0f), 1.0f, mat));
}

/* ---- MIS test (Veach plates) ---- */

inline void buildMis(Scene& scene, bool withHotSpots)
{
    Material* white = scene.take(new MaterialLambertian(vec4(0.8f)));
    struct Wall { float tx, ty, tz; float angle; float ax, ay, az; };
    const Wall walls[6] = {
        { -2.6f, 0.0f, 0.0f, +90.0f, 0.0f, 1.0f, 0.0f },
        { +2.6f, 0.0f, 0.0f, -90.0f, 0.0f, 1.0f, 0.0f },
        { 0.0f, 0.0f, +5.0f, 180.0f, 0.0f, 1.0f, 0.0f },
        { 0.0f, 0.0f, -4.6f, 0.0f, 0.0f, 1.0f, 0.0f },
        { 0.0f, -2.499, 0.0f, +90.0f, 1.0f, 0.0f, 0.0f },
        { 0.0f, -5.0f, 0.0f, -90.0f, 1.0f, 0.0f, 0.0f }
    };
    Mesh* wallMeshes[6];
    for (int i = 0; i < 6; i++) {
        Transformation T;
        T.translate(vec3(walls[i].tx, walls[i].ty, walls[i].tz));
        T.scale(vec3(5.0f));
        T.rotate(toQuat(radians(walls[i].angle), vec3(walls[i].ax, walls[i].ay, walls[i].az)));
        wallMeshes[i] = scene.take(generateQuad(T));
    }
    for (int i = 0; i < 6; i++)
        scene.take(new MeshInstance(wallMeshes[i], white));

    const float roughness[4] = { 0.001f, 0.008, 0.03f, 0.1f };
    Material* plateMaterials[4];
    for (int i = 0; i < 4; i++)
        plateMaterials[i] = scene.take(new MaterialGGX(vec3(1.0f), vec2(roughness[i])));
    const float plateY[4] = { -4.2f, -4.6f, -4.8f, -4.9f };
    const float plateZ[4] = { -4.2f, -3.8f, -3.4f, -3.0f };
    const float plateAngle[4] = { -35.0f, -47.0f, -59.0f, -71.0f };
    for (int i = 0; i < 4; i++) {
        Transformation T(vec3(0.0f, plateY[i], plateZ[i]),
                toQuat(radians(plateAngle[i]), vec3(1.0f, 0.0f, 0.0f)), vec3(2.0f, 0.3f, 1.0f));
        scene.take(new MeshInstance(scene.take(generateQuad(T)), plateMaterials[i]));
    }

    const float lightX[4] = { -1.5f, -0.5f, +0.5f, +1.5f };
    const float lightR[4] = { 0.032f, 0.08f, 0.2f, 0.5f };
    Material* lights[4];
    for (int i = 0; i < 4; i++)
        lights[i] = scene.take(new LightDiffuse(vec3(4.0f)));
    for (int i = 0; i < 4; i++) {
        Transformation T(vec3(lightX[i], -3.5f, -4.0f), quat::null(), vec3(lightR[i]));
        scene.take(new Sphere(lights[i], T), withHotSpots ? HotSpot : ColdSpot);
    }
}

/* ---- a small mixed scene exercising textures, normal maps, mirror, spot light, cube envmap ---- */

inline void buildMixed(Scene& scene)
{
    // checker floor with an image-textured, normal-mapped wall and assorted objects
    Texture* checker = scene.take(new TextureChecker(vec3(0.8f), vec3(0.2f, 0.25f, 0.3f), 8, 8));
    Material* floorMat = scene.take(new MaterialLambertian(vec3(0.5f), checker));
    const quat upright = toQuat(radians(-90.0f), vec3(1.0f, 0.0f, 0.0f));
    scene.take(new MeshInstance(scene.take(generateQuad(Transformation(vec3(0.0f), upright, vec3(4.0f)))), floorMat));

    TGD::Array<uint8_t> img({ size_t(16), size_t(16) }, 3);
    TGD::Array<uint8_t> nrm({ size_t(16), size_t(16) }, 3);
    for (size_t y = 0; y < 16; y++) {
        for (size_t x = 0; x < 16; x++) {
            uint8_t* p = img[{ x, y }];
            p[0] = uint8_t(40 + 13 * x); p[1] = uint8_t(220 - 11 * y); p[2] = uint8_t(((x ^ y) & 1) ? 200 : 60);
            uint8_t* q = nrm[{ x, y }];
            q[0] = uint8_t(128 + ((x & 3) - 1) * 20); q[1] = uint8_t(128 + ((y & 3) - 1) * 20); q[2] = 230;
        }
    }
    Texture* imgTex = scene.take(createTextureImage(img));
    Texture* nrmTex = scene.take(createTextureImage(nrm, LinearizeSRGB_Off));
    Texture* imgTexScaled = scene.take(new TextureTransformer(imgTex, vec2(2.0f, 2.0f), vec2(0.25f, 0.0f),
                vec4(0.9f), vec4(0.05f)));
    Material* wallMat = scene.take(new MaterialLambertian(vec3(0.5f), imgTexScaled));
    wallMat->normalTex = nrmTex;
    scene.take(new MeshInstance(scene.take(generateQuad(Transformation(vec3(0.0f, 2.0f, -3.0f), quat::null(), vec3(4.0f, 2.0f, 1.0f)))), wallMat));

    Material* mirror = scene.take(new MaterialMirror(vec3(0.9f)));
    scene.take(new MeshInstance(scene.take(generateCube(Transformation(vec3(-1.6f, 0.5f, -1.0f),
                            toQuat(radians(30.0f), vec3(0.0f, 1.0f, 0.0f)), vec3(0.5f)))), mirror));
    Material* phong = scene.take(new MaterialModPhong(vec3(0.5f, 0.3f, 0.2f), vec3(0.3f), 40.0f, 0.7f));
    scene.take(new Sphere(vec3(0.0f, 0.6f, -0.5f), 0.6f, phong));
    Material* glass = scene.take(new MaterialGlass(vec3(0.3f, 0.1f, 0.1f), vec3(1.50f, 1.52f, 1.54f)));
    scene.take(new MeshInstance(scene.take(generateIcosahedron(Transformation(vec3(1.5f, 0.7f, -0.6f), quat::null(), vec3(0.6f)))), glass));
    Material* ggxTex = scene.take(new MaterialGGX(vec3(0.9f), vec2(0.2f, 0.05f), imgTex));
    scene.take(new MeshInstance(scene.take(generateTorus(Transformation(vec3(0.0f, 0.25f, 1.0f), upright, vec3(0.8f, 0.8f, 0.6f)),
                        0.5f, 24, 24)), ggxTex));

    Material* spot = scene.take(new MaterialTwoSided(
                scene.take(new LightSpot(radians(100.0f), vec3(30.0f))),
                scene.take(new MaterialLambertian(vec3(0.0f)))));
    scene.take(new MeshInstance(scene.take(generateQuad(Transformation(vec3(0.0f, 3.5f, 0.0f),
                            toQuat(radians(90.0f), vec3(1.0f, 0.0f, 0.0f)), vec3(0.3f)))), spot), HotSpot);
    Material* bulb = scene.take(new LightDiffuse(vec3(6.0f, 5.0f, 4.0f)));
    scene.take(new Sphere(vec3(-2.5f, 2.0f, 1.0f), 0.25f, bulb), HotSpot);

    // cube environment from six constant/checker faces
    Texture* faces[6];
    const float faceColors[6][3] = { { 0.3f, 0.1f, 0.1f }, { 0.1f, 0.3f, 0.1f }, { 0.2f, 0.3f, 0.5f },
        { 0.05f, 0.05f, 0.05f }, { 0.1f, 0.1f, 0.3f }, { 0.3f, 0.3f, 0.1f } };
    for (int i = 0; i < 6; i++)
        faces[i] = scene.take(new TextureChecker(vec3(faceColors[i][0], faceColors[i][1], faceColors[i][2]),
                    vec3(0.02f), 4, 4));
    scene.take(new EnvironmentMapCube(faces[0], faces[1], faces[2], faces[3], faces[4], faces[5]));
}

/* ---- an OBJ/MTL file written on the fly and imported: exercises importIntoScene() (import.hpp) ---- */

inline void buildImported(Scene& scene)
{
    /* every number below is exactly representable, so that two OBJ readers agree on every bit */
    char dirTemplate[] = "/tmp/wpt_import_XXXXXX";
    const char* dir = mkdtemp(dirTemplate);
    if (!dir)
        return;
    const std::string base(dir);
    {
        /* a 4x4 rgb texture in the PPM format the libtgd stand-in reads */
        FILE* f = std::fopen((base + "/checks.ppm").c_str(), "wb");
        std::fprintf(f, "P6\n4 4\n255\n");
        for (int i = 0; i < 16; i++) {
            unsigned char px[3] = { (unsigned char)((i % 2) ? 224 : 64), (unsigned char)((i / 4) % 2 ? 192 : 96), 128 };
            std::fwrite(px, 1, 3, f);
        }
        std::fclose(f);
    }
    {
        FILE* f = std::fopen((base + "/room.mtl").c_str(), "w");
        std::fputs(
            "newmtl floor\nKd 0.5 0.5 0.5\n\n"
            "newmtl painted\nKd 0.75 0.25 0.25\nKs 0.125 0.125 0.125\nNs 32\nmap_Kd -s 2 2 1 checks.ppm\n\n"
            "newmtl lamp\nKd 0 0 0\nKe 8 8 6\n\n"
            "newmtl pane\nKd 0.5 0.75 0.5\nd 0.25\nNi 1.5\n", f);
        std::fclose(f);
    }
    {
        FILE* f = std::fopen((base + "/room.obj").c_str(), "w");
        std::fputs(
            "mtllib room.mtl\n"
            "v -2 0 -2\nv 2 0 -2\nv 2 0 2\nv -2 0 2\n"                 /* 1-4 floor */
            "v -1 0 -1\nv 0 0 -1\nv 0 1 -1\nv -1 1 -1\n"                /* 5-8 box front (quad) */
            "v -1 0 -1.5\nv 0 0 -1.5\nv 0 1 -1.5\nv -1 1 -1.5\n"        /* 9-12 box back */
            "v -0.5 2.5 -0.5\nv 0.5 2.5 -0.5\nv 0.5 2.5 0.5\nv -0.5 2.5 0.5\n"  /* 13-16 lamp */
            "v 0.5 0 0.5\nv 1.5 0 0.5\nv 1 1.5 0.5\n"                    /* 17-19 pane */
            "v 1.25 0 -0.75\nv 1.75 0 -0.75\nv 1.5 0.5 -0.75\n"          /* 20-22 loose triangle, no material */
            "vn 0 1 0\nvn 0 0 1\nvn 0 -2 0\n"
            "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
            "o floor\nusemtl floor\nf 1//1 4//1 3//1 2//1\n"
            "o box\nusemtl painted\nf 5/1/2 6/2/2 7/3/2 8/4/2\nf 6/1 10/2 11/3 7/4\nf -12/1 -14/2 -13/3\n"
            "g lamp\nusemtl lamp\nf 13//3 14//3 15//3 16//3\n"
            "g pane\nusemtl pane\nf 17 18 19\n"
            "usemtl nosuchmaterial\nf 20 21 22\n", f);
        std::fclose(f);
    }
    /* the reference imports materials in an OpenMP loop and takes the meshes in completion order: one thread
     * makes that order the file's order (this header is compiled against the reference as well) */
#ifdef _OPENMP
    const int threadsBefore = omp_get_max_threads();
    omp_set_num_threads(1);
#endif
    importIntoScene(scene, base + "/room.obj", Transformation(vec3(0.0f, 0.0f, -0.5f)));
#ifdef _OPENMP
    omp_set_num_threads(threadsBefore);
#endif
    std::remove((base + "/room.obj").c_str());
    std::remove((base + "/room.mtl").c_str());
    std::remove((base + "/checks.ppm").c_str());
    rmdir(dir);
}

/* ---- registry ---- */

struct View {
    unsigned int width = 512, height = 512;
    unsigned int samplesSqrt = 8;
    float t0 = 0.0f, t1 = 0.0f;
    Parameters params;
    Camera::SurroundMode surround = Camera::Surround_Off;
    float stereoscopicDistance = 0.0f;
    float vfovDegrees = 50.0f;
    float aperture = 0.0f, focusDistance = 1.0f;
    Transformation pose;
    bool tof = false;
};

inline bool buildScene(const std::string& name, Scene& scene, float scale = 1.0f)
{
    if (name == "cornell") buildCornell(scene);
    else if (name == "rtiow") buildRtiow(scene);
    else if (name == "synth" || name == "synth360") buildSynth(scene, scale, scale >= 0.5f ? 2048 : 256, scale >= 0.5f ? 512 : 64);
    else if (name == "tof") buildTof(scene);
    else if (name == "media") buildMedia(scene);
    else if (name == "furnace") buildFurnace(scene);
    else if (name == "mis") buildMis(scene, true);
    else if (name == "mis_nohot") buildMis(scene, false);
    else if (name == "mixed") buildMixed(scene);
    else if (name == "imported") buildImported(scene);
    else return false;
    return true;
}

inline View defaultView(const std::string& name)
{
    View v;
    if (name == "cornell" || name == "media") {
        v.width = v.height = 512;
        v.samplesSqrt = (name == "cornell" ? 8 : 16);
        v.vfovDegrees = 50.0f;
        v.pose = Transformation::fromLookAt(vec3(0.0f, 1.0f, 3.2f), vec3(0.0f, 1.0f, -1.0f), vec3(0.0f, 1.0f, 0.0f));
    } else if (name == "rtiow") {
        v.width = 1920; v.height = 1080; v.samplesSqrt = 16;
        v.vfovDegrees = 30.0f;
        v.aperture = 0.06f; v.focusDistance = 7.0f;
        v.pose = Transformation::fromLookAt(vec3(10.0f, 2.0f, 3.0f), vec3(0.0f, 0.0f, -1.0f));
        v.params.minHitDistance = 0.00035;
    } else if (name == "synth") {
        v.width = 1920; v.height = 1080; v.samplesSqrt = 16;
        v.vfovDegrees = 55.0f;
        v.pose = Transformation::fromLookAt(vec3(0.4f, 1.7f, 9.5f), vec3(0.0f, 1.2f, 0.0f));
    } else if (name == "synth360") {
        v.width = 4096; v.height = 4096; v.samplesSqrt = 16;
        v.surround = Camera::Surround_360;
        v.stereoscopicDistance = 0.07f;
        v.pose = Transformation::fromLookAt(vec3(0.4f, 1.7f, 5.0f), vec3(0.0f, 1.7f, 0.0f));
    } else if (name == "tof") {
        v.width = 352; v.height = 288; v.samplesSqrt = 10;
        v.vfovDegrees = 70.0f;
        v.params.maxPathComponents = 2;
        v.params.rrThreshold = 0.0f;
        v.tof = true;
        v.t0 = 0.0f; v.t1 = 1000.0f / 1e6f;
    } else if (name == "furnace") {
        v.width = v.height = 512; v.samplesSqrt = 10;
        v.vfovDegrees = 40.0f;
        v.pose = Transformation(vec3(0.0f, 0.0f, 5.0f));
        v.params.randomizeRayOverPixel = false;
    } else if (name == "mis" || name == "mis_nohot") {
        v.width = 960; v.height = 540; v.samplesSqrt = 10;
        v.vfovDegrees = 50.0f;
        v.pose = Transformation(vec3(0.0f, -4.5f, -1.2f));
    } else if (name == "imported") {
        v.width = 320; v.height = 240; v.samplesSqrt = 8;
        v.vfovDegrees = 50.0f;
        v.pose = Transformation::fromLookAt(vec3(0.0f, 1.5f, 4.5f), vec3(0.0f, 0.8f, -1.0f));
    } else if (name == "mixed") {
        v.width = 480; v.height = 320; v.samplesSqrt = 8;
        v.vfovDegrees = 55.0f;
        v.pose = Transformation::fromLookAt(vec3(0.5f, 1.8f, 5.0f), vec3(0.0f, 0.8f, 0.0f));
    }
    return v;
}

inline Camera makeCamera(const View& v, float aspectRatio)
{
    Optics optics(Projection(radians(v.vfovDegrees), aspectRatio), LensDistortion(),
            LensDepthOfField(v.aperture, v.focusDistance));
    return Camera(v.surround, v.stereoscopicDistance, optics, v.pose);
}

}
