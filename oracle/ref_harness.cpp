/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, never part of the product path.
 *
 * Builds the *unmodified* WurblPT reference (headers under /root/reference/libwurblpt,
 * compiled from where they lie; see oracle/Makefile) into oracle/_ref/libwpt_ref_*.so and
 * exposes, through a small C API for ctypes:
 *   - scene construction (scenes/configs.hpp compiled against the reference headers, or the
 *     reference example apps' own createScene() functions),
 *   - flattening of a reference Scene into the C-ABI scene description (wurblpt_b200.h), so
 *     the device path can be fed exactly the data the reference renders,
 *   - the reference's own results on that scene: mcpt() frames (wurblpt.hpp:279-436), the
 *     primary-ray closest-hit query of getGroundTruth() (wurblpt.hpp:659-668), closest hits
 *     for arbitrary rays (bvh.hpp:277-311), Prng values (prng.hpp:79-95).
 *
 * Compiled with -fno-access-control so that private members of reference classes can be
 * read while flattening; no reference source is copied or modified.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <vector>

#include <tgd/array.hpp>
#include <tgd/io.hpp>
#include <wurblpt/wurblpt.hpp>

#include "wurblpt_b200.h"
#include "../scenes/configs.hpp"

#ifdef WPT_REF_COUNTERS
#include "ref_counters.hpp"
WptRefCounters wpt_ref_counters[WPT_REF_MAX_THREADS];
#endif

/* The reference example apps, compiled from their own sources. Each app defines
 * createScene()/main() at global scope, so each goes into its own namespace. */
namespace app_cornell {
#define main app_main
#include "wurblpt-cornellbox/wurblpt-cornellbox.cpp"
#undef main
}
namespace app_rtiow {
#define main app_main
#include "wurblpt-rtiow/wurblpt-rtiow.cpp"
#undef main
}
namespace app_tof {
#define main app_main
#include "wurblpt-tof-example/wurblpt-tof-example.cpp"
#undef main
}
namespace app_media {
#define main app_main
#include "wurblpt-participating-media/wurblpt-participating-media.cpp"
#undef main
}
namespace app_mis {
#define main app_main
#include "wurblpt-mis-test/wurblpt-mis-test.cpp"
#undef main
}

using namespace WurblPT;

namespace {

/* ------------------------------------------------------------------ flatten */

struct FlatScene {
    std::vector<wpt_animation> animations;
    std::vector<wpt_keyframe> keyframes;
    std::vector<wpt_texture> textures;
    std::vector<uint8_t> texels;
    std::vector<wpt_material> materials;
    std::vector<wpt_mesh> meshes;
    std::vector<float> vertices;
    std::vector<uint32_t> indices;
    std::vector<wpt_instance> instances;
    std::vector<wpt_sphere> spheres;
    std::vector<wpt_medium> media;
    std::vector<uint32_t> boundary;
    std::vector<wpt_component> components;
    std::vector<float> envM;
    std::vector<int32_t> envMs;
    std::vector<float> envMcs;
    wpt_scene_desc desc;

    std::unordered_map<const Texture*, int> texMap;
    std::unordered_map<const Material*, int> matMap;
    std::unordered_map<const Mesh*, int> meshMap;
    std::unordered_map<const Hitable*, int32_t> hitableId;
    std::string error;
};

void setv(float* dst, const vec4& v) { for (int i = 0; i < 4; i++) dst[i] = v[i]; }
void setv(float* dst, const vec3& v) { for (int i = 0; i < 3; i++) dst[i] = v[i]; }
void setv(float* dst, const vec2& v) { for (int i = 0; i < 2; i++) dst[i] = v[i]; }
void setq(float* dst, const quat& q) { dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w; }

wpt_keyframe toKeyframe(float t, const Transformation& T)
{
    wpt_keyframe k;
    k.t = t;
    setv(k.translation, T.translation);
    setq(k.rotation, T.rotation);
    setv(k.scaling, T.scaling);
    return k;
}

int flattenTexture(FlatScene& F, const Texture* tex);

template<int C, TGD::Type TY, bool SRGB>
bool tryImageTexture(FlatScene& F, const Texture* tex, wpt_texture& out)
{
    auto* t = dynamic_cast<const TextureImage<C, TY, SRGB>*>(tex);
    if (!t)
        return false;
    out.type = WPT_TEX_IMAGE;
    out.width = t->_img.dimension(0);
    out.height = t->_img.dimension(1);
    out.components = C;
    out.texel_type = (TY == TGD::uint8 ? WPT_TEXEL_U8 : TY == TGD::uint16 ? WPT_TEXEL_U16 : WPT_TEXEL_F32);
    out.linearize_srgb = SRGB;
    while (F.texels.size() % 16 != 0)
        F.texels.push_back(0);
    out.texel_offset = F.texels.size();
    const uint8_t* src = static_cast<const uint8_t*>(t->_img.data());
    F.texels.insert(F.texels.end(), src, src + t->_img.dataSize());
    setv(out.coord_factor, t->_coordFactor);
    setv(out.coord_offset, t->_coordOffset);
    setv(out.val_factor, t->_valFactor);
    setv(out.val_offset, t->_valOffset);
    return true;
}

template<int C, TGD::Type TY>
bool tryImageTextureS(FlatScene& F, const Texture* tex, wpt_texture& out)
{
    return tryImageTexture<C, TY, true>(F, tex, out) || tryImageTexture<C, TY, false>(F, tex, out);
}

template<int C>
bool tryImageTextureT(FlatScene& F, const Texture* tex, wpt_texture& out)
{
    return tryImageTextureS<C, TGD::uint8>(F, tex, out) || tryImageTextureS<C, TGD::uint16>(F, tex, out)
        || tryImageTextureS<C, TGD::float32>(F, tex, out);
}

int flattenTexture(FlatScene& F, const Texture* tex)
{
    if (!tex)
        return -1;
    auto it = F.texMap.find(tex);
    if (it != F.texMap.end())
        return it->second;
    wpt_texture t;
    std::memset(&t, 0, sizeof(t));
    t.src_texture = -1;
    if (auto* c = dynamic_cast<const TextureConstant*>(tex)) {
        t.type = WPT_TEX_CONSTANT;
        setv(t.color0, c->_color);
    } else if (auto* c = dynamic_cast<const TextureChecker*>(tex)) {
        t.type = WPT_TEX_CHECKER;
        setv(t.color0, c->_color0);
        setv(t.color1, c->_color1);
        t.horiz = c->_horiz;
        t.vert = c->_vert;
    } else if (auto* c = dynamic_cast<const TextureTransformer*>(tex)) {
        t.type = WPT_TEX_TRANSFORMER;
        t.src_texture = flattenTexture(F, c->_texture);
        setv(t.coord_factor, c->_coordFactor);
        setv(t.coord_offset, c->_coordOffset);
        setv(t.val_factor, c->_valFactor);
        setv(t.val_offset, c->_valOffset);
    } else if (tryImageTextureT<1>(F, tex, t) || tryImageTextureT<2>(F, tex, t)
            || tryImageTextureT<3>(F, tex, t) || tryImageTextureT<4>(F, tex, t)) {
    } else {
        F.error = std::string("unsupported texture type ") + typeid(*tex).name();
        return -1;
    }
    int idx = F.textures.size();
    F.textures.push_back(t);
    F.texMap[tex] = idx;
    return idx;
}

int flattenMaterial(FlatScene& F, const Material* mat)
{
    if (!mat)
        return -1;
    auto it = F.matMap.find(mat);
    if (it != F.matMap.end())
        return it->second;
    wpt_material m;
    std::memset(&m, 0, sizeof(m));
    for (int i = 0; i < 5; i++)
        m.tex[i] = -1;
    m.front = m.back = -1;
    m.normal_tex = flattenTexture(F, mat->normalTex);
    // reserve the slot first so that indices follow take() order even for two-sided materials
    int idx = F.materials.size();
    F.materials.push_back(m);
    F.matMap[mat] = idx;
    if (auto* x = dynamic_cast<const MaterialLambertian*>(mat)) {
        m.type = WPT_MAT_LAMBERTIAN;
        setv(m.v0, x->_albedo);
        m.tex[0] = flattenTexture(F, x->_tex);
        if (x->haveNIR) m.flags |= WPT_MATF_HAVE_NIR;
    } else if (auto* x = dynamic_cast<const MaterialMirror*>(mat)) {
        m.type = WPT_MAT_MIRROR;
        setv(m.v0, x->color);
        m.tex[0] = flattenTexture(F, x->colorTex);
        if (x->haveNIR) m.flags |= WPT_MATF_HAVE_NIR;
    } else if (auto* x = dynamic_cast<const MaterialGlass*>(mat)) {
        m.type = WPT_MAT_GLASS;
        setv(m.v0, x->absorption);
        setv(m.v1, x->refractiveIndexOfMaterial);
        setv(m.v2, x->refractiveIndexOfSurroundingMedium);
        if (x->chromaticDispersion) m.flags |= WPT_MATF_CHROMATIC;
    } else if (auto* x = dynamic_cast<const MaterialGGX*>(mat)) {
        m.type = WPT_MAT_GGX;
        setv(m.v0, x->albedo);
        m.f[0] = x->roughness.x();
        m.f[1] = x->roughness.y();
        m.tex[0] = flattenTexture(F, x->albedoTex);
        m.tex[1] = flattenTexture(F, x->roughnessTex);
    } else if (auto* x = dynamic_cast<const MaterialModPhong*>(mat)) {
        m.type = WPT_MAT_MODPHONG;
        setv(m.v0, x->diffuse);
        setv(m.v1, x->specular);
        setv(m.v2, x->transmissive);
        setv(m.v3, x->emissive);
        m.f[0] = x->shininess;
        m.f[1] = x->opacity;
        m.f[2] = x->indexOfRefraction;
        m.tex[0] = flattenTexture(F, x->diffuseTex);
        m.tex[1] = flattenTexture(F, x->specularTex);
        m.tex[2] = flattenTexture(F, x->shininessTex);
        m.tex[3] = flattenTexture(F, x->opacityTex);
        m.tex[4] = flattenTexture(F, x->emissiveTex);
        if (x->haveNIR) m.flags |= WPT_MATF_HAVE_NIR;
        if (x->diffuseTexHasAlpha) m.flags |= WPT_MATF_DIFFUSE_TEX_ALPHA;
        if (x->specularTexHasAlpha) m.flags |= WPT_MATF_SPECULAR_TEX_ALPHA;
    } else if (auto* x = dynamic_cast<const MaterialPhaseFunctionIsotropic*>(mat)) {
        m.type = WPT_MAT_PHASE_ISOTROPIC;
        setv(m.v0, x->_albedo);
    } else if (auto* x = dynamic_cast<const LightDiffuse*>(mat)) {
        m.type = WPT_MAT_LIGHT_DIFFUSE;
        setv(m.v0, x->_emit);
        m.tex[0] = flattenTexture(F, x->_tex);
    } else if (auto* x = dynamic_cast<const LightSpot*>(mat)) {
        m.type = WPT_MAT_LIGHT_SPOT;
        setv(m.v0, x->_emit);
        m.f[0] = x->_cosHalfOpeningAngle;
        m.tex[0] = flattenTexture(F, x->_tex);
    } else if (auto* x = dynamic_cast<const LightTof*>(mat)) {
        m.type = WPT_MAT_LIGHT_TOF;
        m.f[0] = x->_radiance;
        m.f[1] = x->_cosHalfOpeningAngle;
        m.tex[0] = flattenTexture(F, x->_tex);
    } else if (auto* x = dynamic_cast<const MaterialTwoSided*>(mat)) {
        m.type = WPT_MAT_TWOSIDED;
        m.front = flattenMaterial(F, x->_frontSideMaterial);
        m.back = flattenMaterial(F, x->_backSideMaterial);
    } else if (typeid(*mat) == typeid(Material)) {
        m.type = WPT_MAT_NONE;
    } else {
        F.error = std::string("unsupported material type ") + typeid(*mat).name();
        return -1;
    }
    F.materials[idx] = m;
    return idx;
}

int flattenMesh(FlatScene& F, const Mesh* mesh)
{
    auto it = F.meshMap.find(mesh);
    if (it != F.meshMap.end())
        return it->second;
    wpt_mesh m;
    std::memset(&m, 0, sizeof(m));
    m.vertex_offset = F.vertices.size();
    m.index_offset = F.indices.size();
    m.vertex_count = mesh->vertexCount();
    m.vertex_stride = mesh->vertexSize();
    m.triangle_count = mesh->triangleCount();
    m.flags = (mesh->haveTexCoords ? WPT_MESHF_TEXCOORDS : 0) | (mesh->haveTangents ? WPT_MESHF_TANGENTS : 0);
    F.vertices.insert(F.vertices.end(), mesh->vertices.begin(), mesh->vertices.end());
    F.indices.insert(F.indices.end(), mesh->indices.begin(), mesh->indices.end());
    int idx = F.meshes.size();
    F.meshes.push_back(m);
    F.meshMap[mesh] = idx;
    return idx;
}

struct TriInfo {
    const float* vertices;
    uvec3 idx;
    const Material* material;
    int animation;
    bool transform;
    const MeshInstance* instance;   // only known for transformed triangles
};

template<bool TR, bool AN, bool TC, bool TG>
bool tryTriangle(const Hitable* h, TriInfo& o)
{
    auto* t = dynamic_cast<const HitableTriangle<TR, AN, TC, TG>*>(h);
    if (!t)
        return false;
    o.vertices = t->_vertices;
    o.idx = t->indices();
    o.material = t->material();
    o.animation = t->animationIndex();
    o.transform = TR;
    o.instance = TR ? t->instance() : nullptr;
    return true;
}

bool triangleInfo(const Hitable* h, TriInfo& o)
{
    return tryTriangle<false, false, false, false>(h, o) || tryTriangle<false, false, false, true>(h, o)
        || tryTriangle<false, false, true, false>(h, o) || tryTriangle<false, false, true, true>(h, o)
        || tryTriangle<false, true, false, false>(h, o) || tryTriangle<false, true, false, true>(h, o)
        || tryTriangle<false, true, true, false>(h, o) || tryTriangle<false, true, true, true>(h, o)
        || tryTriangle<true, false, false, false>(h, o) || tryTriangle<true, false, false, true>(h, o)
        || tryTriangle<true, false, true, false>(h, o) || tryTriangle<true, false, true, true>(h, o)
        || tryTriangle<true, true, false, false>(h, o) || tryTriangle<true, true, false, true>(h, o)
        || tryTriangle<true, true, true, false>(h, o) || tryTriangle<true, true, true, true>(h, o);
}

/* Append the hitables of one of the scene's lists as components. Consecutive triangles
 * of one mesh instance become one MESH_INSTANCE component. */
bool flattenHitables(FlatScene& F, const Scene& scene,
        const std::vector<std::unique_ptr<Hitable>>& list, bool hitable,
        const std::unordered_map<const Hitable*, bool>& hotSpots,
        std::vector<std::pair<const HitableMedium*, int>>& mediaToResolve)
{
    size_t i = 0;
    while (i < list.size()) {
        const Hitable* h = list[i].get();
        bool hot = hotSpots.count(h) > 0;
        wpt_component comp;
        std::memset(&comp, 0, sizeof(comp));
        comp.flags = (hitable ? WPT_COMPF_HITABLE : 0) | (hot ? WPT_COMPF_HOTSPOT : 0);
        TriInfo ti;
        if (auto* s = dynamic_cast<const HitableSphere*>(h)) {
            wpt_sphere sp;
            std::memset(&sp, 0, sizeof(sp));
            setv(sp.center, s->_center);
            sp.radius = s->_radius;
            setq(sp.rotation, s->_rotation);
            sp.material = flattenMaterial(F, s->_material);
            sp.animation = s->_animationIndex;
            comp.type = WPT_COMP_SPHERE;
            comp.ref = F.spheres.size();
            F.spheres.push_back(sp);
            F.hitableId[h] = F.hitableId.size();
            i++;
        } else if (auto* m = dynamic_cast<const HitableMedium*>(h)) {
            wpt_medium md;
            std::memset(&md, 0, sizeof(md));
            md.density = m->_density;
            md.material = flattenMaterial(F, m->_material);
            comp.type = WPT_COMP_MEDIUM;
            comp.ref = F.media.size();
            mediaToResolve.push_back(std::make_pair(m, int(F.media.size())));
            F.media.push_back(md);
            F.hitableId[h] = F.hitableId.size();
            i++;
        } else if (triangleInfo(h, ti)) {
            const Mesh* mesh = nullptr;
            for (auto& mp : scene._meshes)
                if (mp->vertices.data() == ti.vertices)
                    mesh = mp.get();
            if (!mesh) {
                F.error = "triangle whose mesh is not owned by the scene";
                return false;
            }
            wpt_instance inst;
            std::memset(&inst, 0, sizeof(inst));
            inst.mesh = flattenMesh(F, mesh);
            inst.material = flattenMaterial(F, ti.material);
            inst.animation = ti.animation;
            inst.has_transform = ti.transform;
            mat4 M(1.0f);
            mat3 N(1.0f);
            if (ti.transform) {
                M = ti.instance->transformationM;
                N = ti.instance->transformationN;
            }
            for (int c = 0; c < 4; c++)
                for (int r = 0; r < 4; r++)
                    inst.M[c * 4 + r] = M[c][r];
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 3; r++)
                    inst.N[c * 3 + r] = N[c][r];
            comp.type = WPT_COMP_MESH_INSTANCE;
            comp.ref = F.instances.size();
            F.instances.push_back(inst);
            // consume the triangles of this instance, checking they are the mesh's triangles in order
            size_t n = mesh->triangleCount();
            for (size_t k = 0; k < n; k++) {
                if (i + k >= list.size()) {
                    F.error = "truncated mesh instance";
                    return false;
                }
                TriInfo tk;
                const Hitable* hk = list[i + k].get();
                if (!triangleInfo(hk, tk) || tk.vertices != ti.vertices || tk.material != ti.material
                        || tk.instance != ti.instance || tk.animation != ti.animation
                        || tk.idx[0] != mesh->indices[3 * k + 0] || tk.idx[1] != mesh->indices[3 * k + 1]
                        || tk.idx[2] != mesh->indices[3 * k + 2] || (hotSpots.count(hk) > 0) != hot) {
                    F.error = "mesh instance triangles are not contiguous";
                    return false;
                }
                F.hitableId[hk] = F.hitableId.size();
            }
            i += n;
        } else {
            F.error = std::string("unsupported hitable type ") + typeid(*h).name();
            return false;
        }
        F.components.push_back(comp);
        if (!F.error.empty())
            return false;
    }
    return true;
}

bool flattenScene(FlatScene& F, const Scene& scene)
{
    // animations
    for (auto& a : scene._animations) {
        auto* kf = dynamic_cast<const AnimationKeyframes*>(a.get());
        if (!kf) {
            F.error = "animation is not an AnimationKeyframes";
            return false;
        }
        wpt_animation wa;
        wa.first_keyframe = F.keyframes.size();
        wa.keyframe_count = kf->_keyframes.size();
        for (auto& k : kf->_keyframes)
            F.keyframes.push_back(toKeyframe(k.t, k.transformation));
        F.animations.push_back(wa);
    }
    // materials in take() order (scene.hpp:105-111), so indices equal Scene::materialIndex()
    {
        std::vector<const Material*> ordered(scene._materialMap.size(), nullptr);
        for (auto& kv : scene._materialMap)
            ordered[kv.second] = kv.first;
        for (const Material* m : ordered)
            flattenMaterial(F, m);
        if (!F.error.empty())
            return false;
    }
    std::unordered_map<const Hitable*, bool> hotSpots;
    for (const Hitable* h : scene._hotSpots)
        hotSpots[h] = true;
    std::vector<std::pair<const HitableMedium*, int>> mediaToResolve;
    if (!flattenHitables(F, scene, scene._hitables, true, hotSpots, mediaToResolve))
        return false;
    if (!flattenHitables(F, scene, scene._unhitables, false, hotSpots, mediaToResolve))
        return false;
    for (auto& mr : mediaToResolve) {
        wpt_medium& md = F.media[mr.second];
        md.first_boundary = F.boundary.size();
        md.boundary_count = mr.first->_boundaryHitables.size();
        for (const Hitable* b : mr.first->_boundaryHitables) {
            auto it = F.hitableId.find(b);
            if (it == F.hitableId.end()) {
                F.error = "medium boundary hitable not owned by the scene";
                return false;
            }
            F.boundary.push_back(it->second);
        }
    }
    // environment map
    wpt_envmap env;
    std::memset(&env, 0, sizeof(env));
    for (int i = 0; i < 6; i++)
        env.tex[i] = -1;
    if (const EnvironmentMap* e = scene.environmentMap()) {
        if (auto* er = dynamic_cast<const EnvironmentMapEquiRect*>(e)) {
            env.type = WPT_ENV_EQUIRECT;
            env.compat = (er->_compatibility == EnvironmentMapEquiRect::CompatibilityMitsuba
                    ? WPT_ENV_COMPAT_MITSUBA : WPT_ENV_COMPAT_SURROUND_VIDEO);
            env.tex[0] = flattenTexture(F, er->_tex);
        } else if (auto* ec = dynamic_cast<const EnvironmentMapCube*>(e)) {
            env.type = WPT_ENV_CUBE;
            for (int i = 0; i < 6; i++)
                env.tex[i] = flattenTexture(F, ec->_cubesides[i]);
        } else {
            F.error = "unsupported environment map type";
            return false;
        }
        env.is_n = e->N;
        F.envM = e->M;
        F.envMs = e->Ms;
        F.envMcs = e->Mcs;
        env.M = F.envM.data();
        env.Ms = F.envMs.data();
        env.Mcs = F.envMcs.data();
    }
    if (!F.error.empty())
        return false;

    wpt_scene_desc& d = F.desc;
    std::memset(&d, 0, sizeof(d));
    d.abi_version = WPT_ABI_VERSION;
    d.n_animations = F.animations.size();   d.animations = F.animations.data();
    d.n_keyframes = F.keyframes.size();     d.keyframes = F.keyframes.data();
    d.n_textures = F.textures.size();       d.textures = F.textures.data();
    d.texel_bytes = F.texels.size();        d.texels = F.texels.data();
    d.n_materials = F.materials.size();     d.materials = F.materials.data();
    d.n_meshes = F.meshes.size();           d.meshes = F.meshes.data();
    d.n_vertex_floats = F.vertices.size();  d.vertices = F.vertices.data();
    d.n_indices = F.indices.size();         d.indices = F.indices.data();
    d.n_instances = F.instances.size();     d.instances = F.instances.data();
    d.n_spheres = F.spheres.size();         d.spheres = F.spheres.data();
    d.n_media = F.media.size();             d.media = F.media.data();
    d.n_boundary = F.boundary.size();       d.boundary_hitables = F.boundary.data();
    d.n_components = F.components.size();   d.components = F.components.data();
    d.envmap = env;
    return true;
}

/* --------------------------------------------------------- camera / sensor */

void flattenCamera(const Camera& cam, wpt_camera_desc& c, std::vector<wpt_keyframe>& keyframeStorage)
{
    std::memset(&c, 0, sizeof(c));
    c.surround_mode = (cam.surroundMode == Camera::Surround_Off ? WPT_SURROUND_OFF
            : cam.surroundMode == Camera::Surround_180 ? WPT_SURROUND_180 : WPT_SURROUND_360);
    c.stereoscopic_distance = cam.stereoscopicDistance;
    c.l = cam.optics.projection.l; c.r = cam.optics.projection.r;
    c.b = cam.optics.projection.b; c.t = cam.optics.projection.t;
    const LensDistortion& ld = cam.optics.distortion;
    c.distortion_type = int(ld.type);
    c.k1 = ld.k1; c.k2 = ld.k2; c.k3 = ld.k3; c.p1 = ld.p1; c.p2 = ld.p2;
    c.b1 = ld.b1; c.b2 = ld.b2; c.b3 = ld.b3; c.b4 = ld.b4;
    c.lens_radius = cam.optics.depthOfField.lensRadius;
    c.focus_dist = cam.optics.depthOfField.focusDist;
    setv(c.pose_translation, cam.transformation.translation);
    setq(c.pose_rotation, cam.transformation.rotation);
    setv(c.pose_scaling, cam.transformation.scaling);
    keyframeStorage.clear();
    if (cam.animation.get()) {
        if (auto* kf = dynamic_cast<const AnimationKeyframes*>(cam.animation.get()))
            for (auto& k : kf->_keyframes)
                keyframeStorage.push_back(toKeyframe(k.t, k.transformation));
    }
    c.n_keyframes = keyframeStorage.size();
    c.keyframes = keyframeStorage.data();
}

Transformation toTransformation(const float* tr, const float* rot, const float* sc)
{
    return Transformation(vec3(tr[0], tr[1], tr[2]), quat(rot[0], rot[1], rot[2], rot[3]), vec3(sc[0], sc[1], sc[2]));
}

Camera makeRefCamera(const wpt_camera_desc& c)
{
    Projection P(c.l, c.r, c.b, c.t);
    LensDistortion LD;
    LD.type = LensDistortion::Type(c.distortion_type);
    LD.k1 = c.k1; LD.k2 = c.k2; LD.k3 = c.k3; LD.p1 = c.p1; LD.p2 = c.p2;
    LD.b1 = c.b1; LD.b2 = c.b2; LD.b3 = c.b3; LD.b4 = c.b4;
    LensDepthOfField DOF;
    DOF.lensRadius = c.lens_radius;
    DOF.focusDist = c.focus_dist;
    Optics optics(P, LD, DOF);
    Camera::SurroundMode mode = (c.surround_mode == WPT_SURROUND_OFF ? Camera::Surround_Off
            : c.surround_mode == WPT_SURROUND_180 ? Camera::Surround_180 : Camera::Surround_360);
    if (c.n_keyframes > 0) {
        std::vector<AnimationKeyframes::Keyframe> kfs;
        for (uint32_t i = 0; i < c.n_keyframes; i++) {
            const wpt_keyframe& k = c.keyframes[i];
            kfs.push_back(AnimationKeyframes::Keyframe(k.t, toTransformation(k.translation, k.rotation, k.scaling)));
        }
        return Camera(mode, c.stereoscopic_distance, optics, new AnimationKeyframes(kfs));
    }
    return Camera(mode, c.stereoscopic_distance, optics,
            toTransformation(c.pose_translation, c.pose_rotation, c.pose_scaling));
}

Parameters makeRefParams(const wpt_params& p)
{
    Parameters params;
    params.maxPathComponents = p.max_path_components;
    params.rrThreshold = p.rr_threshold;
    params.randomizeRayOverPixel = p.randomize_ray_over_pixel != 0;
    params.minHitDistance = p.min_hit_distance;
    return params;
}

/* ------------------------------------------------------------------ handles */

struct RefScene {
    Scene scene;
    FlatScene flat;
    std::vector<int32_t> idOfHitable;       // unused; ids come from flat.hitableId
};

std::vector<std::unique_ptr<RefScene>> g_scenes;
std::string g_error;

RefScene* getScene(int h)
{
    if (h < 0 || size_t(h) >= g_scenes.size() || !g_scenes[h])
        return nullptr;
    return g_scenes[h].get();
}

}

extern "C" {

const char* ref_last_error(void)
{
    return g_error.c_str();
}

/* variant 0: scenes/configs.hpp; variant 1: the reference example app's own createScene() */
int ref_scene_create(const char* name, float scale, int variant)
{
    std::unique_ptr<RefScene> rs(new RefScene);
    std::string n(name);
    bool ok = true;
    if (variant == 0) {
        ok = wptcfg::buildScene(n, rs->scene, scale);
    } else {
        if (n == "cornell") app_cornell::createScene(rs->scene, 0, 0, 0, false);
        else if (n == "cornell_sphere_glass") app_cornell::createScene(rs->scene, 1, 1, 2, false);
        else if (n == "cornell_moving") app_cornell::createScene(rs->scene, 0, 0, 0, true);
        else if (n == "rtiow") app_rtiow::createRandomScene(rs->scene);
        else if (n == "tof") app_tof::createScene(rs->scene);
        else if (n == "media") app_media::createScene(rs->scene);
        else if (n == "mis") app_mis::createScene(rs->scene, true);
        else if (n == "mis_nohot") app_mis::createScene(rs->scene, false);
        else ok = false;
    }
    if (!ok) {
        g_error = "unknown scene '" + n + "'";
        return -1;
    }
    if (!flattenScene(rs->flat, rs->scene)) {
        g_error = rs->flat.error;
        return -1;
    }
    g_scenes.push_back(std::move(rs));
    return int(g_scenes.size()) - 1;
}

void ref_scene_destroy(int h)
{
    if (getScene(h))
        g_scenes[h].reset();
}

const wpt_scene_desc* ref_scene_desc(int h)
{
    RefScene* rs = getScene(h);
    return rs ? &rs->flat.desc : nullptr;
}

int ref_scene_update_bvh(int h, float t0, float t1)
{
    RefScene* rs = getScene(h);
    if (!rs)
        return -1;
    rs->scene.updateBVH(t0, t1);
    return 0;
}

/* Default view of a named config, as flat descriptions. The camera key frames (if any) live
 * in static storage valid until the next call. */
int ref_default_view(const char* name, wpt_camera_desc* cam, wpt_sensor_desc* sensor, wpt_params* params,
        uint32_t* samplesSqrt, float* t0, float* t1)
{
    static std::vector<wpt_keyframe> keyframeStorage;
    wptcfg::View v = wptcfg::defaultView(name);
    Camera camera = wptcfg::makeCamera(v, float(v.width) / v.height
            / (v.stereoscopicDistance > 0.0f && v.surround == Camera::Surround_Off ? 1.0f : 1.0f));
    flattenCamera(camera, *cam, keyframeStorage);
    std::memset(sensor, 0, sizeof(*sensor));
    sensor->width = v.width;
    sensor->height = v.height;
    if (v.tof) {
        SensorTofAmcw s(4, 4);
        sensor->type = WPT_SENSOR_TOF_AMCW;
        sensor->pixel_area = s.pixelArea;
        sensor->exposure_time = s.exposureTime;
        sensor->contrast = s.contrast;
        sensor->phase_index = 0;
        sensor->phase_count = s.phaseImageCount;
        sensor->modulation_frequency = s.modulationFrequency;
    } else {
        sensor->type = WPT_SENSOR_RGB;
        sensor->min_dist_to_light = 0.0f;
        sensor->max_dist_to_light = std::numeric_limits<float>::max();
        sensor->min_path_len = 0.0f;
        sensor->max_path_len = std::numeric_limits<float>::max();
    }
    params->max_path_components = v.params.maxPathComponents;
    params->rr_threshold = v.params.rrThreshold;
    params->randomize_ray_over_pixel = v.params.randomizeRayOverPixel;
    params->min_hit_distance = v.params.minHitDistance;
    *samplesSqrt = v.samplesSqrt;
    *t0 = v.t0;
    *t1 = v.t1;
    return 0;
}

/* The reference's linearised top-level tree (BVH::_nodes, bvh.hpp:217-246) in the layout of
 * wpt_scene_get_binary_nodes(). Returns the node count, or -1. */
long long ref_bvh_nodes(int h, uint64_t capacity, float* boxes, uint32_t* link, int32_t* leaf)
{
    RefScene* rs = getScene(h);
    if (!rs || !rs->scene._bvh)
        return -1;
    const std::vector<BVHNodeLinear>& nodes = rs->scene._bvh->_nodes;
    if (!boxes)
        return (long long)nodes.size();
    if (capacity < nodes.size())
        return -1;
    for (size_t i = 0; i < nodes.size(); i++) {
        for (int k = 0; k < 3; k++) {
            boxes[i * 6 + k] = nodes[i].aabb.lo[k];
            boxes[i * 6 + 3 + k] = nodes[i].aabb.hi[k];
        }
        if (nodes[i].child2IndexMulTwoPlusOne % 2 == 1) {
            link[i] = nodes[i].child2IndexMulTwoPlusOne / 2;
            leaf[i] = -1;
        } else {
            link[i] = 0;
            auto it = rs->flat.hitableId.find(nodes[i].hitable);
            leaf[i] = (it == rs->flat.hitableId.end() ? -2 : it->second);
        }
    }
    return (long long)nodes.size();
}

/* mcpt() of the reference (wurblpt.hpp:279-436). Returns wall seconds spent inside mcpt(),
 * or a negative value on error. The BVH must have been updated for [t0,t1]. */
double ref_render(int h, const wpt_camera_desc* cam, const wpt_sensor_desc* sensor, const wpt_params* params,
        uint32_t samplesSqrt, float t0, float t1, float* frame)
{
    RefScene* rs = getScene(h);
    if (!rs)
        return -1.0;
    Camera camera = makeRefCamera(*cam);
    Parameters p = makeRefParams(*params);
    FILE* devnull = std::fopen("/dev/null", "w");
    FILE* savedStderr = stderr;
    if (devnull && !std::getenv("WPT_REF_VERBOSE"))
        stderr = devnull;       // mcpt() prints progress text
    double seconds;
    size_t n = size_t(sensor->width) * sensor->height * 3;
    if (sensor->type == WPT_SENSOR_TOF_AMCW) {
        SensorTofAmcw s(sensor->width, sensor->height);
        s.pixelArea = sensor->pixel_area;
        s.exposureTime = sensor->exposure_time;
        s.contrast = sensor->contrast;
        s.phaseImageCount = sensor->phase_count;
        s.modulationFrequency = sensor->modulation_frequency;
        s.setPhaseIndex(sensor->phase_index);
        auto start = std::chrono::steady_clock::now();
        mcpt(s, camera, rs->scene, samplesSqrt, t0, t1, p);
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        std::memcpy(frame, s.energy().data(), n * sizeof(float));
    } else {
        SensorRGB s(sensor->width, sensor->height, sensor->min_dist_to_light, sensor->max_dist_to_light,
                sensor->min_path_len, sensor->max_path_len);
        auto start = std::chrono::steady_clock::now();
        mcpt(s, camera, rs->scene, samplesSqrt, t0, t1, p);
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        std::memcpy(frame, s.result().data(), n * sizeof(float));
    }
    stderr = savedStderr;
    if (devnull)
        std::fclose(devnull);
    return seconds;
}

/* Primary-ray closest hit through pixel centres, randomness off: the loop of
 * getGroundTruth() (wurblpt.hpp:659-668). */
int ref_primary_hits(int h, const wpt_camera_desc* cam, uint32_t width, uint32_t height, float t,
        float minHitDistance, int32_t* ids, float* a)
{
    RefScene* rs = getScene(h);
    if (!rs)
        return -1;
    Camera camera = makeRefCamera(*cam);
    const Camera::RayHelper rayHelper = camera.getRayHelper(t, width, height);
    std::vector<AnimationCache> caches(omp_get_max_threads(), AnimationCache(rs->scene.animations()));
    #pragma omp parallel for schedule(dynamic, 256)
    for (size_t pixel = 0; pixel < size_t(width) * height; pixel++) {
        unsigned int y = pixel / width;
        unsigned int x = pixel % width;
        Prng prng(pixel);
        AnimationCache& cache = caches[omp_get_thread_num()];
        float p = (x + 0.5f) / width;
        float q = (y + 0.5f) / height;
        Ray ray = camera.getRay(p, q, t, t, rayHelper, prng, false);
        cache.init(ray.time);
        HitRecord hr = rs->scene.bvh().hit(ray, RayIntersectionHelper(ray), minHitDistance, maxval,
                minHitDistance, cache, prng);
        if (hr.haveHit) {
            auto it = rs->flat.hitableId.find(hr.hitable);
            ids[pixel] = (it == rs->flat.hitableId.end() ? -2 : it->second);
            a[pixel] = hr.a;
        } else {
            ids[pixel] = -1;
            a[pixel] = 0.0f;
        }
    }
    return 0;
}

/* getGroundTruth() of the reference (wurblpt.hpp:643-759), all images. out->aov[i] / out->materials as in
 * wpt_get_ground_truth(); every pointer must be valid. The BVH must have been updated for (t0, t0). */
int ref_ground_truth(int h, const wpt_camera_desc* cam, uint32_t width, uint32_t height,
        float t0, float tPrev, float tNext, float minHitDistance, const wpt_ground_truth* out)
{
    RefScene* rs = getScene(h);
    if (!rs)
        return -1;
    Camera camera = makeRefCamera(*cam);
    Parameters p;
    p.minHitDistance = minHitDistance;
    SensorRGB sensor(width, height);
    FILE* devnull = std::fopen("/dev/null", "w");
    FILE* savedStderr = stderr;
    if (devnull && !std::getenv("WPT_REF_VERBOSE"))
        stderr = devnull;       // progress text
    unsigned int bits = GroundTruth::All;
    if (cam->surround_mode != WPT_SURROUND_OFF || cam->stereoscopic_distance > 0.0f)
        bits &= ~(GroundTruth::PixelSpaceOffsetToPrev | GroundTruth::PixelSpaceOffsetToNext);   // asserts, camera.hpp:207-208
    GroundTruth gt = getGroundTruth(sensor, camera, rs->scene, t0, tPrev, tNext, bits, p);
    stderr = savedStderr;
    if (devnull)
        std::fclose(devnull);
    const TGD::Array<float>* images[WPT_GT_FLOAT_IMAGES] = {
        &gt.worldSpacePositions, &gt.worldSpaceGeometryNormals, &gt.worldSpaceGeometryTangents,
        &gt.worldSpaceMaterialNormals, &gt.worldSpaceMaterialTangents,
        &gt.cameraSpacePositions, &gt.cameraSpaceGeometryNormals, &gt.cameraSpaceGeometryTangents,
        &gt.cameraSpaceMaterialNormals, &gt.cameraSpaceMaterialTangents,
        &gt.cameraSpaceDepths, &gt.cameraSpaceDistances, &gt.texCoords,
        &gt.worldSpaceOffsetToPrev, &gt.worldSpaceOffsetToNext, &gt.cameraSpaceOffsetToPrev, &gt.cameraSpaceOffsetToNext,
        &gt.pixelSpaceOffsetToPrev, &gt.pixelSpaceOffsetToNext };
    const size_t pixels = size_t(width) * height;
    for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++) {
        if (bits & (1u << i))
            std::memcpy(out->aov[i], images[i]->data(), pixels * WPT_GT_COMPONENTS[i] * sizeof(float));
        else
            std::memset(out->aov[i], 0, pixels * WPT_GT_COMPONENTS[i] * sizeof(float));
    }
    std::memcpy(out->materials, gt.materials.data(), pixels * sizeof(int));
    return 0;
}

/* BVH::hit() (bvh.hpp:277-311) for caller-supplied rays. Directions must be unit length. */
int ref_trace_rays(int h, const float* origins, const float* directions, uint64_t n, float time,
        float minHitDistance, int32_t* ids, float* a)
{
    RefScene* rs = getScene(h);
    if (!rs)
        return -1;
    std::vector<AnimationCache> caches(omp_get_max_threads(), AnimationCache(rs->scene.animations()));
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint64_t i = 0; i < n; i++) {
        Prng prng(i);
        AnimationCache& cache = caches[omp_get_thread_num()];
        Ray ray(vec3(origins + 3 * i), vec3(directions + 3 * i), time, vec4(1.0f));
        cache.init(time);
        HitRecord hr = rs->scene.bvh().hit(ray, RayIntersectionHelper(ray), minHitDistance, maxval,
                minHitDistance, cache, prng);
        if (hr.haveHit) {
            auto it = rs->flat.hitableId.find(hr.hitable);
            ids[i] = (it == rs->flat.hitableId.end() ? -2 : it->second);
            a[i] = hr.a;
        } else {
            ids[i] = -1;
            a[i] = 0.0f;
        }
    }
    return 0;
}

/* Camera rays of the reference (camera.hpp:123-185), randomness off, pixel centres:
 * out[pixel*6 .. +5] = origin xyz, direction xyz. */
int ref_camera_rays(const wpt_camera_desc* cam, uint32_t width, uint32_t height, float t, float* out)
{
    Camera camera = makeRefCamera(*cam);
    const Camera::RayHelper rayHelper = camera.getRayHelper(t, width, height);
    for (size_t pixel = 0; pixel < size_t(width) * height; pixel++) {
        unsigned int y = pixel / width;
        unsigned int x = pixel % width;
        Prng prng(pixel);
        Ray ray = camera.getRay((x + 0.5f) / width, (y + 0.5f) / height, t, t, rayHelper, prng, false);
        for (int k = 0; k < 3; k++) {
            out[pixel * 6 + k] = ray.origin[k];
            out[pixel * 6 + 3 + k] = ray.direction[k];
        }
    }
    return 0;
}

/* out[i*nDraws + k] = k-th in01() of Prng(seeds[i]) (prng.hpp:79-95) */
int ref_prng_kat(const uint32_t* seeds, uint32_t nSeeds, uint32_t nDraws, float* out)
{
    for (uint32_t i = 0; i < nSeeds; i++) {
        Prng prng(seeds[i]);
        for (uint32_t k = 0; k < nDraws; k++)
            out[size_t(i) * nDraws + k] = prng.in01();
    }
    return 0;
}

int ref_threads(void)
{
    return omp_get_max_threads();
}

const char* ref_build_flags(void)
{
#ifdef WPT_REF_FLAGS
    return WPT_REF_FLAGS;
#else
    return "unknown";
#endif
}

/* Traversal counters of the instrumented build (0 elsewhere):
 * out = { path-loop queries, all BVH::hit calls, node visits, leaf tests } */
int ref_counters(uint64_t* out, int reset)
{
    out[0] = out[1] = out[2] = out[3] = 0;
#ifdef WPT_REF_COUNTERS
    for (int i = 0; i < WPT_REF_MAX_THREADS; i++) {
        out[0] += wpt_ref_counters[i].rays;
        out[1] += wpt_ref_counters[i].calls;
        out[2] += wpt_ref_counters[i].nodes;
        out[3] += wpt_ref_counters[i].leaves;
        if (reset)
            wpt_ref_counters[i] = WptRefCounters();
    }
    return 1;
#else
    (void)reset;
    return 0;
#endif
}

}
