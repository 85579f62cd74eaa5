/* scene.hpp -- the Scene container (mirror of reference libwurblpt/scene.hpp:48-241: same take()
 * overloads, return types and ownership rules).
 *
 * Differences that follow from moving the hot path to the GPU:
 *  - updateBVH(t0,t1) flattens the scene into the C ABI's plain arrays, hands them to
 *    wpt_scene_create() (upload) and wpt_scene_update_bvh() (bounds, SAH tree, device layouts);
 *  - the scene keeps the resulting device handle; mcpt() renders from it.
 * Failures are reported like the reference reports its own: a message on stderr, then abort().
 */
#pragma once

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "animation.hpp"
#include "animation_keyframes.hpp"
#include "texture.hpp"
#include "material.hpp"
#include "envmap.hpp"
#include "hitable.hpp"
#include "mesh.hpp"
#include "sphere.hpp"
#include "medium.hpp"
#include "scene_component.hpp"
#include "flatten.hpp"

namespace WurblPT {

typedef enum
{
    ColdSpot,
    HotSpot
} HotSpotType;

namespace detail {

[[noreturn]] inline void die(const std::string& what, const char* message)
{
    fprintf(stderr, "%s: %s\n", what.c_str(), message);
    abort();
}

/* Which GPUs render: by default ALL of the box (or those listed in WURBLPT_DEVICES, e.g. "0,1,2,3") work on every
 * frame together (wurblpt_b200.h, WPT_DEVICE_ALL; MPICoordinator::worldSize() reports how many). WURBLPT_DEVICE
 * selects a single CUDA device instead; "none" prepares scenes on the host only (flattening and tree build can
 * then be inspected, rendering is refused). */
inline int deviceFromEnvironment()
{
    const char* e = std::getenv("WURBLPT_DEVICE");
    if (!e || !*e)
        return WPT_DEVICE_ALL;
    if (std::string(e) == "none")
        return WPT_DEVICE_NONE;
    return std::atoi(e);
}

}

/* The scene takes ownership of everything handed to take(). */
class Scene
{
private:
    struct Taken {
        const SceneComponent* component;
        size_t first, count;        // range in _hitables or _unhitables
        bool hotSpot;
    };

    std::vector<std::unique_ptr<Animation>> _animations;
    std::vector<const Animation*> _animationsForCaching;
    std::vector<std::unique_ptr<Mesh>> _meshes;
    std::vector<std::unique_ptr<SceneComponent>> _components;
    std::vector<std::unique_ptr<Hitable>> _hitables;
    std::vector<const Hitable*> _hotSpots;
    std::vector<std::unique_ptr<Hitable>> _unhitables;
    std::vector<Taken> _takenHitable, _takenUnhitable;
    std::unique_ptr<EnvironmentMap> _envmap;
    bool _bvhNeedsRebuild;
    float _bvhT0, _bvhT1;
    std::map<const Material*, int> _materialMap;
    std::vector<const Material*> _materialsInOrder;
    std::vector<std::string> _materialNames;

    /* device side */
    std::unique_ptr<Flattener> _flat;
    wpt_scene* _device;
    bool _descriptionStale;
    bool _hasBakedAnimations;           /* some Animation is not an AnimationKeyframes (sampled into key frames) or some
                                           Texture is user-defined (sampled into an image): the description depends on time */
    std::vector<float> _bakeTimes;      /* the times the current description samples those animations at */

    void flattenRange(Flattener& F, const std::vector<Taken>& taken, const std::vector<std::unique_ptr<Hitable>>& list,
            bool hitable, int& nextId, std::vector<std::pair<const Medium*, int>>& mediaToResolve)
    {
        for (const Taken& t : taken) {
            wpt_component comp;
            std::memset(&comp, 0, sizeof(comp));
            comp.flags = (hitable ? WPT_COMPF_HITABLE : 0) | (t.hotSpot ? WPT_COMPF_HOTSPOT : 0);
            if (const MeshInstance* mi = dynamic_cast<const MeshInstance*>(t.component)) {
                wpt_instance inst;
                std::memset(&inst, 0, sizeof(inst));
                inst.mesh = F.mesh(mi->mesh);
                inst.material = F.material(mi->material);
                inst.animation = mi->animationIndex;
                inst.has_transform = !mi->transformation.isIdentity();
                mat4 M(1.0f);
                mat3 N(1.0f);
                if (inst.has_transform) {
                    M = mi->transformationM;
                    N = mi->transformationN;
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
            } else if (const Sphere* sp = dynamic_cast<const Sphere*>(t.component)) {
                wpt_sphere s;
                sp->flatten(F, s);
                comp.type = WPT_COMP_SPHERE;
                comp.ref = F.spheres.size();
                F.spheres.push_back(s);
            } else if (const Medium* md = dynamic_cast<const Medium*>(t.component)) {
                wpt_medium m;
                std::memset(&m, 0, sizeof(m));
                m.density = md->density();
                m.material = F.material(md->phaseFunction());
                comp.type = WPT_COMP_MEDIUM;
                comp.ref = F.media.size();
                mediaToResolve.push_back(std::make_pair(md, int(F.media.size())));
                F.media.push_back(m);
            } else {
                F.fail("a user-defined SceneComponent cannot be intersected on the device");
                return;
            }
            for (size_t k = 0; k < t.count; k++)
                list[t.first + k]->_globalId = nextId++;
            F.components.push_back(comp);
        }
    }

    /* Build the C-ABI description. Global hitable ids number the hitables of take() first, then
     * those of takeUnhitable() (the same convention as the oracle's flattening of a reference Scene). */
    void flatten(const std::vector<float>& bakeTimes = std::vector<float>(1, 0.0f))
    {
        _flat.reset(new Flattener);
        Flattener& F = *_flat;
        _hasBakedAnimations = false;
        _bakeTimes = bakeTimes;
        F.bakeTime = bakeTimes.empty() ? 0.0f : bakeTimes.front();
        for (const auto& a : _animations) {
            const AnimationKeyframes* kf = dynamic_cast<const AnimationKeyframes*>(a.get());
            std::vector<AnimationKeyframes::Keyframe> baked;
            if (!kf) {
                /* a user-defined Animation: the device only knows key frames, so at() is sampled at the times
                 * this description is made for (exact when t0 == t1; see detail::bakeTimesFor) */
                _hasBakedAnimations = true;
                for (float t : bakeTimes)
                    baked.push_back(AnimationKeyframes::Keyframe(t, a->at(t)));
            }
            const std::vector<AnimationKeyframes::Keyframe>& keys = (kf ? kf->keyframes() : baked);
            wpt_animation wa;
            wa.first_keyframe = F.keyframes.size();
            wa.keyframe_count = keys.size();
            for (const auto& k : keys) {
                wpt_keyframe wk;
                wk.t = k.t;
                detail::store(wk.translation, k.transformation.translation, 3);
                wk.rotation[0] = k.transformation.rotation.x; wk.rotation[1] = k.transformation.rotation.y;
                wk.rotation[2] = k.transformation.rotation.z; wk.rotation[3] = k.transformation.rotation.w;
                detail::store(wk.scaling, k.transformation.scaling, 3);
                F.keyframes.push_back(wk);
            }
            F.animations.push_back(wa);
        }
        for (const Material* m : _materialsInOrder)
            F.material(m);
        int nextId = 0;
        std::vector<std::pair<const Medium*, int>> mediaToResolve;
        flattenRange(F, _takenHitable, _hitables, true, nextId, mediaToResolve);
        flattenRange(F, _takenUnhitable, _unhitables, false, nextId, mediaToResolve);
        for (const auto& mr : mediaToResolve) {
            wpt_medium& m = F.media[mr.second];
            m.first_boundary = F.boundary.size();
            m.boundary_count = mr.first->boundary().size();
            for (const Hitable* b : mr.first->boundary()) {
                if (!b || b->globalId() < 0) {
                    F.fail("a medium boundary hitable is not owned by this scene");
                    break;
                }
                F.boundary.push_back(b->globalId());
            }
        }
        if (_envmap && !_envmap->flatten(F))
            F.fail("a user-defined EnvironmentMap cannot be evaluated on the device");
        if (F.bakedTextures)
            _hasBakedAnimations = true;     /* the description depends on the time it was made for */
        if (!F.error.empty())
            detail::die("WurblPT scene", F.error.c_str());
    }

    void releaseDevice()
    {
        if (_device)
            wpt_scene_destroy(_device);
        _device = nullptr;
    }

    void rememberTaken(std::vector<Taken>& taken, const SceneComponent* component, size_t first, size_t count, bool hot)
    {
        taken.push_back(Taken{ component, first, count, hot });
        _descriptionStale = true;
    }

public:
    Scene() : _bvhNeedsRebuild(true), _bvhT0(0.0f), _bvhT1(0.0f), _device(nullptr), _descriptionStale(true),
        _hasBakedAnimations(false) {}
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    ~Scene() { releaseDevice(); }

    int take(Animation* anim)
    {
        _animations.push_back(std::unique_ptr<Animation>(anim));
        _animationsForCaching.push_back(anim);
        _descriptionStale = true;
        return _animations.size() - 1;
    }

    Mesh* take(Mesh* mesh)
    {
        _meshes.push_back(std::unique_ptr<Mesh>(mesh));
        return mesh;
    }

    Texture* take(Texture* tex)
    {
        _components.push_back(std::unique_ptr<SceneComponent>(tex));
        return tex;
    }

    Material* take(Material* mat, const std::string& name = std::string())
    {
        _components.push_back(std::unique_ptr<SceneComponent>(mat));
        _materialMap.insert(std::pair<const Material*, int>(mat, _materialNames.size()));
        _materialsInOrder.push_back(mat);
        _materialNames.push_back(name);
        _descriptionStale = true;
        return mat;
    }

    std::vector<const Hitable*> take(SceneComponent* component, HotSpotType hotSpotType = ColdSpot)
    {
        std::vector<Hitable*> hitables = component->createHitables();
        if (hotSpotType == HotSpot)
            _hotSpots.insert(_hotSpots.end(), hitables.begin(), hitables.end());
        rememberTaken(_takenHitable, component, _hitables.size(), hitables.size(), hotSpotType == HotSpot);
        _hitables.reserve(_hitables.size() + hitables.size());
        for (Hitable* h : hitables)
            _hitables.push_back(std::unique_ptr<Hitable>(h));
        _components.push_back(std::unique_ptr<SceneComponent>(component));
        _bvhNeedsRebuild = true;
        return std::vector<const Hitable*>(hitables.begin(), hitables.end());
    }

    std::vector<const Hitable*> takeUnhitable(SceneComponent* component)
    {
        std::vector<Hitable*> hitables = component->createHitables();
        rememberTaken(_takenUnhitable, component, _unhitables.size(), hitables.size(), false);
        _unhitables.reserve(_unhitables.size() + hitables.size());
        for (Hitable* h : hitables)
            _unhitables.push_back(std::unique_ptr<Hitable>(h));
        _components.push_back(std::unique_ptr<SceneComponent>(component));
        _bvhNeedsRebuild = true;
        return std::vector<const Hitable*>(hitables.begin(), hitables.end());
    }

    EnvironmentMap* take(EnvironmentMap* envmap)
    {
        _envmap = std::unique_ptr<EnvironmentMap>(envmap);
        _descriptionStale = true;
        return envmap;
    }

    /* (Re)builds the description and uploads it when needed: first use, objects taken since, or sampled
     * ("baked") animations that were sampled for other times. */
    void prepareDevice(const std::vector<float>& bakeTimes)
    {
        if (_device && !_descriptionStale && !(_hasBakedAnimations && _bakeTimes != bakeTimes))
            return;
        /* Note: a material's public members (e.g. normalTex) may have been changed after take();
         * the description is rebuilt whenever the set of taken objects changed. */
        releaseDevice();
        flatten(bakeTimes);
        wpt_scene_desc d = _flat->desc();
        if (wpt_scene_create(&d, detail::deviceFromEnvironment(), &_device) != WPT_OK)
            detail::die("WurblPT scene upload", wpt_last_error());
        _descriptionStale = false;
    }

    bool bvhNeedsUpdate(float t0 = 0.0f, float t1 = 0.0f) const
    {
        return _bvhNeedsRebuild || (_animations.size() > 0 && (_bvhT0 > t0 || _bvhT1 < t1));
    }

    void updateBVH(float t0 = 0.0f, float t1 = 0.0f)
    {
        if (!bvhNeedsUpdate(t0, t1)) {
            fprintf(stderr, "Bounding volume hierarchy does not need updating.\n");
            return;
        }
        prepareDevice(detail::bakeTimesFor(t0, t1));
        fprintf(stderr, "Building bounding volume hierarchy for %zu hitables for %.3fs-%.3fs\n", _hitables.size(), t0, t1);
        if (wpt_scene_update_bvh(_device, t0, t1) != WPT_OK)
            detail::die("WurblPT updateBVH", wpt_last_error());
        wpt_scene_info info;
        if (wpt_scene_get_info(_device, &info) == WPT_OK)
            fprintf(stderr, "Linearizing bounding volume hierarchy with %zu nodes on %zu levels (device: %zu 4-wide nodes on %zu levels)\n",
                    size_t(info.binary_nodes), size_t(info.binary_depth), size_t(info.wide_nodes), size_t(info.wide_depth));
        _bvhNeedsRebuild = false;
        _bvhT0 = t0;
        _bvhT1 = t1;
    }

    const std::vector<const Animation*>& animations() const { return _animationsForCaching; }
    const std::vector<const Hitable*>& hotSpots() const { return _hotSpots; }
    const EnvironmentMap* environmentMap() const { return _envmap.get(); }

    int materialIndex(const Material* mat) const
    {
        auto it = _materialMap.find(mat);
        return it != _materialMap.end() ? it->second : -1;
    }
    const std::vector<std::string>& materialNames() const { return _materialNames; }

    /* ---- B200 build only ---- */

    /* device handle (valid after updateBVH) */
    wpt_scene* deviceScene() const { return _device; }

    /* Scene::materialIndex() of every material of the flattened description (-1: a material that is only
     * reachable through another one and was never given to take()); for getGroundTruth() */
    std::vector<int32_t> materialIdsOfDescription() const
    {
        std::vector<int32_t> ids(_flat ? _flat->materials.size() : 0, -1);
        if (_flat) {
            for (const auto& kv : _flat->materialIndex) {
                auto it = _materialMap.find(kv.first);
                if (it != _materialMap.end())
                    ids[kv.second] = it->second;
            }
        }
        return ids;
    }

    /* the flattened description of the last updateBVH() (for inspection and tests) */
    const Flattener* flattened() const { return _flat.get(); }

    /* flatten without touching a device; (t0, t1) only matters for animations that are sampled into key frames */
    wpt_scene_desc describe(float t0 = 0.0f, float t1 = 0.0f)
    {
        flatten(detail::bakeTimesFor(t0, t1));
        return _flat->desc();
    }

    bool hasBakedAnimations() const { return _hasBakedAnimations; }

    /* mcpt(t0, t1) on a description that was sampled for other times (user-defined textures are functions of
     * time, and nothing obliges the caller to call updateBVH() again when the scene has no animations):
     * sample again for (t0, t1) and rebuild the tree. Logically const. */
    void prepareRender(float t0, float t1) const
    {
        if (!_hasBakedAnimations || _bakeTimes == detail::bakeTimesFor(t0, t1))
            return;
        Scene* self = const_cast<Scene*>(this);
        self->prepareDevice(detail::bakeTimesFor(t0, t1));
        if (wpt_scene_update_bvh(self->_device, t0, t1) != WPT_OK)
            detail::die("WurblPT updateBVH", wpt_last_error());
    }

    /* getGroundTruth() evaluates animations at tPrev and tNext as well: sampled animations get key frames at
     * exactly these times and the tree for t0 is rebuilt. Logically const (the scene is unchanged). */
    void prepareGroundTruth(float t0, float tPrev, float tNext) const
    {
        if (!_hasBakedAnimations || (tPrev == t0 && tNext == t0))
            return;
        std::vector<float> times;
        times.push_back(tPrev);
        if (t0 > tPrev)
            times.push_back(t0);
        if (tNext > t0)
            times.push_back(tNext);
        Scene* self = const_cast<Scene*>(this);
        self->prepareDevice(times);
        if (wpt_scene_update_bvh(self->_device, t0, t0) != WPT_OK)
            detail::die("WurblPT updateBVH", wpt_last_error());
    }
};

}
