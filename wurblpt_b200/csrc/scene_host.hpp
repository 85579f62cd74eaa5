/*
 * scene_host.hpp -- host-side scene state of the path-tracing core.
 *
 * Holds a deep copy of a wpt_scene_desc, and per Scene::updateBVH(t0,t1) call (reference
 * scene.hpp:151-169) the bounds of every hitable, the SAH tree in the reference's linear
 * layout (bvh.hpp:217-246), the 4-wide tree derived from it, and the primitive arrays in the
 * order the device wants them. No CUDA here; device upload lives in device_scene.cu.
 */
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "wurblpt_b200.h"
#include "host_math.hpp"

namespace wpt {

/* Leaf references used on the device: (type << 29) | slot. */
enum : uint32_t { PRIM_TRIANGLE = 0, PRIM_SPHERE = 1, PRIM_MEDIUM = 2, PRIM_TRIANGLE_RUN = 3 };
/* wide-tree leaves may hold a run of up to 4 consecutive triangle slots: (3 << 29) | (count-1) << 27 | first slot */
constexpr uint32_t RUN_COUNT_SHIFT = 27;
constexpr uint32_t RUN_SLOT_MASK = (1u << RUN_COUNT_SHIFT) - 1u;
constexpr uint32_t REF_TYPE_SHIFT = 29;
constexpr uint32_t REF_SLOT_MASK = (1u << REF_TYPE_SHIFT) - 1u;
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;
constexpr uint32_t INNER = 0xFFFFFFFFu;
constexpr int32_t EMPTY_CHILD = 0x7FFFFFFF;

/* One hitable of the expanded component list (global hitable id = index). */
struct Prim {
    uint32_t type;      /* PRIM_* */
    uint32_t ref;       /* instance / sphere / medium index */
    uint32_t tri;       /* triangle number within the mesh */
    uint32_t flags;     /* WPT_COMPF_* of its component */
};

/* Reference layout (BVHNodeLinear, bvh.hpp:217-225): 32 bytes, child 1 is the next node.
 * leaf == INNER: inner node, `link` = index of child 2. Otherwise `leaf` holds the leaf
 * reference (global hitable id while building, device leaf reference afterwards). */
struct BNode {
    float lo[3];
    float hi[3];
    uint32_t link;
    uint32_t leaf;
};
static_assert(sizeof(BNode) == 32, "binary node must be 32 bytes");

/* 4-wide node, 128 bytes = one cache line = eight 16-byte vector loads.
 * child[i] >= 0: index of an inner Node4; child[i] < 0: ~(leaf reference); EMPTY_CHILD: unused
 * (its box is a far-away point no ray reaches). */
struct Node4 {
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int32_t child[4];
    uint32_t pad[4];
};
static_assert(sizeof(Node4) == 128, "wide node must be 128 bytes");

struct HostScene {
    /* ---- deep copy of the description ---- */
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
    wpt_envmap env;
    std::vector<float> envM;
    std::vector<int32_t> envMs;
    std::vector<float> envMcs;

    /* ---- derived once ---- */
    std::vector<Prim> prims;                /* by global hitable id */
    std::vector<uint32_t> topLevel;         /* global ids in BVH input order (Scene::_hitables) */
    std::vector<uint32_t> hotSpotIds;       /* global ids in Scene::_hotSpots order */

    /* ---- per updateBVH(t0, t1) ---- */
    bool haveBvh = false;
    float bvhT0 = 0.0f, bvhT1 = 0.0f;
    std::vector<BNode> bnodes;              /* top-level tree at 0, nested medium trees after it */
    std::vector<uint32_t> mediumRoot;       /* per medium: root index into bnodes, NO_NODE if the boundary is empty */
    std::vector<Node4> nodes4;
    int32_t root4 = EMPTY_CHILD;            /* encoded like a Node4 child */
    uint32_t binaryDepth = 0, wideDepth = 0;
    uint32_t wideLeafTriangles = 4;         /* max triangles per leaf of the wide tree */
    uint64_t topLevelNodes = 0;
    std::vector<uint32_t> refOfGlobal;      /* global id -> device leaf reference */
    std::vector<uint32_t> triGlobal, sphGlobal, medGlobal;  /* slot -> global id */
    std::vector<float> triData;             /* 12 floats per triangle slot */
    std::vector<uint32_t> triShade;         /* 4 words per triangle slot: offsets of the three vertices' attribute
                                               blocks in `vertices` (floats), mesh flags */
    std::vector<float> sphData;             /* 12 floats per sphere slot */
    std::vector<float> medData;             /* 4 floats per medium slot */
    std::vector<uint32_t> hotSpotRefs;      /* device leaf references of the hot spots */
    std::vector<uint8_t> triClass, sphClass, medClass;  /* material class of every slot (device_scene.cuh MC_*) */
    uint32_t classMask = 0;                 /* classes present */
    uint8_t materialClass(int32_t material) const;
    double buildSeconds = 0.0;
    uint64_t nTriSlots = 0, nSphSlots = 0, nWideNodes = 0;  /* sizes of the device arrays (either builder) */
    float sceneLo[3] = { 0, 0, 0 }, sceneHi[3] = { 0, 0, 0 };  /* box of the top level (either builder) */

    /* ---- inputs of the device builder (bvh_device.cu), in hitable order ---- */
    bool deviceBuilt = false;               /* the current tree was built on the device: there is no binary tree */
    std::vector<float> devBox;              /* 8 floats per hitable: lo xyz | type, hi xyz | index among its type */
    std::vector<float> devTriPrim;          /* 12 floats per triangle (hitable order) */
    std::vector<uint32_t> devTriShadePrim;  /* 4 words per triangle */
    std::vector<uint8_t> devTriClassPrim, devSphClassPrim;
    std::vector<float> devSphPrim;          /* 12 floats per sphere */

    std::string error;

    /* Copy and validate a description. Returns WPT_OK or an error code (message in `error`). */
    int ingest(const wpt_scene_desc* desc);

    /* Scene::updateBVH(t0, t1). */
    int updateBvh(float t0, float t1);

    /* The host part of Scene::updateBVH(t0, t1) when the tree is built on the device: bounds of every hitable
     * over [t0,t1] and the primitive data in hitable order. WPT_E_UNSUPPORTED (no message): this scene needs the
     * host builder (media with nested boundary trees, fewer than two hitables). */
    int prepareDeviceBuild(float t0, float t1);
    /* Book-keeping after launchDeviceBuild(). */
    void finishDeviceBuild(float t0, float t1, uint32_t wideNodes, uint32_t depth, uint32_t triangles, double seconds);

    /* AnimationKeyframes::at(t) (animation_keyframes.hpp:191-215) for scene animation `index`. */
    Trs animationAt(int index, float t) const;
};

/* Key-frame interpolation shared with the camera animation. */
Trs keyframesAt(const wpt_keyframe* keyframes, uint32_t count, float t);

}
