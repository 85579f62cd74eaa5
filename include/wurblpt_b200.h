/*
 * wurblpt_b200.h -- C ABI of the B200-native WurblPT path-tracing core.
 *
 * This is the drop-in boundary for the one hot path this repository replaces:
 *   mcpt() -> tracePath() -> BVH::hit()      (reference libwurblpt/wurblpt.hpp:108-436,
 *                                              libwurblpt/bvh.hpp:277-311)
 * together with Scene::updateBVH() (reference libwurblpt/scene.hpp:151-169) which
 * prepares the acceleration structure that path reads.
 *
 * The reference is a header-only C++ library with a *source* API (no binary plugin
 * ABI). The inline C++ API mirror in include/wurblpt/ flattens its Scene / Camera /
 * Sensor / Parameters objects into the plain structs below and calls these entry
 * points; nothing else crosses the boundary. Only plain pointers and sizes appear
 * here -- no C++ types, no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative WPT_E_* code on failure;
 *     wpt_last_error() returns a thread-local human-readable message. The C++
 *     wrapper turns failures into fprintf(stderr)+abort(), matching the reference's
 *     only error convention (bvh.hpp:260-263, mpi.hpp:191-194).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with WPT_E_NO_DEVICE.
 *   - vec4 colour channels are (r, g, b, nir) as in the reference
 *     (material_lambertian.hpp:53-59); quaternions are stored (x, y, z, w)
 *     (gvm.hpp quaternion); matrices are column-major (gvm.hpp matrix).
 *   - frames are w*h*3 float32, pixel index = y*w + x, row 0 = bottom of the image
 *     (wurblpt.hpp:339-341, camera.hpp:148).
 */
#ifndef WURBLPT_B200_H
#define WURBLPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WPT_ABI_VERSION 1u

/* error codes */
#define WPT_OK              0
#define WPT_E_INVALID      -1   /* invalid argument / malformed scene description */
#define WPT_E_NO_DEVICE    -2   /* no usable CUDA device (there is no CPU fallback) */
#define WPT_E_CUDA         -3   /* a CUDA runtime call or kernel failed */
#define WPT_E_UNSUPPORTED  -4   /* scene uses a feature outside the device's closed set */
#define WPT_E_STATE        -5   /* call order violated, e.g. render before update_bvh */

/* ---------------------------------------------------------------- animations */

/* One key frame: time + Transformation (reference animation_keyframes.hpp:70-78,
 * transformation.hpp:47-62). */
typedef struct wpt_keyframe {
    float t;
    float translation[3];
    float rotation[4];          /* x, y, z, w */
    float scaling[3];
} wpt_keyframe;

/* AnimationKeyframes (animation_keyframes.hpp:191-215): a sorted run of key frames. */
typedef struct wpt_animation {
    uint32_t first_keyframe;
    uint32_t keyframe_count;
} wpt_animation;

/* ------------------------------------------------------------------ textures */

enum {
    WPT_TEX_CONSTANT    = 0,    /* TextureConstant    texture.hpp:160-180 */
    WPT_TEX_CHECKER     = 1,    /* TextureChecker     texture.hpp:182-205 */
    WPT_TEX_IMAGE       = 2,    /* TextureImage       texture_image.hpp:37-233 */
    WPT_TEX_TRANSFORMER = 3     /* TextureTransformer texture.hpp:207-246 */
};

enum { WPT_TEXEL_U8 = 0, WPT_TEXEL_U16 = 1, WPT_TEXEL_F32 = 2 };

typedef struct wpt_texture {
    int32_t type;
    int32_t src_texture;        /* TRANSFORMER: index of the wrapped texture */
    int32_t width, height;      /* IMAGE */
    int32_t components;         /* IMAGE: 1..4 as stored */
    int32_t texel_type;         /* IMAGE: WPT_TEXEL_* */
    int32_t linearize_srgb;     /* IMAGE: decode sRGB on rgb channels (texture_image.hpp:235-262) */
    int32_t horiz, vert;        /* CHECKER */
    int32_t _pad;
    uint64_t texel_offset;      /* IMAGE: byte offset into wpt_scene_desc.texels, x fastest, row 0 first */
    float color0[4];            /* CONSTANT colour / CHECKER colour 0 */
    float color1[4];            /* CHECKER colour 1 */
    float coord_factor[2];      /* IMAGE, TRANSFORMER */
    float coord_offset[2];
    float val_factor[4];
    float val_offset[4];
} wpt_texture;

/* ----------------------------------------------------------------- materials */

enum {
    WPT_MAT_NONE            = 0,    /* base Material: no scatter, no emission (material.hpp:158-191) */
    WPT_MAT_LAMBERTIAN      = 1,    /* material_lambertian.hpp:61-102 */
    WPT_MAT_MIRROR          = 2,    /* material_mirror.hpp:53-62 */
    WPT_MAT_GLASS           = 3,    /* material_glass.hpp:91-152 */
    WPT_MAT_GGX             = 4,    /* material_ggx.hpp:173-257 */
    WPT_MAT_MODPHONG        = 5,    /* material_modphong.hpp:183-327 */
    WPT_MAT_PHASE_ISOTROPIC = 6,    /* material_phase_function_isotropic.hpp:51-59 */
    WPT_MAT_LIGHT_DIFFUSE   = 7,    /* light_diffuse.hpp:50-61 */
    WPT_MAT_LIGHT_SPOT      = 8,    /* light_spot.hpp:45-60 */
    WPT_MAT_LIGHT_TOF       = 9,    /* light_tof.hpp:55-74 */
    WPT_MAT_TWOSIDED        = 10    /* material.hpp:273-334 */
};

#define WPT_MATF_HAVE_NIR            1u  /* 4th channel given explicitly (vec4 ctor), else mean(rgb) */
#define WPT_MATF_CHROMATIC           2u  /* GLASS: chromaticDispersion (material_glass.hpp:66-69) */
#define WPT_MATF_DIFFUSE_TEX_ALPHA   4u  /* MODPHONG: diffuseTexHasAlpha */
#define WPT_MATF_SPECULAR_TEX_ALPHA  8u  /* MODPHONG: specularTexHasAlpha */

/*
 * Parameter slots per type (unused slots are ignored):
 *   LAMBERTIAN      v0 albedo                               tex0 albedo
 *   MIRROR          v0 color                                tex0 color
 *   GLASS           v0 absorption, v1 ior material, v2 ior surrounding
 *   GGX             v0 albedo, f0 f1 roughness x y          tex0 albedo, tex1 roughness
 *   MODPHONG        v0 diffuse, v1 specular, v2 transmissive, v3 emissive,
 *                   f0 shininess, f1 opacity, f2 index of refraction
 *                   tex0 diffuse, tex1 specular, tex2 shininess, tex3 opacity, tex4 emissive
 *   PHASE_ISOTROPIC v0 albedo
 *   LIGHT_DIFFUSE   v0 emission                             tex0
 *   LIGHT_SPOT      v0 emission, f0 cos(half opening angle) tex0
 *   LIGHT_TOF       f0 radiance, f1 cos(half opening angle) tex0
 *   TWOSIDED        front, back (material indices, -1 = null pointer)
 */
typedef struct wpt_material {
    int32_t type;
    int32_t normal_tex;         /* Material::normalTex, -1 = none (material.hpp:132) */
    int32_t tex[5];             /* texture indices, -1 = none */
    uint32_t flags;
    int32_t front, back;
    int32_t _pad[2];
    float v0[4], v1[4], v2[4], v3[4];
    float f[4];
} wpt_material;

/* ------------------------------------------------------------------ geometry */

#define WPT_MESHF_TEXCOORDS 1u
#define WPT_MESHF_TANGENTS  2u

/* Mesh (mesh.hpp:38-153): interleaved vertices pos(3) nrm(3) [tc(2)] [tng(3)]. */
typedef struct wpt_mesh {
    uint64_t vertex_offset;     /* in floats, into wpt_scene_desc.vertices */
    uint64_t index_offset;      /* in uint32, into wpt_scene_desc.indices */
    uint32_t vertex_count;
    uint32_t vertex_stride;     /* 6, 8 or 11 floats */
    uint32_t triangle_count;
    uint32_t flags;             /* WPT_MESHF_* */
} wpt_mesh;

/* MeshInstance (mesh.hpp:159-189). */
typedef struct wpt_instance {
    int32_t mesh;
    int32_t material;           /* -1 = nullptr (e.g. medium boundaries) */
    int32_t animation;          /* -1 = none */
    int32_t has_transform;      /* !transformation.isIdentity() */
    float M[16];                /* transformationM, column-major */
    float N[9];                 /* transformationN, column-major */
    float _pad[3];
} wpt_instance;

/* HitableSphere (hitable_sphere.hpp:33-86). */
typedef struct wpt_sphere {
    float center[3];
    float radius;
    float rotation[4];
    int32_t material;
    int32_t animation;
    int32_t _pad[2];
} wpt_sphere;

/* HitableMedium (hitable_medium.hpp:36-98). The boundary is a run of global hitable
 * ids (see wpt_component) in wpt_scene_desc.boundary_hitables. */
typedef struct wpt_medium {
    float density;
    int32_t material;           /* phase function */
    uint32_t first_boundary;
    uint32_t boundary_count;
} wpt_medium;

enum { WPT_COMP_MESH_INSTANCE = 0, WPT_COMP_SPHERE = 1, WPT_COMP_MEDIUM = 2 };

#define WPT_COMPF_HITABLE 1u    /* Scene::take (in the BVH); unset = Scene::takeUnhitable (scene.hpp:127-135) */
#define WPT_COMPF_HOTSPOT 2u    /* HotSpotType::HotSpot (scene.hpp:48-52,113-125) */

/*
 * One Scene::take(SceneComponent*) call. Components expand, in order, into the
 * scene's hitables: a mesh instance into one hitable per triangle, a sphere or a
 * medium into one. The position of a hitable in that expansion (over ALL components,
 * hitable or not) is its *global hitable id*: the id wpt_primary_hits() reports and
 * medium boundaries refer to.
 */
typedef struct wpt_component {
    int32_t type;               /* WPT_COMP_* */
    int32_t ref;                /* index into instances / spheres / media */
    uint32_t flags;             /* WPT_COMPF_* */
    uint32_t _pad;
} wpt_component;

/* ------------------------------------------------------------ environment map */

enum { WPT_ENV_NONE = 0, WPT_ENV_EQUIRECT = 1, WPT_ENV_CUBE = 2 };
enum { WPT_ENV_COMPAT_MITSUBA = 0, WPT_ENV_COMPAT_SURROUND_VIDEO = 1 };

/* EnvironmentMap (envmap.hpp:44-285). The importance tables are those built by
 * initializeImportanceSampling(N) (envmap.hpp:121-158); is_n = 0 means it was not called. */
typedef struct wpt_envmap {
    int32_t type;
    int32_t compat;             /* EQUIRECT */
    int32_t tex[6];             /* EQUIRECT: tex[0]; CUBE: posx negx posy negy posz negz */
    int32_t is_n;
    int32_t _pad;
    const float* M;             /* is_n*is_n */
    const int32_t* Ms;
    const float* Mcs;
} wpt_envmap;

/* --------------------------------------------------------------------- scene */

typedef struct wpt_scene_desc {
    uint32_t abi_version;       /* WPT_ABI_VERSION */
    uint32_t n_animations;
    const wpt_animation* animations;
    uint32_t n_keyframes;
    uint32_t n_textures;
    const wpt_keyframe* keyframes;
    const wpt_texture* textures;
    uint64_t texel_bytes;
    const uint8_t* texels;
    uint32_t n_materials;
    uint32_t n_meshes;
    const wpt_material* materials;
    const wpt_mesh* meshes;
    uint64_t n_vertex_floats;
    const float* vertices;
    uint64_t n_indices;
    const uint32_t* indices;
    uint32_t n_instances;
    uint32_t n_spheres;
    const wpt_instance* instances;
    const wpt_sphere* spheres;
    uint32_t n_media;
    uint32_t n_boundary;
    const wpt_medium* media;
    const uint32_t* boundary_hitables;
    uint32_t n_components;
    uint32_t _pad;
    const wpt_component* components;
    wpt_envmap envmap;
} wpt_scene_desc;

/* -------------------------------------------------------------------- camera */

enum { WPT_SURROUND_OFF = 0, WPT_SURROUND_180 = 1, WPT_SURROUND_360 = 2 };
enum { WPT_DISTORTION_NONE = 0, WPT_DISTORTION_RADIAL_AND_PLANAR = 1,
       WPT_DISTORTION_RADIAL_ONLY = 2, WPT_DISTORTION_OPENCV = 3 };

/* Camera + Optics (camera.hpp:42-107, optics.hpp:37-351). A camera animation that
 * is an AnimationKeyframes is passed as key frames; n_keyframes = 0 uses `pose`. */
typedef struct wpt_camera_desc {
    int32_t surround_mode;
    float stereoscopic_distance;
    float l, r, b, t;           /* Projection: frustum at near = 1 (optics.hpp:40) */
    int32_t distortion_type;
    float k1, k2, k3, p1, p2;
    float b1, b2, b3, b4;
    float lens_radius, focus_dist;
    float pose_translation[3];
    float pose_rotation[4];
    float pose_scaling[3];
    uint32_t n_keyframes;
    const wpt_keyframe* keyframes;
} wpt_camera_desc;

/* -------------------------------------------------------------------- sensor */

enum { WPT_SENSOR_RGB = 0, WPT_SENSOR_TOF_AMCW = 1 };

/* SensorRGB (sensor_rgb.hpp:41-87) / SensorTofAmcw (sensor_tof_amcw.hpp:227-259). */
typedef struct wpt_sensor_desc {
    int32_t type;
    uint32_t width, height;
    /* RGB gating */
    float min_dist_to_light, max_dist_to_light;
    float min_path_len, max_path_len;
    /* ToF AMCW */
    float pixel_area;           /* um^2 */
    float exposure_time;        /* us */
    float contrast;
    uint32_t phase_index;       /* setPhaseIndex(), enters tau(), sensor_tof_amcw.hpp:97-105 */
    uint32_t phase_count;       /* phaseImageCount */
    uint32_t _pad;
    double modulation_frequency;/* Hz; f/c is formed in double then cast (sensor_tof_amcw.hpp:107-110) */
} wpt_sensor_desc;

/* Parameters (wurblpt.hpp:79-96). */
typedef struct wpt_params {
    uint32_t max_path_components;
    float rr_threshold;
    int32_t randomize_ray_over_pixel;
    float min_hit_distance;
} wpt_params;

/* ------------------------------------------------------------ render control */

enum {
    WPT_TRAVERSAL_WIDE      = 0,    /* 4-wide BVH, near-to-far ordered (default) */
    WPT_TRAVERSAL_REFERENCE = 1,    /* binary BVH walked in the reference's storage order (bvh.hpp:277-311) */
    WPT_TRAVERSAL_POOL      = 2,    /* queries only: the 4-wide walk as the wavefront pipeline's pooled trace kernel runs it
                                       (scenes without media and animated hitables); renders treat it like WIDE */
    WPT_TRAVERSAL_PERSISTENT = 3,   /* queries only: likewise through the pipeline's persistent per-lane trace kernel, so that
                                       parity tests reach it bit for bit */
    WPT_TRAVERSAL_QUAD      = 4     /* queries only: likewise through the pipeline's quad trace kernel (four lanes per query:
                                       one child of a wide node / one triangle of a leaf each) */
};

enum {
    WPT_PIPELINE_AUTO       = 0,    /* the faster one for the scene (currently: wavefront) */
    WPT_PIPELINE_MEGAKERNEL = 1,    /* one persistent kernel, path state in registers */
    WPT_PIPELINE_WAVEFRONT  = 2     /* shade / trace kernel pairs, path state in HBM */
};

typedef struct wpt_render_opts {
    int32_t traversal;          /* WPT_TRAVERSAL_* */
    int32_t substreams;         /* PRNG sub-streams per pixel: 1 = the reference's one stream per pixel
                                   (wurblpt.hpp:342), 2..128 (power of two) split a pixel's samples over
                                   lanes, 0 = choose from image size */
    uint32_t tile_first;        /* this call renders pixel tiles tile_first, tile_first+tile_stride, ... */
    uint32_t tile_stride;       /* 0 or 1 = all tiles; N = one of N GPUs (the MPICoordinator block split, mpi.hpp:241-254) */
    int32_t frame_on_device;    /* `frame` is a device pointer; untouched tiles are left as they are */
    int32_t pipeline;           /* WPT_PIPELINE_* */
    void* stream;               /* cudaStream_t to launch on (NULL = default stream) */
    void* shared_counter;       /* NULL, or a device pointer to a zeroed 64-bit counter that the GPUs of SEVERAL
                                   calls rendering the same frame draw their blocks of work items from (memory of this
                                   device, of a peer, or opened with wpt_ipc_counter_open in another process): the
                                   MPICoordinator hand-out (mpi.hpp:74-107) as one atomic. tile_stride must be 0 or 1;
                                   every call renders whatever blocks it draws, into a frame that is zero elsewhere,
                                   and the frames are added up (one NCCL reduce in bench.py) */
} wpt_render_opts;

typedef struct wpt_stats {
    uint64_t samples;           /* camera paths started */
    uint64_t rays;              /* closest-hit queries issued from the path loop (wurblpt.hpp:131,204,236) */
    uint64_t nested_queries;    /* medium-boundary queries (hitable_medium.hpp:75-79) */
    uint64_t tiles;             /* pixel tiles rendered by this call */
    float kernel_ms;            /* device time of the render kernel (CUDA events) */
    float total_ms;             /* host wall time of the call including copies */
    uint32_t launches;          /* kernels launched by this call */
    uint32_t _pad;
} wpt_stats;

typedef struct wpt_scene_info {
    uint64_t hitables;          /* in the top-level BVH */
    uint64_t triangles, spheres, media;
    uint64_t binary_nodes;      /* reference-layout nodes (2n-1) */
    uint64_t wide_nodes;
    uint32_t binary_depth, wide_depth;
    uint32_t hot_spots;
    uint32_t warp_tested_media; /* media whose boundary (a mesh of up to 32 triangles) the trace kernel tests with a whole
                                   warp instead of walking the nested boundary tree per lane (hitable_medium.hpp:69-98) */
    uint64_t device_bytes;
    double build_seconds;       /* host SAH build + collapse */
    double upload_seconds;
} wpt_scene_info;

typedef struct wpt_scene wpt_scene;

/* ------------------------------------------------------------- entry points */

/* Library / ABI version string. */
const char* wpt_version(void);

/* Message for the last failure on the calling thread. */
const char* wpt_last_error(void);

/* Number of usable CUDA devices (0 without a GPU). */
int wpt_device_count(void);

/* sizeof() of the ABI structs as compiled into the library, in declaration order
 * (0 = wpt_keyframe ... 16 = wpt_scene_info); lets a foreign-language binding verify its layout. */
size_t wpt_abi_sizeof(int which);

#define WPT_DEVICE_NONE (-1)
#define WPT_DEVICE_ALL  (-2)

/* Number of devices a scene created with WPT_DEVICE_ALL will use: every usable CUDA device, or those listed in the
 * environment variable WURBLPT_DEVICES ("all", or comma separated device indices). This is what
 * MPICoordinator::worldSize() reports (mpi.hpp:286). */
int wpt_planned_device_count(void);

/* Ingest a scene description: copies what it needs (the caller may free `desc`
 * afterwards) and uploads time-independent data to `device`. Replaces the sum of
 * Scene::take() calls (scene.hpp:86-141). */
int wpt_scene_create(const wpt_scene_desc* desc, int device, wpt_scene** scene);
/* device = WPT_DEVICE_NONE creates a host-only scene: wpt_scene_update_bvh() and the inspection
 * calls work (they are pure host code), every rendering call fails with WPT_E_NO_DEVICE.
 * device = WPT_DEVICE_ALL replicates the scene on several GPUs of the box (see wpt_planned_device_count): the
 * acceleration structure is built once on the host and uploaded to each; wpt_render() then renders ONE frame on
 * all of them. This replaces running the reference under MPI (MPICoordinatorMain / MPICoordinatorChild,
 * mpi.hpp:42-150): the block hand-out of the coordinator thread (mpi.hpp:74-107) is one atomic counter in the first
 * device's memory that the windows of every device draw blocks of 256 work items from (over NVLink for the other
 * devices), submitBlock()'s transfer of finished pixels (mpi.hpp:133-149) is one device-to-device copy of each
 * frame, added on the first device. The other entry points (queries, ground truth) use the first device. */

/* A work-item counter shared by several PROCESSES (one per GPU): 256 bytes of device memory on `device` holding two
 * counters (offsets 0 and 128, used for alternate frames so that one can be cleared while the other is in use),
 * exported / imported with CUDA IPC handles (64 bytes). The pointers go into wpt_render_opts::shared_counter. */
int wpt_ipc_counter_create(int device, void** counters, unsigned char handle[64]);
int wpt_ipc_counter_open(int device, const unsigned char handle[64], void** counters);
int wpt_ipc_counter_close(int device, void* counters, int opened);
/* Zeroes ONE of the two counters (`counter` = counters or counters + 128 bytes) on `stream` (a cudaStream_t). */
int wpt_ipc_counter_clear(void* counter, void* stream);

/* 1 + the number of replicas (0 for a host-only scene). */
int wpt_scene_device_count(const wpt_scene* scene);

/* Which builder wpt_scene_update_bvh() runs (default: host; the environment variable WURBLPT_BVH_BUILDER=device selects
 * the device builder for every scene):
 *   WPT_BUILDER_HOST_SAH     the reference's builder restated on the host (bvh.hpp:93-193): the tree equals the
 *                            reference's node for node, both walks are available
 *   WPT_BUILDER_DEVICE_LBVH  bounds on the host with the reference's formulas, tree on the device (Morton codes,
 *                            sort, binary radix tree, bottom-up boxes, 4-wide collapse; wurblpt_b200/csrc/bvh_device.cu):
 *                            milliseconds instead of 0.3 s for 1 M triangles, for scenes that are rebuilt every frame.
 *                            Closest hits are those of the host-built tree (exact distance ties may be settled differently);
 *                            only the 4-wide walk is available; scenes with media use the host builder regardless. */
enum { WPT_BUILDER_HOST_SAH = 0, WPT_BUILDER_DEVICE_LBVH = 1 };
int wpt_scene_set_builder(wpt_scene* scene, int builder);

/* Scene::updateBVH(t0, t1) (scene.hpp:151-169): bound every hitable over [t0,t1]
 * (hitable_triangle.hpp:336-395, hitable_sphere.hpp:97-110, hitable_medium.hpp:59-67),
 * build the SAH tree (bvh.hpp:93-193), derive the device layouts and upload them. */
int wpt_scene_update_bvh(wpt_scene* scene, float t0, float t1);

int wpt_scene_get_info(const wpt_scene* scene, wpt_scene_info* info);

void wpt_scene_destroy(wpt_scene* scene);

/* Inspection: the top-level tree in the reference's linear layout (BVHNodeLinear, bvh.hpp:217-246),
 * binary_nodes entries: boxes[i*6..] = lo xyz, hi xyz; inner nodes have leaf[i] = -1 and link[i] =
 * index of child 2 (child 1 is i+1); leaves have leaf[i] = global hitable id. */
int wpt_scene_get_binary_nodes(const wpt_scene* scene, uint64_t capacity, float* boxes, uint32_t* link, int32_t* leaf);

/* mcpt() (wurblpt.hpp:279-436): render samples_sqrt^2 paths per pixel into `frame`
 * (w*h*3 floats; host memory unless opts->frame_on_device). `opts` and `stats` may be NULL.
 *
 * Concurrency: a scene owns the scratch memory of its calls (counters, path state, queues, staging buffers), so
 * only ONE call per scene (wpt_render, wpt_primary_hits, wpt_trace_rays, wpt_get_ground_truth) may be in flight
 * at a time, from any thread or stream; different scenes are independent. With opts->frame_on_device and
 * stats == NULL the call returns when the work has been queued on opts->stream: synchronise that stream before
 * the next call on the same scene.
 * opts->frame_on_device: `frame` is device memory on the scene's (first) device and the caller clears it first;
 * pixels outside the tiles selected by tile_first / tile_stride are left as they are (that is what lets the frames
 * of several processes be added up). */
int wpt_render(wpt_scene* scene, const wpt_camera_desc* camera, const wpt_sensor_desc* sensor,
        const wpt_params* params, uint32_t samples_sqrt, float t0, float t1,
        const wpt_render_opts* opts, float* frame, wpt_stats* stats);

/* The primary-ray closest-hit query of getGroundTruth() (wurblpt.hpp:659-668): one ray
 * through each pixel centre at time t, randomness off. ids[pixel] = global hitable id or -1,
 * a[pixel] = hit distance (0 where id = -1). Host pointers. `counters`, if not NULL, receives
 * { tree nodes visited, leaf (hitable) tests, nested medium-boundary queries } summed over all
 * rays: with WPT_TRAVERSAL_REFERENCE these are the reference's own visit counts, the n and l of
 * the algorithmic-bytes formula. */
int wpt_primary_hits(wpt_scene* scene, const wpt_camera_desc* camera,
        uint32_t width, uint32_t height, float t, float min_hit_distance,
        const wpt_render_opts* opts, int32_t* ids, float* a, uint64_t* counters);

/* getGroundTruth() (wurblpt.hpp:643-759): per pixel, the hit record of the pixel-centre ray at time t0 as
 * a set of images. `bits` selects them (the GroundTruth::* constants, wurblpt.hpp:458-478, same bit
 * positions); out->aov[i] receives the image of bit i, WPT_GT_COMPONENTS[i] floats per pixel, pixel
 * index y*width+x like the frame; bit 19 (materials) goes to out->materials, one int32 per pixel:
 * Scene::materialIndex() of the hitable's material, taken from material_ids[description material index]
 * when material_ids is not NULL (else the description index itself), -1 where nothing was hit.
 * Pixels without a hit hold zeros. Host pointers; pointers of unselected bits are not touched.
 * Offsets to the previous / next frame need t_prev <= t0 <= t_next and, for the pixel-space ones, a
 * perspective camera (camera.hpp:205-217). The BVH must have been updated for (t0, t0). */
enum {
    WPT_GT_WORLD_POSITIONS = 1 << 0, WPT_GT_WORLD_GEOMETRY_NORMALS = 1 << 1, WPT_GT_WORLD_GEOMETRY_TANGENTS = 1 << 2,
    WPT_GT_WORLD_MATERIAL_NORMALS = 1 << 3, WPT_GT_WORLD_MATERIAL_TANGENTS = 1 << 4,
    WPT_GT_CAMERA_POSITIONS = 1 << 5, WPT_GT_CAMERA_GEOMETRY_NORMALS = 1 << 6, WPT_GT_CAMERA_GEOMETRY_TANGENTS = 1 << 7,
    WPT_GT_CAMERA_MATERIAL_NORMALS = 1 << 8, WPT_GT_CAMERA_MATERIAL_TANGENTS = 1 << 9,
    WPT_GT_CAMERA_DEPTHS = 1 << 10, WPT_GT_CAMERA_DISTANCES = 1 << 11, WPT_GT_TEXCOORDS = 1 << 12,
    WPT_GT_WORLD_OFFSET_TO_PREV = 1 << 13, WPT_GT_WORLD_OFFSET_TO_NEXT = 1 << 14,
    WPT_GT_CAMERA_OFFSET_TO_PREV = 1 << 15, WPT_GT_CAMERA_OFFSET_TO_NEXT = 1 << 16,
    WPT_GT_PIXEL_OFFSET_TO_PREV = 1 << 17, WPT_GT_PIXEL_OFFSET_TO_NEXT = 1 << 18,
    WPT_GT_MATERIALS = 1 << 19, WPT_GT_ALL = (1 << 20) - 1
};
#define WPT_GT_FLOAT_IMAGES 19
static const uint32_t WPT_GT_COMPONENTS[WPT_GT_FLOAT_IMAGES] = { 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1, 1, 2, 3, 3, 3, 3, 2, 2 };

typedef struct wpt_ground_truth {
    float* aov[WPT_GT_FLOAT_IMAGES];
    int32_t* materials;
} wpt_ground_truth;

int wpt_get_ground_truth(wpt_scene* scene, const wpt_camera_desc* camera, uint32_t width, uint32_t height,
        float t0, float t_prev, float t_next, float min_hit_distance, uint32_t bits,
        const int32_t* material_ids, uint32_t n_material_ids,
        const wpt_render_opts* opts, const wpt_ground_truth* out);

/* Closest-hit queries for caller-supplied rays (origin xyz, unit direction xyz per ray;
 * host pointers). Used to check BVH::hit() parity on secondary-ray distributions.
 * A ray with a non-finite origin or a direction whose squared length is not within 1e-5 of 1 (what the reference's Ray
 * constructor asserts, ray.hpp:47-49) is refused with WPT_E_INVALID before anything runs. */
int wpt_trace_rays(wpt_scene* scene, const float* origins, const float* directions, uint64_t n_rays,
        float time, float min_hit_distance, const wpt_render_opts* opts, int32_t* ids, float* a, uint64_t* counters);

/* Prng known-answer hook (prng.hpp:79-95): out[i*n_draws + k] = k-th in01() of Prng(seeds[i]),
 * evaluated on the device. Host pointers. */
int wpt_prng_kat(const uint32_t* seeds, uint32_t n_seeds, uint32_t n_draws, float* out);

#ifdef __cplusplus
}
#endif

#endif
