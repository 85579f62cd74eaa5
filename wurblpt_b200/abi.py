"""ctypes mirror of include/wurblpt_b200.h (the C ABI of the path-tracing core).

Field order and types must match the header exactly; tests/test_abi.py checks the
struct sizes against the compiled library (wpt_abi_sizeof).
"""
import ctypes as C

WPT_ABI_VERSION = 1

WPT_OK = 0
WPT_E_INVALID = -1
WPT_E_NO_DEVICE = -2
WPT_E_CUDA = -3
WPT_E_UNSUPPORTED = -4
WPT_E_STATE = -5
WPT_DEVICE_NONE, WPT_DEVICE_ALL = -1, -2
WPT_BUILDER_HOST_SAH, WPT_BUILDER_DEVICE_LBVH = 0, 1

WPT_TEX_CONSTANT, WPT_TEX_CHECKER, WPT_TEX_IMAGE, WPT_TEX_TRANSFORMER = range(4)
WPT_TEXEL_U8, WPT_TEXEL_U16, WPT_TEXEL_F32 = range(3)
(WPT_MAT_NONE, WPT_MAT_LAMBERTIAN, WPT_MAT_MIRROR, WPT_MAT_GLASS, WPT_MAT_GGX, WPT_MAT_MODPHONG,
 WPT_MAT_PHASE_ISOTROPIC, WPT_MAT_LIGHT_DIFFUSE, WPT_MAT_LIGHT_SPOT, WPT_MAT_LIGHT_TOF,
 WPT_MAT_TWOSIDED) = range(11)
WPT_COMP_MESH_INSTANCE, WPT_COMP_SPHERE, WPT_COMP_MEDIUM = range(3)
WPT_COMPF_HITABLE, WPT_COMPF_HOTSPOT = 1, 2
WPT_ENV_NONE, WPT_ENV_EQUIRECT, WPT_ENV_CUBE = range(3)
WPT_SURROUND_OFF, WPT_SURROUND_180, WPT_SURROUND_360 = range(3)
WPT_DISTORTION_NONE, WPT_DISTORTION_RADIAL_AND_PLANAR, WPT_DISTORTION_RADIAL_ONLY, WPT_DISTORTION_OPENCV = range(4)
WPT_SENSOR_RGB, WPT_SENSOR_TOF_AMCW = range(2)
WPT_TRAVERSAL_WIDE, WPT_TRAVERSAL_REFERENCE, WPT_TRAVERSAL_POOL, WPT_TRAVERSAL_PERSISTENT, WPT_TRAVERSAL_QUAD = range(5)
WPT_PIPELINE_AUTO, WPT_PIPELINE_MEGAKERNEL, WPT_PIPELINE_WAVEFRONT = range(3)

f32, i32, u32, u64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint64


class wpt_keyframe(C.Structure):
    _fields_ = [("t", f32), ("translation", f32 * 3), ("rotation", f32 * 4), ("scaling", f32 * 3)]


class wpt_animation(C.Structure):
    _fields_ = [("first_keyframe", u32), ("keyframe_count", u32)]


class wpt_texture(C.Structure):
    _fields_ = [("type", i32), ("src_texture", i32), ("width", i32), ("height", i32),
                ("components", i32), ("texel_type", i32), ("linearize_srgb", i32),
                ("horiz", i32), ("vert", i32), ("_pad", i32), ("texel_offset", u64),
                ("color0", f32 * 4), ("color1", f32 * 4), ("coord_factor", f32 * 2),
                ("coord_offset", f32 * 2), ("val_factor", f32 * 4), ("val_offset", f32 * 4)]


class wpt_material(C.Structure):
    _fields_ = [("type", i32), ("normal_tex", i32), ("tex", i32 * 5), ("flags", u32),
                ("front", i32), ("back", i32), ("_pad", i32 * 2),
                ("v0", f32 * 4), ("v1", f32 * 4), ("v2", f32 * 4), ("v3", f32 * 4), ("f", f32 * 4)]


class wpt_mesh(C.Structure):
    _fields_ = [("vertex_offset", u64), ("index_offset", u64), ("vertex_count", u32),
                ("vertex_stride", u32), ("triangle_count", u32), ("flags", u32)]


class wpt_instance(C.Structure):
    _fields_ = [("mesh", i32), ("material", i32), ("animation", i32), ("has_transform", i32),
                ("M", f32 * 16), ("N", f32 * 9), ("_pad", f32 * 3)]


class wpt_sphere(C.Structure):
    _fields_ = [("center", f32 * 3), ("radius", f32), ("rotation", f32 * 4),
                ("material", i32), ("animation", i32), ("_pad", i32 * 2)]


class wpt_medium(C.Structure):
    _fields_ = [("density", f32), ("material", i32), ("first_boundary", u32), ("boundary_count", u32)]


class wpt_component(C.Structure):
    _fields_ = [("type", i32), ("ref", i32), ("flags", u32), ("_pad", u32)]


class wpt_envmap(C.Structure):
    _fields_ = [("type", i32), ("compat", i32), ("tex", i32 * 6), ("is_n", i32), ("_pad", i32),
                ("M", C.POINTER(f32)), ("Ms", C.POINTER(i32)), ("Mcs", C.POINTER(f32))]


class wpt_scene_desc(C.Structure):
    _fields_ = [("abi_version", u32), ("n_animations", u32),
                ("animations", C.POINTER(wpt_animation)),
                ("n_keyframes", u32), ("n_textures", u32),
                ("keyframes", C.POINTER(wpt_keyframe)), ("textures", C.POINTER(wpt_texture)),
                ("texel_bytes", u64), ("texels", C.POINTER(C.c_uint8)),
                ("n_materials", u32), ("n_meshes", u32),
                ("materials", C.POINTER(wpt_material)), ("meshes", C.POINTER(wpt_mesh)),
                ("n_vertex_floats", u64), ("vertices", C.POINTER(f32)),
                ("n_indices", u64), ("indices", C.POINTER(u32)),
                ("n_instances", u32), ("n_spheres", u32),
                ("instances", C.POINTER(wpt_instance)), ("spheres", C.POINTER(wpt_sphere)),
                ("n_media", u32), ("n_boundary", u32),
                ("media", C.POINTER(wpt_medium)), ("boundary_hitables", C.POINTER(u32)),
                ("n_components", u32), ("_pad", u32),
                ("components", C.POINTER(wpt_component)),
                ("envmap", wpt_envmap)]


class wpt_camera_desc(C.Structure):
    _fields_ = [("surround_mode", i32), ("stereoscopic_distance", f32),
                ("l", f32), ("r", f32), ("b", f32), ("t", f32),
                ("distortion_type", i32),
                ("k1", f32), ("k2", f32), ("k3", f32), ("p1", f32), ("p2", f32),
                ("b1", f32), ("b2", f32), ("b3", f32), ("b4", f32),
                ("lens_radius", f32), ("focus_dist", f32),
                ("pose_translation", f32 * 3), ("pose_rotation", f32 * 4), ("pose_scaling", f32 * 3),
                ("n_keyframes", u32), ("keyframes", C.POINTER(wpt_keyframe))]


class wpt_sensor_desc(C.Structure):
    _fields_ = [("type", i32), ("width", u32), ("height", u32),
                ("min_dist_to_light", f32), ("max_dist_to_light", f32),
                ("min_path_len", f32), ("max_path_len", f32),
                ("pixel_area", f32), ("exposure_time", f32), ("contrast", f32),
                ("phase_index", u32), ("phase_count", u32), ("_pad", u32),
                ("modulation_frequency", C.c_double)]


class wpt_params(C.Structure):
    _fields_ = [("max_path_components", u32), ("rr_threshold", f32),
                ("randomize_ray_over_pixel", i32), ("min_hit_distance", f32)]


class wpt_render_opts(C.Structure):
    _fields_ = [("traversal", i32), ("substreams", i32), ("tile_first", u32), ("tile_stride", u32),
                ("frame_on_device", i32), ("pipeline", i32), ("stream", C.c_void_p), ("shared_counter", C.c_void_p)]


class wpt_stats(C.Structure):
    _fields_ = [("samples", u64), ("rays", u64), ("nested_queries", u64), ("tiles", u64),
                ("kernel_ms", f32), ("total_ms", f32), ("launches", u32), ("_pad", u32)]


class wpt_scene_info(C.Structure):
    _fields_ = [("hitables", u64), ("triangles", u64), ("spheres", u64), ("media", u64),
                ("binary_nodes", u64), ("wide_nodes", u64),
                ("binary_depth", u32), ("wide_depth", u32), ("hot_spots", u32), ("warp_tested_media", u32),
                ("device_bytes", u64), ("build_seconds", C.c_double), ("upload_seconds", C.c_double)]


# getGroundTruth() images: bit i selects image i (GroundTruth::*, reference wurblpt.hpp:458-478)
WPT_GT_NAMES = ["world_positions", "world_geometry_normals", "world_geometry_tangents", "world_material_normals",
                "world_material_tangents", "camera_positions", "camera_geometry_normals", "camera_geometry_tangents",
                "camera_material_normals", "camera_material_tangents", "camera_depths", "camera_distances", "texcoords",
                "world_offset_to_prev", "world_offset_to_next", "camera_offset_to_prev", "camera_offset_to_next",
                "pixel_offset_to_prev", "pixel_offset_to_next"]
WPT_GT_COMPONENTS = [3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1, 1, 2, 3, 3, 3, 3, 2, 2]
WPT_GT_FLOAT_IMAGES = 19
WPT_GT_MATERIALS = 1 << 19
WPT_GT_ALL = (1 << 20) - 1
WPT_GT_PIXEL_OFFSETS = (1 << 17) | (1 << 18)


class wpt_ground_truth(C.Structure):
    _fields_ = [("aov", C.c_void_p * WPT_GT_FLOAT_IMAGES), ("materials", C.c_void_p)]


def ground_truth_buffers(width, height, bits=WPT_GT_ALL):
    """numpy images for the selected bits and the wpt_ground_truth that points at them."""
    import numpy as np
    images = {}
    out = wpt_ground_truth()
    for i, (name, comps) in enumerate(zip(WPT_GT_NAMES, WPT_GT_COMPONENTS)):
        if bits & (1 << i):
            images[name] = np.zeros((height, width, comps), dtype=np.float32)
            out.aov[i] = images[name].ctypes.data
    if bits & WPT_GT_MATERIALS:
        images["materials"] = np.full((height, width), -1, dtype=np.int32)
        out.materials = images["materials"].ctypes.data
    return images, out


STRUCTS = [wpt_keyframe, wpt_animation, wpt_texture, wpt_material, wpt_mesh, wpt_instance,
           wpt_sphere, wpt_medium, wpt_component, wpt_envmap, wpt_scene_desc, wpt_camera_desc,
           wpt_sensor_desc, wpt_params, wpt_render_opts, wpt_stats, wpt_scene_info]


def default_params():
    """Parameters() defaults (reference wurblpt.hpp:89-95)."""
    return wpt_params(128, 1.0, 1, 0.00001)


def rgb_sensor(width, height):
    """SensorRGB(w, h) with the default (open) gates (reference sensor_rgb.hpp:41-46)."""
    s = wpt_sensor_desc()
    s.type = WPT_SENSOR_RGB
    s.width, s.height = width, height
    fmax = 3.4028234663852886e38
    s.min_dist_to_light, s.max_dist_to_light = 0.0, fmax
    s.min_path_len, s.max_path_len = 0.0, fmax
    return s
