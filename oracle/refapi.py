"""oracle/refapi.py -- TEST INFRASTRUCTURE (ctypes wrapper of oracle/_ref/libwpt_ref_*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this. The product package (wurblpt_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

from wurblpt_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(flavour):
    return os.path.join(_HERE, "_ref", "libwpt_ref_%s.so" % flavour)


def available(flavour="strict"):
    return os.path.exists(lib_path(flavour))


def load(flavour="strict"):
    if flavour in _LIBS:
        return _LIBS[flavour]
    lib = C.CDLL(lib_path(flavour), mode=getattr(os, "RTLD_LOCAL", 0))
    lib.ref_last_error.restype = C.c_char_p
    lib.ref_scene_create.argtypes = [C.c_char_p, C.c_float, C.c_int]
    lib.ref_scene_desc.restype = C.POINTER(abi.wpt_scene_desc)
    lib.ref_scene_desc.argtypes = [C.c_int]
    lib.ref_scene_update_bvh.argtypes = [C.c_int, C.c_float, C.c_float]
    lib.ref_scene_destroy.argtypes = [C.c_int]
    lib.ref_default_view.argtypes = [C.c_char_p, C.POINTER(abi.wpt_camera_desc), C.POINTER(abi.wpt_sensor_desc),
                                     C.POINTER(abi.wpt_params), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.ref_render.restype = C.c_double
    lib.ref_render.argtypes = [C.c_int, C.POINTER(abi.wpt_camera_desc), C.POINTER(abi.wpt_sensor_desc),
                               C.POINTER(abi.wpt_params), C.c_uint32, C.c_float, C.c_float, C.c_void_p]
    lib.ref_primary_hits.argtypes = [C.c_int, C.POINTER(abi.wpt_camera_desc), C.c_uint32, C.c_uint32,
                                     C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.ref_ground_truth.argtypes = [C.c_int, C.POINTER(abi.wpt_camera_desc), C.c_uint32, C.c_uint32, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.POINTER(abi.wpt_ground_truth)]
    lib.ref_trace_rays.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p]
    lib.ref_camera_rays.argtypes = [C.POINTER(abi.wpt_camera_desc), C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
    lib.ref_prng_kat.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.ref_build_flags.restype = C.c_char_p
    lib.ref_counters.argtypes = [C.c_void_p, C.c_int]
    _LIBS[flavour] = lib
    return lib


class View:
    """Default camera / sensor / parameters of a named config (scenes/configs.hpp)."""

    def __init__(self, name, flavour="strict"):
        lib = load(flavour)
        self.camera = abi.wpt_camera_desc()
        self.sensor = abi.wpt_sensor_desc()
        self.params = abi.wpt_params()
        s, t0, t1 = C.c_uint32(), C.c_float(), C.c_float()
        lib.ref_default_view(name.encode(), C.byref(self.camera), C.byref(self.sensor), C.byref(self.params),
                             C.byref(s), C.byref(t0), C.byref(t1))
        # copy camera key frames out of the library's static storage
        self._keyframes = None
        if self.camera.n_keyframes:
            self._keyframes = (abi.wpt_keyframe * self.camera.n_keyframes)()
            C.memmove(self._keyframes, self.camera.keyframes, C.sizeof(self._keyframes))
            self.camera.keyframes = C.cast(self._keyframes, C.POINTER(abi.wpt_keyframe))
        self.samples_sqrt, self.t0, self.t1 = s.value, t0.value, t1.value

    def resized(self, width, height):
        """Same view at another resolution of the same aspect ratio class."""
        self.sensor.width, self.sensor.height = width, height
        return self


class RefScene:
    """A scene built and rendered by the unmodified reference."""

    def __init__(self, name, scale=1.0, variant=0, flavour="strict"):
        self.lib = load(flavour)
        self.name = name
        self.h = self.lib.ref_scene_create(name.encode(), scale, variant)
        if self.h < 0:
            raise RuntimeError("ref_scene_create(%s): %s" % (name, self.lib.ref_last_error().decode()))

    @property
    def desc(self):
        return self.lib.ref_scene_desc(self.h).contents

    def update_bvh(self, t0=0.0, t1=0.0):
        self.lib.ref_scene_update_bvh(self.h, t0, t1)

    def render(self, camera, sensor, params, samples_sqrt, t0=0.0, t1=0.0):
        frame = np.zeros((sensor.height, sensor.width, 3), dtype=np.float32)
        seconds = self.lib.ref_render(self.h, C.byref(camera), C.byref(sensor), C.byref(params), samples_sqrt,
                                      t0, t1, frame.ctypes.data)
        if seconds < 0:
            raise RuntimeError("ref_render failed")
        return frame, seconds

    def primary_hits(self, camera, width, height, t=0.0, min_hit=1e-5):
        ids = np.zeros((height, width), dtype=np.int32)
        a = np.zeros((height, width), dtype=np.float32)
        self.lib.ref_primary_hits(self.h, C.byref(camera), width, height, t, min_hit, ids.ctypes.data, a.ctypes.data)
        return ids, a

    def ground_truth(self, camera, width, height, t0=0.0, t_prev=None, t_next=None, min_hit=1e-5):
        """getGroundTruth() of the reference, all images (pixel offsets are zero for surround / stereo cameras)."""
        images, out = abi.ground_truth_buffers(width, height)
        rc = self.lib.ref_ground_truth(self.h, C.byref(camera), width, height, t0, t0 if t_prev is None else t_prev,
                                       t0 if t_next is None else t_next, min_hit, C.byref(out))
        if rc != 0:
            raise RuntimeError("ref_ground_truth failed")
        return images

    def trace_rays(self, origins, directions, time=0.0, min_hit=1e-5):
        origins = np.ascontiguousarray(origins, dtype=np.float32)
        directions = np.ascontiguousarray(directions, dtype=np.float32)
        n = origins.shape[0]
        ids = np.zeros(n, dtype=np.int32)
        a = np.zeros(n, dtype=np.float32)
        self.lib.ref_trace_rays(self.h, origins.ctypes.data, directions.ctypes.data, n, time, min_hit,
                                ids.ctypes.data, a.ctypes.data)
        return ids, a

    def counters(self, reset=True):
        out = (C.c_uint64 * 4)()
        have = self.lib.ref_counters(out, 1 if reset else 0)
        return dict(have=bool(have), rays=out[0], calls=out[1], nodes=out[2], leaves=out[3])

    def close(self):
        if self.h >= 0:
            self.lib.ref_scene_destroy(self.h)
            self.h = -1


def camera_rays(camera, width, height, t=0.0, flavour="strict"):
    out = np.zeros((height * width, 6), dtype=np.float32)
    load(flavour).ref_camera_rays(C.byref(camera), width, height, t, out.ctypes.data)
    return out[:, :3].copy(), out[:, 3:].copy()


def prng_kat(seeds, n_draws, flavour="strict"):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.zeros((len(seeds), n_draws), dtype=np.float32)
    load(flavour).ref_prng_kat(seeds.ctypes.data, len(seeds), n_draws, out.ctypes.data)
    return out


def bvh_nodes(scene):
    """The reference's linearised BVH of a RefScene: (boxes[n,6], link[n], leaf[n])."""
    lib = scene.lib
    lib.ref_bvh_nodes.restype = C.c_longlong
    lib.ref_bvh_nodes.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    n = lib.ref_bvh_nodes(scene.h, 0, None, None, None)
    if n < 0:
        raise RuntimeError("reference BVH not built")
    boxes = np.zeros((n, 6), dtype=np.float32)
    link = np.zeros(n, dtype=np.uint32)
    leaf = np.zeros(n, dtype=np.int32)
    lib.ref_bvh_nodes(scene.h, n, boxes.ctypes.data, link.ctypes.data, leaf.ctypes.data)
    return boxes, link, leaf
