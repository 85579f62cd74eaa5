"""ctypes binding of wurblpt_b200/lib/libwurblpt_b200.so (the C ABI in include/wurblpt_b200.h).

This is the product path from Python: bench.py, __graft_entry__.smoke() and the GPU tests call
the kernels through here. It never imports anything under oracle/. If the library is missing
or no CUDA device is present, calls fail loudly (WptError); there is no fallback.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# WPT_LIB selects an experiment build (see csrc/Makefile VARIANT=...); the default is the product library
LIB_PATH = os.environ.get("WPT_LIB") or os.path.join(_HERE, "lib", "libwurblpt_b200.so")

EXPORTS = ["wpt_version", "wpt_last_error", "wpt_device_count", "wpt_abi_sizeof", "wpt_scene_create",
           "wpt_scene_update_bvh", "wpt_scene_get_info", "wpt_scene_destroy", "wpt_scene_get_binary_nodes",
           "wpt_render", "wpt_primary_hits", "wpt_get_ground_truth", "wpt_trace_rays", "wpt_prng_kat",
           "wpt_planned_device_count", "wpt_scene_device_count", "wpt_scene_set_builder",
           "wpt_ipc_counter_create", "wpt_ipc_counter_open", "wpt_ipc_counter_close", "wpt_ipc_counter_clear"]

_lib = None


class WptError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("wurblpt_b200 error %d: %s" % (code, message))
        self.code = code


def lib():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WptError(abi.WPT_E_NO_DEVICE, "%s not found: build it with `python -c 'import __graft_entry__ as g; "
                       "g.build()'` or `make -C wurblpt_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.wpt_version.restype = C.c_char_p
    L.wpt_last_error.restype = C.c_char_p
    L.wpt_abi_sizeof.restype = C.c_size_t
    L.wpt_abi_sizeof.argtypes = [C.c_int]
    L.wpt_scene_create.argtypes = [C.POINTER(abi.wpt_scene_desc), C.c_int, C.POINTER(C.c_void_p)]
    L.wpt_scene_update_bvh.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.wpt_scene_get_info.argtypes = [C.c_void_p, C.POINTER(abi.wpt_scene_info)]
    L.wpt_scene_destroy.argtypes = [C.c_void_p]
    L.wpt_scene_destroy.restype = None
    L.wpt_scene_get_binary_nodes.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.wpt_render.argtypes = [C.c_void_p, C.POINTER(abi.wpt_camera_desc), C.POINTER(abi.wpt_sensor_desc),
                             C.POINTER(abi.wpt_params), C.c_uint32, C.c_float, C.c_float,
                             C.POINTER(abi.wpt_render_opts), C.c_void_p, C.POINTER(abi.wpt_stats)]
    L.wpt_primary_hits.argtypes = [C.c_void_p, C.POINTER(abi.wpt_camera_desc), C.c_uint32, C.c_uint32, C.c_float,
                                   C.c_float, C.POINTER(abi.wpt_render_opts), C.c_void_p, C.c_void_p, C.c_void_p]
    L.wpt_get_ground_truth.argtypes = [C.c_void_p, C.POINTER(abi.wpt_camera_desc), C.c_uint32, C.c_uint32, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_void_p, C.c_uint32,
                                       C.POINTER(abi.wpt_render_opts), C.POINTER(abi.wpt_ground_truth)]
    L.wpt_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_float,
                                 C.POINTER(abi.wpt_render_opts), C.c_void_p, C.c_void_p, C.c_void_p]
    L.wpt_prng_kat.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.wpt_scene_device_count.argtypes = [C.c_void_p]
    L.wpt_scene_set_builder.argtypes = [C.c_void_p, C.c_int]
    L.wpt_ipc_counter_create.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
    L.wpt_ipc_counter_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    L.wpt_ipc_counter_close.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.wpt_ipc_counter_clear.argtypes = [C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _check(rc):
    if rc != abi.WPT_OK:
        raise WptError(rc, lib().wpt_last_error().decode())


def device_count():
    return lib().wpt_device_count()


def ipc_counter_create(device):
    """(device pointer, 64-byte handle) of a pair of work-item counters other processes can open."""
    p = C.c_void_p()
    handle = C.create_string_buffer(64)
    _check(lib().wpt_ipc_counter_create(device, C.byref(p), handle))
    return p.value, handle.raw


def ipc_counter_open(device, handle):
    p = C.c_void_p()
    _check(lib().wpt_ipc_counter_open(device, handle, C.byref(p)))
    return p.value


def ipc_counter_close(device, pointer, opened):
    _check(lib().wpt_ipc_counter_close(device, C.c_void_p(pointer), 1 if opened else 0))


def ipc_counter_clear(pointer, stream=None):
    _check(lib().wpt_ipc_counter_clear(C.c_void_p(pointer), C.c_void_p(stream) if stream else None))


def make_opts(traversal=abi.WPT_TRAVERSAL_WIDE, substreams=0, tile_first=0, tile_stride=1,
              frame_on_device=False, stream=None, pipeline=abi.WPT_PIPELINE_AUTO, shared_counter=None):
    o = abi.wpt_render_opts()
    o.pipeline = pipeline
    o.traversal = traversal
    o.substreams = substreams
    o.tile_first = tile_first
    o.tile_stride = tile_stride
    o.frame_on_device = 1 if frame_on_device else 0
    o.stream = stream
    o.shared_counter = shared_counter
    return o


class Scene:
    """A scene resident on one GPU (wpt_scene). `desc` is a wpt_scene_desc (e.g. HostScene.desc)."""

    def __init__(self, desc, device=0):
        self._h = C.c_void_p()
        self.device = device
        _check(lib().wpt_scene_create(C.byref(desc), device, C.byref(self._h)))

    def set_builder(self, builder):
        """abi.WPT_BUILDER_HOST_SAH (default) or abi.WPT_BUILDER_DEVICE_LBVH for the following update_bvh() calls."""
        _check(lib().wpt_scene_set_builder(self._h, builder))
        return self

    def update_bvh(self, t0=0.0, t1=0.0):
        """Scene::updateBVH(t0, t1)."""
        _check(lib().wpt_scene_update_bvh(self._h, t0, t1))
        return self

    def device_count(self):
        """Devices that render a frame of this scene (more than one only for device=abi.WPT_DEVICE_ALL)."""
        return lib().wpt_scene_device_count(self._h)

    def info(self):
        i = abi.wpt_scene_info()
        _check(lib().wpt_scene_get_info(self._h, C.byref(i)))
        return i

    def binary_nodes(self):
        n = self.info().binary_nodes
        boxes = np.zeros((n, 6), dtype=np.float32)
        link = np.zeros(n, dtype=np.uint32)
        leaf = np.zeros(n, dtype=np.int32)
        _check(lib().wpt_scene_get_binary_nodes(self._h, n, boxes.ctypes.data, link.ctypes.data, leaf.ctypes.data))
        return boxes, link, leaf

    def render(self, camera, sensor, params, samples_sqrt, t0=0.0, t1=0.0, opts=None, frame=None):
        """mcpt(): returns (frame[h, w, 3] float32 host array, wpt_stats)."""
        if frame is None:
            frame = np.zeros((sensor.height, sensor.width, 3), dtype=np.float32)
        stats = abi.wpt_stats()
        _check(lib().wpt_render(self._h, C.byref(camera), C.byref(sensor), C.byref(params), samples_sqrt, t0, t1,
                                C.byref(opts) if opts is not None else None, frame.ctypes.data, C.byref(stats)))
        return frame, stats

    def render_device(self, camera, sensor, params, samples_sqrt, frame_ptr, t0=0.0, t1=0.0, opts=None,
                      want_stats=False):
        """mcpt() into a device frame (pointer). Asynchronous on opts.stream unless want_stats."""
        if opts is None:
            opts = make_opts(frame_on_device=True)
        opts.frame_on_device = 1
        stats = abi.wpt_stats() if want_stats else None
        _check(lib().wpt_render(self._h, C.byref(camera), C.byref(sensor), C.byref(params), samples_sqrt, t0, t1,
                                C.byref(opts), C.c_void_p(frame_ptr), C.byref(stats) if want_stats else None))
        return stats

    def primary_hits(self, camera, width, height, t=0.0, min_hit=1e-5, opts=None, count=False):
        ids = np.zeros((height, width), dtype=np.int32)
        a = np.zeros((height, width), dtype=np.float32)
        counters = (C.c_uint64 * 3)()
        _check(lib().wpt_primary_hits(self._h, C.byref(camera), width, height, t, min_hit,
                                      C.byref(opts) if opts is not None else None, ids.ctypes.data, a.ctypes.data,
                                      counters if count else None))
        if count:
            return ids, a, dict(nodes=counters[0], leaves=counters[1], nested=counters[2])
        return ids, a

    def ground_truth(self, camera, width, height, t0=0.0, t_prev=None, t_next=None, min_hit=1e-5,
                     bits=abi.WPT_GT_ALL, material_ids=None, opts=None):
        """getGroundTruth(): dict of images (abi.WPT_GT_NAMES + "materials") for the selected bits."""
        images, out = abi.ground_truth_buffers(width, height, bits)
        ids = None if material_ids is None else np.ascontiguousarray(material_ids, dtype=np.int32)
        _check(lib().wpt_get_ground_truth(self._h, C.byref(camera), width, height, t0,
                                          t0 if t_prev is None else t_prev, t0 if t_next is None else t_next,
                                          min_hit, bits, None if ids is None else ids.ctypes.data,
                                          0 if ids is None else ids.size,
                                          C.byref(opts) if opts is not None else None, C.byref(out)))
        return images

    def trace_rays(self, origins, directions, time=0.0, min_hit=1e-5, opts=None, count=False):
        origins = np.ascontiguousarray(origins, dtype=np.float32)
        directions = np.ascontiguousarray(directions, dtype=np.float32)
        n = origins.shape[0]
        ids = np.zeros(n, dtype=np.int32)
        a = np.zeros(n, dtype=np.float32)
        counters = (C.c_uint64 * 3)()
        _check(lib().wpt_trace_rays(self._h, origins.ctypes.data, directions.ctypes.data, n, time, min_hit,
                                    C.byref(opts) if opts is not None else None, ids.ctypes.data, a.ctypes.data,
                                    counters if count else None))
        if count:
            return ids, a, dict(nodes=counters[0], leaves=counters[1], nested=counters[2])
        return ids, a

    def close(self):
        if self._h:
            lib().wpt_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prng_kat(seeds, n_draws):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.zeros((len(seeds), n_draws), dtype=np.float32)
    _check(lib().wpt_prng_kat(seeds.ctypes.data, len(seeds), n_draws, out.ctypes.data))
    return out
