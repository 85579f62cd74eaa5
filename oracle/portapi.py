"""oracle/portapi.py -- TEST INFRASTRUCTURE (ctypes wrapper of oracle/_build/liboracle_port.so, the
CPU restatement in oracle/port.cpp). Never imported by the product package."""
import ctypes as C
import os

import numpy as np

from wurblpt_b200 import abi

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_port.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.port_prng_kat.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.port_primary_hits.argtypes = [C.POINTER(abi.wpt_scene_desc), C.POINTER(abi.wpt_camera_desc), C.c_uint32,
                                        C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
        L.port_trace_rays.argtypes = [C.POINTER(abi.wpt_scene_desc), C.c_void_p, C.c_void_p, C.c_uint64, C.c_float,
                                      C.c_void_p, C.c_void_p]
        L.port_bvh_nodes.restype = C.c_longlong
        L.port_bvh_nodes.argtypes = [C.POINTER(abi.wpt_scene_desc), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_medium_boundary.restype = C.c_longlong
        L.port_medium_boundary.argtypes = [C.POINTER(abi.wpt_scene_desc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                           C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def prng_kat(seeds, n_draws):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.zeros((len(seeds), n_draws), dtype=np.float32)
    lib().port_prng_kat(seeds.ctypes.data, len(seeds), n_draws, out.ctypes.data)
    return out


def primary_hits(desc, camera, width, height, min_hit=1e-5):
    ids = np.zeros((height, width), dtype=np.int32)
    a = np.zeros((height, width), dtype=np.float32)
    rc = lib().port_primary_hits(C.byref(desc), C.byref(camera), width, height, min_hit, ids.ctypes.data, a.ctypes.data)
    if rc != 0:
        raise NotImplementedError("oracle port does not cover this scene / camera (rc=%d)" % rc)
    return ids, a


def trace_rays(desc, origins, directions, min_hit=1e-5):
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    directions = np.ascontiguousarray(directions, dtype=np.float32)
    n = origins.shape[0]
    ids = np.zeros(n, dtype=np.int32)
    a = np.zeros(n, dtype=np.float32)
    rc = lib().port_trace_rays(C.byref(desc), origins.ctypes.data, directions.ctypes.data, n, min_hit,
                               ids.ctypes.data, a.ctypes.data)
    if rc != 0:
        raise NotImplementedError("oracle port does not cover this scene (rc=%d)" % rc)
    return ids, a


def bvh_nodes(desc):
    n = lib().port_bvh_nodes(C.byref(desc), 0, None, None, None)
    if n < 0:
        raise NotImplementedError("oracle port does not cover this scene")
    boxes = np.zeros((n, 6), dtype=np.float32)
    link = np.zeros(n, dtype=np.uint32)
    leaf = np.zeros(n, dtype=np.int32)
    lib().port_bvh_nodes(C.byref(desc), n, boxes.ctypes.data, link.ctypes.data, leaf.ctypes.data)
    return boxes, link, leaf


def medium_boundary(desc, medium, origins, directions, min_hit=1e-5):
    """The two boundary queries of HitableMedium::hit for every ray: (nested, replay, corners) - see port_medium_boundary."""
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    directions = np.ascontiguousarray(directions, dtype=np.float32)
    n = origins.shape[0]
    nested = np.zeros((n, 2), dtype=np.float32)
    replay = np.zeros((n, 2), dtype=np.float32)
    corners = np.zeros((max(1, int(desc.media[medium].boundary_count)), 3, 3), dtype=np.float32)
    count = lib().port_medium_boundary(C.byref(desc), medium, origins.ctypes.data, directions.ctypes.data, n, min_hit,
                                       nested.ctypes.data, replay.ctypes.data, corners.ctypes.data)
    if count < 0:
        raise NotImplementedError("oracle port does not cover this medium's boundary (rc=%d)" % count)
    return nested, replay, corners[:count]
