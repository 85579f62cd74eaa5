"""Scene descriptions (wpt_scene_desc) as numpy arrays: copy out, rebuild, save, load, compare.

A description is a set of plain arrays (include/wurblpt_b200.h). Holding it as a dict of numpy
arrays makes it easy to store small scenes as fixtures (tests/golden/*.npz), to compare the
flattening done by the C++ API mirror with the one done from the reference's own Scene objects,
and to hand a scene to the library from Python.
"""
import ctypes as C

import numpy as np

from . import abi

# (count field, pointer field, ctypes element type)
_ARRAYS = [
    ("n_animations", "animations", abi.wpt_animation),
    ("n_keyframes", "keyframes", abi.wpt_keyframe),
    ("n_textures", "textures", abi.wpt_texture),
    ("texel_bytes", "texels", C.c_uint8),
    ("n_materials", "materials", abi.wpt_material),
    ("n_meshes", "meshes", abi.wpt_mesh),
    ("n_vertex_floats", "vertices", C.c_float),
    ("n_indices", "indices", C.c_uint32),
    ("n_instances", "instances", abi.wpt_instance),
    ("n_spheres", "spheres", abi.wpt_sphere),
    ("n_media", "media", abi.wpt_medium),
    ("n_boundary", "boundary_hitables", C.c_uint32),
    ("n_components", "components", abi.wpt_component),
]


def _copy_out(ptr, count, ctype):
    """Copy `count` elements of `ctype` at `ptr` into a fresh uint8-backed numpy array."""
    nbytes = int(count) * C.sizeof(ctype)
    buf = np.zeros(nbytes, dtype=np.uint8)
    if nbytes:
        C.memmove(buf.ctypes.data, ptr, nbytes)
    return buf


def desc_to_arrays(desc):
    """Deep-copy a wpt_scene_desc into {name: np.uint8 array of raw bytes} (+ envmap fields)."""
    out = {}
    for nfield, pfield, ctype in _ARRAYS:
        out[pfield] = _copy_out(getattr(desc, pfield), getattr(desc, nfield), ctype)
    env = desc.envmap
    out["env_header"] = np.array([env.type, env.compat] + list(env.tex) + [env.is_n], dtype=np.int32)
    n2 = env.is_n * env.is_n
    out["env_M"] = _copy_out(env.M, n2, C.c_float)
    out["env_Ms"] = _copy_out(env.Ms, n2, C.c_int32)
    out["env_Mcs"] = _copy_out(env.Mcs, n2, C.c_float)
    return out


class HostScene:
    """A wpt_scene_desc whose arrays are owned by numpy buffers held by this object."""

    def __init__(self, arrays):
        self.arrays = {k: np.ascontiguousarray(v) for k, v in arrays.items()}
        d = abi.wpt_scene_desc()
        d.abi_version = abi.WPT_ABI_VERSION
        for nfield, pfield, ctype in _ARRAYS:
            buf = self.arrays[pfield]
            setattr(d, nfield, buf.nbytes // C.sizeof(ctype))
            setattr(d, pfield, C.cast(buf.ctypes.data, C.POINTER(ctype)))
        h = self.arrays["env_header"]
        d.envmap.type, d.envmap.compat = int(h[0]), int(h[1])
        for i in range(6):
            d.envmap.tex[i] = int(h[2 + i])
        d.envmap.is_n = int(h[8])
        d.envmap.M = C.cast(self.arrays["env_M"].ctypes.data, C.POINTER(C.c_float))
        d.envmap.Ms = C.cast(self.arrays["env_Ms"].ctypes.data, C.POINTER(C.c_int32))
        d.envmap.Mcs = C.cast(self.arrays["env_Mcs"].ctypes.data, C.POINTER(C.c_float))
        self.desc = d

    @classmethod
    def from_desc(cls, desc):
        return cls(desc_to_arrays(desc))

    @classmethod
    def load(cls, path):
        with np.load(path) as z:
            return cls({k: z[k] for k in z.files})

    def save(self, path):
        np.savez_compressed(path, **self.arrays)


def struct_view(buf, ctype):
    """View a raw byte buffer as a ctypes array of `ctype` (no copy)."""
    n = buf.nbytes // C.sizeof(ctype)
    return (ctype * n).from_buffer(buf)


def diff_descs(a, b):
    """Names of arrays that differ between two array dicts (empty list = identical)."""
    bad = []
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b or a[k].shape != b[k].shape or not np.array_equal(a[k], b[k]):
            bad.append(k)
    return bad
