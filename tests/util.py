"""Shared helpers for the tests: golden fixtures, image metrics."""
import ctypes as C
import os

import numpy as np

from wurblpt_b200 import abi
from wurblpt_b200 import scene_desc as sd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_SCENES = ["cornell", "rtiow", "mis", "mis_nohot", "media", "tof", "furnace", "mixed", "synth", "synth360", "imported"]


class Golden:
    """One tests/golden/<name>.npz: scene description, view, and the reference's outputs."""

    def __init__(self, name):
        self.name = name
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            self.z = {k: z[k] for k in z.files}
        self.host = sd.HostScene({k[len("scene_"):]: v for k, v in self.z.items() if k.startswith("scene_")})
        self.camera = abi.wpt_camera_desc.from_buffer_copy(self.z["camera"].tobytes())
        self._keys = None
        if self.camera.n_keyframes:
            self._keys = np.ascontiguousarray(self.z["camera_keyframes"])
            self.camera.keyframes = C.cast(self._keys.ctypes.data, C.POINTER(abi.wpt_keyframe))
        else:
            self.camera.keyframes = None
        self.sensor = abi.wpt_sensor_desc.from_buffer_copy(self.z["sensor"].tobytes())
        self.params = abi.wpt_params.from_buffer_copy(self.z["params"].tobytes())
        self.t0 = float(self.z["t0"])
        self.t1 = float(self.z["t1"])
        self.scale = float(self.z["scale"])

    @property
    def desc(self):
        return self.host.desc


def box_downsample(img, k):
    h, w = img.shape[:2]
    h2, w2 = (h // k) * k, (w // k) * k
    return img[:h2, :w2].reshape(h2 // k, k, w2 // k, k, -1).mean(axis=(1, 3))


def rel_rmse(a, b, floor=1e-2):
    """Relative RMSE of a against b: sqrt(mean((a-b)^2 / (b^2 + floor)))."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2 / (b ** 2 + floor))))


def compare_hits(ids, dist, ref_ids, ref_dist):
    """Classify per-ray differences between a device closest-hit result and the reference's.

    Returns dict(total, same, ties, near, wrong):
      same   identical id
      ties   different id, bit-equal (or 1-ulp) distance: an exact tie settled differently
      near   different id, distances within 4 ulp or 1e-6 relative: rounding at an edge
      wrong  everything else (must be 0)
    """
    ids = ids.ravel(); ref_ids = ref_ids.ravel()
    dist = dist.ravel(); ref_dist = ref_dist.ravel()
    diff = ids != ref_ids
    n_diff = int(diff.sum())
    ties = near = wrong = 0
    if n_diff:
        d = dist[diff]; r = ref_dist[diff]
        both = (ids[diff] >= 0) & (ref_ids[diff] >= 0)
        ulp = np.abs(d.view(np.int32).astype(np.int64) - r.view(np.int32).astype(np.int64))
        is_tie = both & (ulp <= 1)
        is_near = both & ~is_tie & ((ulp <= 4) | (np.abs(d - r) <= 1e-6 * np.abs(r)))
        ties = int(is_tie.sum()); near = int(is_near.sum())
        wrong = n_diff - ties - near
    same = ids.size - n_diff
    bit_equal_dist = int((dist[~diff].view(np.int32) == ref_dist[~diff].view(np.int32)).sum())
    return dict(total=int(ids.size), same=same, ties=ties, near=near, wrong=wrong, dist_bit_equal=bit_equal_dist)


def read_png(path):
    """Reader for 8-bit non-interlaced PNG files with filter type 0 rows, which is what the libtgd stand-in writes
    (include/tgd/png.hpp). Returns (height, width, components) uint8 with row 0 = BOTTOM row, like every other
    array the stand-in stores."""
    import struct
    import zlib
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG file"
    pos, idat, header = 8, b"", None
    while pos < len(data):
        n = struct.unpack(">I", data[pos:pos + 4])[0]
        kind = data[pos + 4:pos + 8]
        if kind == b"IHDR":
            header = struct.unpack(">IIBBBBB", data[pos + 8:pos + 8 + 13])
        elif kind == b"IDAT":
            idat += data[pos + 8:pos + 8 + n]
        pos += 12 + n
    w, h, depth, colour, _, _, interlace = header
    comps = {0: 1, 2: 3, 4: 2, 6: 4}[colour]
    assert depth == 8 and interlace == 0
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + w * comps)
    assert (rows[:, 0] == 0).all(), "only filter type 0 is handled here"
    return rows[::-1, 1:].reshape(h, w, comps).copy()


def read_wptarray(path):
    """Reader for the container the libtgd stand-in writes (include/tgd/io.hpp)."""
    with open(path, "rb") as f:
        header = f.readline().decode().split()
        assert header[0] == "WPTARRAY"
        dtype = {"float32": np.float32, "uint8": np.uint8, "uint16": np.uint16, "int32": np.int32}[header[1]]
        comps, ndims = int(header[2]), int(header[3])
        dims = [int(x) for x in header[4:4 + ndims]]
        data = np.frombuffer(f.read(), dtype=dtype)
    return data.reshape(dims[1], dims[0], comps)


VIEW_GOLDENS = ["cornell_radial_planar", "cornell_radial_only", "cornell_opencv", "mixed_opencv", "cornell_surround180",
                "mixed_surround180_stereo", "cornell_gated", "mis_gated"]


class GoldenView(Golden):
    """tests/golden/view_<name>.npz (make_view_golden.py): a base fixture's scene seen through another camera /
    sensor, with the reference's primary hits and render for that view."""

    def __init__(self, name):
        with np.load(os.path.join(GOLDEN, "view_" + name + ".npz")) as z:
            v = {k: z[k] for k in z.files}
        super().__init__(str(v["base"]))
        self.view_name = name
        self.camera = abi.wpt_camera_desc.from_buffer_copy(v["camera"].tobytes())
        self.camera.keyframes = None
        self.sensor = abi.wpt_sensor_desc.from_buffer_copy(v["sensor"].tobytes())
        self.params = abi.wpt_params.from_buffer_copy(v["params"].tobytes())
        self.t0, self.t1 = float(v["t0"]), float(v["t1"])
        for k in ("hit_ids", "hit_dist", "render", "render_spp_sqrt"):
            self.z[k] = v[k]
