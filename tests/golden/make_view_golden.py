"""Regenerates tests/golden/view_*.npz from the compiled, unmodified reference (oracle/_ref): the scenes of
make_golden.py seen through cameras and sensors the default views do not use --

  the three LensDistortion models (optics.hpp:226-309; the OpenCV one undistorts iteratively),
  Camera::Surround_180 (camera.hpp:160-169),
  SensorRGB with its path-length and distance-to-light gates set (sensor_rgb.hpp:63-80).

Each file holds the camera / sensor / params structs, the reference's primary-ray closest hits through pixel
centres, and a low-resolution reference render; the scene comes from the base fixture <scene>.npz.

    make -C oracle ref && python tests/golden/make_view_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refapi  # noqa: E402
from wurblpt_b200 import abi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def radial_and_planar(view):
    c = view.camera
    c.distortion_type = abi.WPT_DISTORTION_RADIAL_AND_PLANAR
    c.k1, c.k2, c.p1, c.p2 = -0.21, 0.06, 0.004, -0.003


def radial_only(view):
    c = view.camera
    c.distortion_type = abi.WPT_DISTORTION_RADIAL_ONLY
    c.b1, c.b2, c.b3, c.b4 = 0.12, -0.03, 0.008, -0.001


def opencv(view):
    c = view.camera
    c.distortion_type = abi.WPT_DISTORTION_OPENCV
    c.k1, c.k2, c.k3, c.p1, c.p2 = -0.28, 0.09, -0.012, 0.0015, -0.0011


def surround_180(view):
    view.camera.surround_mode = abi.WPT_SURROUND_180


def surround_180_stereo(view):
    view.camera.surround_mode = abi.WPT_SURROUND_180
    view.camera.stereoscopic_distance = 0.065


def gated(lo_frac, hi_frac, light_lo_frac, light_hi_frac):
    def f(view, extent):
        s = view.sensor
        s.min_path_len, s.max_path_len = lo_frac * extent, hi_frac * extent
        s.min_dist_to_light, s.max_dist_to_light = light_lo_frac * extent, light_hi_frac * extent
    f.needs_extent = True
    return f


# name, base scene, scale, modifier, (hit width, height), (render width, height, samples_sqrt)
VIEWS = [
    ("cornell_radial_planar", "cornell", 1.0, radial_and_planar, (96, 96), (64, 64, 12)),
    ("cornell_radial_only", "cornell", 1.0, radial_only, (96, 96), (64, 64, 12)),
    ("cornell_opencv", "cornell", 1.0, opencv, (96, 96), (64, 64, 12)),
    ("mixed_opencv", "mixed", 1.0, opencv, (96, 64), (96, 64, 10)),
    ("cornell_surround180", "cornell", 1.0, surround_180, (96, 96), (64, 64, 12)),
    ("mixed_surround180_stereo", "mixed", 1.0, surround_180_stereo, (96, 96), (64, 64, 10)),
    # gates in units of the scene's extent (distance of the farthest primary hit): paths between 1.2 and 2.6
    # extents long whose light is 0.15 .. 0.9 extents from the last vertex
    ("cornell_gated", "cornell", 1.0, gated(1.2, 2.6, 0.15, 0.9), (32, 32), (64, 64, 16)),
    ("mis_gated", "mis", 1.0, gated(0.0, 1.6, 0.0, 0.8), (32, 18), (96, 54, 12)),
]


def raw(struct):
    return np.frombuffer(bytes(struct), dtype=np.uint8).copy()


def main():
    only = sys.argv[1:]
    for name, base, scale, modify, (hw, hh), (rw, rh, spp) in VIEWS:
        if only and name not in only:
            continue
        view = refapi.View(base)
        scene = refapi.RefScene(base, scale, 0)
        scene.update_bvh(view.t0, view.t1)
        if getattr(modify, "needs_extent", False):
            _, d0 = scene.primary_hits(view.camera, 64, 64, view.t0, view.params.min_hit_distance)
            modify(view, float(d0.max()))
        else:
            modify(view)
        assert view.camera.n_keyframes == 0, "static cameras only"
        ids, dist = scene.primary_hits(view.camera, hw, hh, view.t0, view.params.min_hit_distance)
        view.sensor.width, view.sensor.height = rw, rh
        frame, seconds = scene.render(view.camera, view.sensor, view.params, spp, view.t0, view.t1)
        path = os.path.join(OUT, "view_" + name + ".npz")
        np.savez_compressed(path, base=np.array(base), camera=raw(view.camera), sensor=raw(view.sensor), params=raw(view.params),
                            t0=np.float32(view.t0), t1=np.float32(view.t1), hit_ids=ids, hit_dist=dist, render=frame,
                            render_spp_sqrt=np.int32(spp))
        print("view_%-26s %6.1f KiB  hit pixels %d/%d  render %dx%dx%d in %.2fs  mean %s" % (
            name, os.path.getsize(path) / 1024, int((ids >= 0).sum()), ids.size, rw, rh, spp * spp, seconds,
            np.round(frame.reshape(-1, 3).mean(axis=0), 5)))
        scene.close()


if __name__ == "__main__":
    main()
