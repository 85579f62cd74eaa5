"""Regenerates tests/golden/*.npz from the compiled, unmodified reference (oracle/_ref).

Run in the build container (needs /root/reference at oracle build time):
    make -C oracle ref && python tests/golden/make_golden.py

Per scene the file holds
  - the flattened scene description (wurblpt_b200.scene_desc arrays),
  - the default view (camera / sensor / params as raw struct bytes),
  - the reference's linearised BVH (boxes, link, leaf),
  - primary-ray closest hits at a small resolution (ids, distances),
  - a low-resolution reference render (mcpt of the reference, strict build).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refapi  # noqa: E402
from wurblpt_b200 import scene_desc as sd  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name, scale, (hit width, hit height), (render width, height, samples_sqrt)
SCENES = [
    ("cornell", 1.0, (96, 96), (64, 64, 16)),
    ("rtiow", 1.0, (128, 72), (96, 54, 12)),
    ("mis", 1.0, (96, 54), (96, 54, 12)),
    ("mis_nohot", 1.0, (96, 54), (96, 54, 12)),
    ("media", 1.0, (96, 96), (48, 48, 16)),
    ("tof", 1.0, (88, 72), (88, 72, 10)),
    ("furnace", 1.0, (64, 64), (64, 64, 10)),
    ("mixed", 1.0, (96, 64), (96, 64, 12)),
    ("synth", 0.004, (128, 72), (96, 54, 8)),
    ("synth360", 0.004, (128, 128), (64, 64, 8)),     # C4: 360 degree top-bottom stereo camera
    ("imported", 1.0, (96, 72), (80, 60, 12)),        # OBJ/MTL import: the description is the REFERENCE importer's
]


# getGroundTruth() fixtures: name, scale, (width, height), (t0, tPrev, tNext)
GROUND_TRUTH = [
    ("tof", 1.0, (44, 36), (1.0, 0.96, 1.04)),       # animated hitables: the flow images are non-zero
    ("mixed", 1.0, (48, 32), (0.0, 0.0, 0.0)),       # spheres, textures, two-sided materials
]


def raw(struct):
    return np.frombuffer(bytes(struct), dtype=np.uint8).copy()


def main():
    if not sys.argv[1:]:
        write_prng_kat()
    write_scenes()
    write_ground_truth()


def write_ground_truth():
    only = sys.argv[1:]
    for name, scale, (w, h), (t0, tp, tn) in GROUND_TRUTH:
        if only and ("gt_" + name) not in only:
            continue
        view = refapi.View(name)
        scene = refapi.RefScene(name, scale, 0)
        scene.update_bvh(t0, t0)
        images = scene.ground_truth(view.camera, w, h, t0, tp, tn, view.params.min_hit_distance)
        path = os.path.join(OUT, "gt_" + name + ".npz")
        np.savez_compressed(path, times=np.array([t0, tp, tn], dtype=np.float32), scale=np.float32(scale), **images)
        print("gt_%-7s %7.1f KiB  %dx%d, %d pixels hit" % (name, os.path.getsize(path) / 1024, w, h,
                                                         int((images["materials"] >= 0).sum())))
        scene.close()


def write_prng_kat():
    kat_seeds = np.array([0, 1, 42, 4095, 262143, 2073599, 0xFFFFFFFF], dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, "prng_kat.npz"), seeds=kat_seeds, values=refapi.prng_kat(kat_seeds, 16))


def write_scenes():
    only = sys.argv[1:]        # optional: regenerate just these scenes
    for name, scale, (hw, hh), (rw, rh, spp) in SCENES:
        if only and name not in only:
            continue
        view = refapi.View(name)
        scene = refapi.RefScene(name, scale, 0)
        scene.update_bvh(view.t0, view.t1)
        arrays = sd.desc_to_arrays(scene.desc)
        boxes, link, leaf = refapi.bvh_nodes(scene)
        ids, dist = scene.primary_hits(view.camera, hw, hh, view.t0, view.params.min_hit_distance)
        view.sensor.width, view.sensor.height = rw, rh
        frame, seconds = scene.render(view.camera, view.sensor, view.params, spp, view.t0, view.t1)
        cam = view.camera
        cam_keys = np.zeros(0, dtype=np.uint8)
        if cam.n_keyframes:
            cam_keys = np.frombuffer(C.string_at(cam.keyframes, C.sizeof(cam.keyframes._type_) * cam.n_keyframes),
                                     dtype=np.uint8).copy()
        out = {"scene_" + k: v for k, v in arrays.items()}
        out.update(camera=raw(cam), camera_keyframes=cam_keys, sensor=raw(view.sensor), params=raw(view.params),
                   t0=np.float32(view.t0), t1=np.float32(view.t1), scale=np.float32(scale),
                   bvh_boxes=boxes, bvh_link=link, bvh_leaf=leaf,
                   hit_ids=ids, hit_dist=dist, render=frame, render_spp_sqrt=np.int32(spp))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-10s %7.1f KiB  hitables=%d  hit pixels=%d/%d  render %dx%dx%d in %.2fs" % (
            name, os.path.getsize(path) / 1024, len(leaf) // 2 + 1, int((ids >= 0).sum()), ids.size, rw, rh, spp * spp,
            seconds))
        scene.close()


if __name__ == "__main__":
    main()
