"""Regenerates tests/golden/app_*.npz: box-filtered frames of the reference's example apps rendered by the
REFERENCE itself (oracle/_ref/apps/*, built by `make -C oracle refapps` from the unchanged app sources against the
reference's headers). The GPU tests run the same unchanged apps against this repository's API mirror and compare.

    make -C oracle refapps && python tests/golden/make_app_golden.py [toomuch] [noise] [playground]
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import box_downsample, read_png, read_wptarray  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
APPS = os.path.join(ROOT, "oracle", "_ref", "apps")


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = [int(v) for v in f.readline().split()]
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(h, w, 3)


def toomuch():
    with tempfile.TemporaryDirectory() as d:
        t = time.time()
        subprocess.run([os.path.join(APPS, "wurblpt-toomuch"), "0"], cwd=d, check=True, capture_output=True, timeout=1800)
        hdr = read_wptarray(os.path.join(d, "toomuch-0000.tgd"))
        small = box_downsample(hdr, 4).astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "app_toomuch.npz"), frame0=small, box=np.int32(4))
        print("toomuch frame 0: %s -> %s, mean %s, %.1f s" % (hdr.shape, small.shape, hdr.reshape(-1, 3).mean(axis=0), time.time() - t))


def noise():
    with tempfile.TemporaryDirectory() as d:
        t = time.time()
        subprocess.run([os.path.join(APPS, "wurblpt-noise-textures")], cwd=d, check=True, capture_output=True, timeout=3600)
        out = {}
        for k in range(10):
            img = read_png(os.path.join(d, "noise-example-%d.png" % k))
            out["t%d" % k] = box_downsample(img.astype(np.float32), 8).astype(np.float16)
        np.savez_compressed(os.path.join(OUT, "app_noise_textures.npz"), box=np.int32(8), **out)
        print("noise-textures: 10 frames %s -> %s, %.1f s" % (img.shape, out["t0"].shape, time.time() - t))


def playground():
    with tempfile.TemporaryDirectory() as d:
        t = time.time()
        p = subprocess.Popen([os.path.join(APPS, "wurblpt-material-playground")], cwd=d, stdout=subprocess.DEVNULL,
                             stderr=subprocess.DEVNULL)
        try:
            while not os.path.exists(os.path.join(d, "frame-0003.pnm")):     # frames 0..2 are complete once 3 exists
                time.sleep(1.0)
                assert p.poll() is None and time.time() - t < 3600
        finally:
            p.kill()
        out = {}
        for k in range(3):
            img = read_ppm(os.path.join(d, "frame-%04d.pnm" % k))
            out["frame%d" % k] = box_downsample(img.astype(np.float32), 8).astype(np.float16)
        np.savez_compressed(os.path.join(OUT, "app_material_playground.npz"), box=np.int32(8), **out)
        print("material-playground: frames 0..2 %s -> %s, %.1f s" % (img.shape, out["frame0"].shape, time.time() - t))


if __name__ == "__main__":
    which = sys.argv[1:] or ["toomuch", "noise", "playground"]
    for w in which:
        {"toomuch": toomuch, "noise": noise, "playground": playground}[w]()
