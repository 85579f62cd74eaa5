"""TEST INFRASTRUCTURE (uses oracle/): ad-hoc timing on the GPU box (development aid, not the bench contract): renders a config with
the scene taken from the oracle harness and prints Msamples/s, Mrays/s and the reference's CPU time."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refapi  # noqa: E402
from wurblpt_b200 import abi, capi  # noqa: E402
from util import box_downsample, rel_rmse  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="synth")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-spp", type=int, default=0, help="also time the reference with this samples_sqrt")
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--substreams", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--save", default="")
    ap.add_argument("--stride", type=int, default=1, help="render only every stride-th tile (the share of one of N ranks)")
    a = ap.parse_args()
    view = refapi.View(a.scene)
    if a.width:
        view.sensor.width, view.sensor.height = a.width, a.height
    spp = a.spp or view.samples_sqrt
    t = time.time()
    r = refapi.RefScene(a.scene, a.scale, 0)
    print("scene built in %.2fs" % (time.time() - t))
    s = capi.Scene(r.desc, device=0)
    s.update_bvh(view.t0, view.t1)
    i = s.info()
    print("hitables %d tris %d binary nodes %d depth %d | wide nodes %d depth %d | build %.2fs upload %.2fs device MB %.1f" % (
        i.hitables, i.triangles, i.binary_nodes, i.binary_depth, i.wide_nodes, i.wide_depth, i.build_seconds,
        i.upload_seconds, i.device_bytes / 1e6))
    opts = capi.make_opts(traversal=a.mode, substreams=a.substreams, tile_first=0, tile_stride=a.stride)
    frame = None
    for k in range(a.repeat):
        frame, st = s.render(view.camera, view.sensor, view.params, spp, view.t0, view.t1, opts)
        print("run %d: %dx%d spp %d: kernel %.1f ms total %.1f ms | %.1f Msamples/s %.1f Mrays/s | rays/sample %.2f nested %d" % (
            k, view.sensor.width, view.sensor.height, spp * spp, st.kernel_ms, st.total_ms,
            st.samples / st.kernel_ms / 1e3, st.rays / st.kernel_ms / 1e3, st.rays / max(1, st.samples), st.nested_queries))
    if a.save:
        np.save(a.save, frame)
    if a.cpu_spp:
        for flavour in ("fast", "strict"):
            if not refapi.available(flavour):
                continue
            rr = refapi.RefScene(a.scene, a.scale, 0, flavour=flavour)
            rr.update_bvh(view.t0, view.t1)
            f, sec = rr.render(view.camera, view.sensor, view.params, a.cpu_spp, view.t0, view.t1)
            n = view.sensor.width * view.sensor.height * a.cpu_spp ** 2
            print("reference (%s, %d threads): %d samples in %.2fs = %.3f Msamples/s" % (
                flavour, rr.lib.ref_threads(), n, sec, n / sec / 1e6))
            if flavour == "strict" or not refapi.available("strict"):
                m1, m2 = frame.reshape(-1, 3).mean(0), f.reshape(-1, 3).mean(0)
                print("  channel means gpu", m1, "ref", m2, "relRMSE(8x8)", rel_rmse(box_downsample(frame, 8), box_downsample(f, 8), floor=1e-2 * float(m2.mean()) ** 2))


if __name__ == "__main__":
    main()
