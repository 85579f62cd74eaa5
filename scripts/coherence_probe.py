"""How much does the order of incoherent rays matter to the closest-hit query? (GPU box)

Bounce rays leaving the surfaces seen by the synth camera are traced in three orders through
wpt_trace_rays (WPT_QUERY_TIMING=1 prints the kernel time to stderr): pixel order, a random permutation
(what the wavefront queues look like after a few path components), and sorted by (direction octant,
Morton code of the origin cell). The answers must be identical; only the time differs.
usage: WPT_QUERY_TIMING=1 python scripts/coherence_probe.py [width height]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wurblpt_b200 import abi, capi, scenes  # noqa: E402


def morton3(q):
    def spread(v):
        v = v.astype(np.uint64) & 0x3FF
        v = (v | (v << 16)) & 0x30000FF
        v = (v | (v << 8)) & 0x300F00F
        v = (v | (v << 4)) & 0x30C30C3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def main():
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
    built = scenes.BuiltScene("synth", 1.0)
    view = scenes.View("synth")
    scene = capi.Scene(built.desc, device=0).update_bvh(view.t0, view.t1)
    built.close()
    gt = scene.ground_truth(view.camera, w, h, view.t0, bits=3)      # world positions, world geometry normals
    pos = gt["world_positions"].reshape(-1, 3)
    nrm = gt["world_geometry_normals"].reshape(-1, 3)
    ok = np.linalg.norm(nrm, axis=1) > 0.5
    pos, nrm = pos[ok], nrm[ok]
    n = pos.shape[0]
    rng = np.random.default_rng(7)
    for bounce in (1, 2):
        # cosine-distributed directions about the normal
        u1, u2 = rng.random(n), rng.random(n)
        r, phi = np.sqrt(u1), 2 * np.pi * u2
        local = np.stack([r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(0, 1 - u1))], axis=1)
        a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
        t = np.cross(nrm, a)
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        b = np.cross(nrm, t)
        d = (local[:, :1] * t + local[:, 1:2] * b + local[:, 2:3] * nrm).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        o = (pos + 1e-4 * nrm).astype(np.float32)
        lo, hi = o.min(axis=0), o.max(axis=0)
        cell = np.clip(((o - lo) / (hi - lo + 1e-9) * 64).astype(np.int64), 0, 63)
        octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2)
        key = (octant << 18) | morton3(cell)
        orders = {"pixel order": np.arange(n), "random": rng.permutation(n), "sorted octant+cell": np.argsort(key, kind="stable"),
                  "sorted cell only": np.argsort(morton3(cell), kind="stable")}
        ref_ids = None
        for name, perm in orders.items():
            sys.stderr.write("bounce %d, %s: " % (bounce, name))
            sys.stderr.flush()
            ids, dist = scene.trace_rays(o[perm], d[perm], view.t0, 1e-5)
            back = np.empty_like(ids)
            back[perm] = ids
            if ref_ids is None:
                ref_ids, ref_perm_dist = back, dist
            else:
                assert (back == ref_ids).all(), "answers depend on the order"
        # next bounce starts where these rays ended (misses are dropped)
        ids, dist = scene.trace_rays(o, d, view.t0, 1e-5)
        hit = ids >= 0
        pos = o[hit] + dist[hit, None] * d[hit]
        # geometric normal unknown here: flip a random direction into the hemisphere facing back along the ray
        nrm = -d[hit]
        n = pos.shape[0]
    scene.close()


if __name__ == "__main__":
    main()
