"""Node visits and leaf tests per query of the wide walk on the 1M-triangle scene (GPU box), for A/B of box-test variants.
usage: [WPT_LIB=...] WPT_QUERY_TIMING=1 python scripts/slab_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wurblpt_b200 import capi, scenes  # noqa: E402

built = scenes.BuiltScene("synth", 1.0)
view = scenes.View("synth")
scene = capi.Scene(built.desc, device=0).update_bvh(view.t0, view.t1)
built.close()
ids, a, c = scene.primary_hits(view.camera, 960, 540, view.t0, 1e-4, count=True)
n = ids.size
print("primary: nodes/ray %.2f leaves/ray %.2f hit %.3f" % (c["nodes"] / n, c["leaves"] / n, float((ids >= 0).mean())))
rng = np.random.default_rng(5)
m = 400000
o = np.stack([rng.uniform(-6, 6, m), rng.uniform(0.05, 4, m), rng.uniform(-10, 10, m)], axis=1).astype(np.float32)
d = rng.normal(size=(m, 3)).astype(np.float32)
d = (d / np.sqrt((d ** 2).sum(axis=1, keepdims=True))).astype(np.float32)
ids2, a2, c2 = scene.trace_rays(o, d, view.t0, 1e-4, count=True)
print("random: nodes/ray %.2f leaves/ray %.2f hit %.3f" % (c2["nodes"] / m, c2["leaves"] / m, float((ids2 >= 0).mean())))
# per-ray cost distribution: trace in chunks sorted by |smallest direction component|
k = np.abs(d).min(axis=1)
order = np.argsort(k)
for lo, hi in ((0, 400), (400, 4000), (4000, 40000), (40000, m)):
    sel = order[lo:hi]
    _, _, cc = scene.trace_rays(o[sel], d[sel], view.t0, 1e-4, count=True)
    print("  min|d| in [%.2e, %.2e]: nodes/ray %.1f leaves/ray %.1f" % (k[sel].min(), k[sel].max(), cc["nodes"] / len(sel), cc["leaves"] / len(sel)))
np.save(os.path.join(ROOT, "gpurun_out", "probe_ids_%s.npy" % (os.path.basename(os.environ.get("WPT_LIB", "default")))), ids2)
