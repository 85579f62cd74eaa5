"""TEST INFRASTRUCTURE (uses oracle/). Regenerates profiles/r1_algorithmic_work.json: the reference's own node visits / leaf tests per BVH::hit()
(SURVEY.md section 8d), counted by the instrumented build of the unmodified reference (oracle/_ref/libwpt_ref_instr.so)
on a reduced image of each benchmark config. bytes/query = 32 n + 68 l, flops/query = 25 n + 50 l.
Run in the build container: make -C oracle ref && python tests/tools/algorithmic_work.py [config ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refapi  # noqa: E402

# config: (scale, width, height, samples_sqrt)
CONFIGS = {
    "synth": (1.0, 480, 270, 4),
    "cornell": (1.0, 256, 256, 8),
    "rtiow": (1.0, 480, 270, 8),
    "synth360": (1.0, 512, 512, 4),
    "tof": (1.0, 352, 288, 6),
    "media": (1.0, 256, 256, 8),
}


def main():
    path = os.path.join(ROOT, "profiles", "r1_algorithmic_work.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    for name in (sys.argv[1:] or CONFIGS):
        scale, w, h, spp = CONFIGS[name]
        view = refapi.View(name, flavour="instr")
        view.sensor.width, view.sensor.height = w, h
        scene = refapi.RefScene(name, scale, 0, flavour="instr")
        scene.update_bvh(view.t0, view.t1)
        scene.counters(reset=True)
        scene.render(view.camera, view.sensor, view.params, spp, view.t0, view.t1)
        c = scene.counters(reset=True)
        scene.close()
        # `rays` = closest-hit queries issued by tracePath; `calls` also counts nested medium-boundary queries
        q = c["rays"]
        n, l = c["nodes"] / q, c["leaves"] / q
        out[name] = {"width": w, "height": h, "samples_sqrt": spp, "queries": q,
                     "rays_per_sample": q / (w * h * spp * spp), "bvh_calls_per_query": c["calls"] / q,
                     "nodes_per_query": n, "leaves_per_query": l,
                     "bytes_per_query": 32 * n + 68 * l, "flops_per_query": 25 * n + 50 * l}
        print(name, json.dumps(out[name]))
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
