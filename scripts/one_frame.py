"""One mcpt() of a bench workload and nothing else (for profiler passes): prints the frame's statistics as JSON.
usage: python scripts/one_frame.py [workload] (see bench.py WORKLOADS)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from wurblpt_b200 import capi, scenes  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "synth_profile"
scene_name, scale, width, height, samples_sqrt = bench.WORKLOADS[workload]
built = scenes.BuiltScene(scene_name, scale)
view = scenes.View(scene_name)
view.sensor.width, view.sensor.height = width, height
scene = capi.Scene(built.desc, device=0).update_bvh(view.t0, view.t1)
built.close()
frame, st = scene.render(view.camera, view.sensor, view.params, samples_sqrt, view.t0, view.t1)
print(json.dumps({"workload": workload, "scene": scene_name, "samples": st.samples, "rays": st.rays, "launches": st.launches,
                  "kernel_ms": st.kernel_ms, "mean": float(frame.mean())}))
