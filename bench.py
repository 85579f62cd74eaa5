#!/usr/bin/env python
"""bench.py -- throughput of the path-tracing hot path (mcpt -> tracePath -> BVH::hit) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

A "step" is one mcpt() of the workload: the 1M-triangle synthetic scene (BASELINE.json config C3:
GGX materials, HDR environment map with importance sampling, 1920x1080, 256 spp). Rank r of N renders
the pixel tiles r, r+N, ... of the same image and the frames are combined with one NCCL reduce, so the
total work is fixed as N grows ("scaling": "strong").

Keys of the JSON line (see the task contract):
  value      Msamples/s, scene resident in HBM, frame left on the device, CUDA-event timed
  e2e        the same through the C-ABI call a user of the library makes (wpt_render with a HOST frame
             buffer): parameter upload + kernel + device->host copy of the frame inside the timed region
  roofline   the render kernel against measured HBM bandwidth, using the reference's algorithmic bytes
             per closest-hit query (32 B per node visit + 68 B per leaf test)
  cpu_baseline  the unmodified reference (oracle/_ref, compiled with its release flags) timed on this
             box's host cores on a bounded sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, scale, width, height, samples_sqrt)
    "synth": ("synth", 1.0, 1920, 1080, 16),
    "cornell": ("cornell", 1.0, 512, 512, 8),
    "rtiow": ("rtiow", 1.0, 1920, 1080, 16),
    # C4: omnidirectional stereo, 4.3 Gsamples per step (about 8 s on one B200)
    "ods": ("synth360", 1.0, 4096, 4096, 16),
    # C5: one ToF phase image of the animated scene at full quality; participating media
    "tof": ("tof", 1.0, 352, 288, 30),
    "media": ("media", 1.0, 512, 512, 16),
    # reduced copies for profiler runs (ncu replays every launch ~40 times)
    "synth_profile": ("synth", 1.0, 960, 540, 4),
    "synth_64spp": ("synth", 1.0, 1920, 1080, 8),
    "cornell_profile": ("cornell", 1.0, 512, 512, 4),
    "media_profile": ("media", 1.0, 512, 512, 4),
}


def load_json(path, default):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


class ClockSampler:
    """Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_reference_flavour():
    """The build of the reference this host's CPU runs fastest: the reference is compiled with -march=native
    (wurblpt-cornellbox/CMakeLists.txt:17); the binaries are built elsewhere, so the ISA level is chosen here."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((line for line in f if line.startswith("flags")), "")
    except OSError:
        flags = ""
    need = ["avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"]
    return "fast_v4" if all((" " + n) in flags for n in need) else "fast"


def reference_timing(workload, samples_sqrt, flavour=None):
    """mcpt() of the unmodified reference on the host cores: (Msamples/s, threads, seconds, sample text)."""
    from oracle import refapi
    scene_name, scale, width, height, _ = WORKLOADS[workload]
    if flavour is None:
        flavour = host_reference_flavour()
    if not refapi.available(flavour):
        flavour = "fast"
    if not refapi.available(flavour):
        flavour = "strict"
    view = refapi.View(scene_name, flavour=flavour)
    view.sensor.width, view.sensor.height = width, height
    scene = refapi.RefScene(scene_name, scale, 0, flavour=flavour)
    scene.update_bvh(view.t0, view.t1)
    _, seconds = scene.render(view.camera, view.sensor, view.params, samples_sqrt, view.t0, view.t1)
    samples = width * height * samples_sqrt * samples_sqrt
    threads = scene.lib.ref_threads()
    flags = scene.lib.ref_build_flags().decode()
    scene.close()
    text = "%s %dx%d at %d spp (%d samples) of the %d-spp workload, reference built with %s" % (
        scene_name, width, height, samples_sqrt ** 2, samples, WORKLOADS[workload][4] ** 2, flags)
    return samples / seconds / 1e6, threads, seconds, text


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation, rank 0 only, on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to its children; the reference's OpenMP loop must get every
    # core this process may run on (set before libgomp is loaded with the reference library)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    if os.environ.get("OMP_NUM_THREADS") not in (None, str(cores)) and not os.environ.get("WPT_BENCH_REEXEC"):
        os.environ["OMP_NUM_THREADS"] = str(cores)
        os.environ["WPT_BENCH_REEXEC"] = "1"
        sys.stdout.flush()
        os.execv(sys.executable, [sys.executable] + sys.argv)   # a clean process: no OpenMP runtime initialised yet
    os.environ["OMP_NUM_THREADS"] = str(cores)
    workload = args.workload
    spp = args.cpu_samples_sqrt
    values = []
    threads, text = 0, ""
    for i in range(args.warmup + args.steps):
        v, threads, seconds, text = reference_timing(workload, spp)
        if i >= args.warmup:
            values.append((v, seconds))
    value = sum(v for v, _ in values) / len(values)
    ms = 1e3 * sum(s for _, s in values) / len(values)
    scene_name, scale, width, height, full_spp = WORKLOADS[workload]
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s %dx%d %dspp" % (workload, width, height, full_spp ** 2),
                   "step": "one mcpt() of the reference on the host cores at %d spp (bounded sample)" % (spp * spp)},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "reference", "sample": text},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def measure_config(workload, device, stream, args):
    """Msamples/s and Mrays/s of one mcpt() of another config (frame resident, CUDA events; 1 warm-up + 2 timed)."""
    import torch
    from wurblpt_b200 import capi, scenes
    scene_name, scale, width, height, samples_sqrt = WORKLOADS[workload]
    built = scenes.BuiltScene(scene_name, scale)
    view = scenes.View(scene_name)
    view.sensor.width, view.sensor.height = width, height
    scene = capi.Scene(built.desc, device=device.index)
    scene.update_bvh(view.t0, view.t1)
    built.close()
    frame = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
    opts = capi.make_opts(frame_on_device=True, stream=stream.cuda_stream, traversal=args.traversal)
    ms, rays = [], 0
    for i in range(3):
        frame.zero_()
        st = scene.render_device(view.camera, view.sensor, view.params, samples_sqrt, frame.data_ptr(), view.t0, view.t1,
                                 opts, want_stats=True)
        if i > 0:
            ms.append(st.kernel_ms)
            rays = st.rays
    scene.close()
    samples = width * height * samples_sqrt * samples_sqrt
    t = sum(ms) / len(ms)
    return {"workload": "%s %dx%d %dspp" % (workload, width, height, samples_sqrt ** 2), "msamples_per_s": samples / t / 1e3,
            "mrays_per_s": rays / t / 1e3, "ms_per_step": t, "launches": int(st.launches)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="synth", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-samples-sqrt", type=int, default=4, help="samples_sqrt of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short measurements of the other configs")
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--substreams", type=int, default=0)
    ap.add_argument("--pipeline", default="auto", choices=["auto", "megakernel", "wavefront"])
    ap.add_argument("--work", default="dynamic", choices=["dynamic", "static"],
                    help="N > 1 ranks: 'dynamic' = every rank draws blocks of work items from one counter in rank 0's GPU "
                         "memory (CUDA IPC, NVLink); 'static' = rank r renders the tiles r, r+N, ...")
    ap.add_argument("--inprocess", action="store_true",
                    help="one process renders each frame on ALL visible GPUs (wpt_scene_create with WPT_DEVICE_ALL, "
                         "the path the reference's apps take through mcpt()); not for use under torchrun")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from wurblpt_b200 import abi, capi, multi, scenes

    if not torch.cuda.is_available() or capi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the library has no CPU fallback")
    rank, world, local = multi.init_process_group()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)

    scene_name, scale, width, height, samples_sqrt = WORKLOADS[args.workload]
    t = time.time()
    built = scenes.BuiltScene(scene_name, scale)
    view = scenes.View(scene_name)
    view.sensor.width, view.sensor.height = width, height
    build_scene_s = time.time() - t
    if args.inprocess and world > 1:
        raise SystemExit("--inprocess drives all GPUs from one process; do not start it under torchrun")
    scene = capi.Scene(built.desc, device=abi.WPT_DEVICE_ALL if args.inprocess else local)
    scene.update_bvh(view.t0, view.t1)
    devices_used = scene.device_count()
    built.close()
    info = scene.info()

    stream = torch.cuda.current_stream(device)
    frame = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
    host_frame = torch.zeros((height, width, 3), dtype=torch.float32).pin_memory()
    pipeline = {"auto": abi.WPT_PIPELINE_AUTO, "megakernel": abi.WPT_PIPELINE_MEGAKERNEL,
                "wavefront": abi.WPT_PIPELINE_WAVEFRONT}[args.pipeline]
    shared = None
    if world > 1 and args.work == "dynamic":
        shared = multi.SharedWork(rank, world, local)
        opts = capi.make_opts(traversal=args.traversal, substreams=args.substreams, frame_on_device=True,
                              stream=stream.cuda_stream, pipeline=pipeline)
    else:
        opts = multi.rank_opts(rank, world, substreams=args.substreams, stream=stream.cuda_stream, traversal=args.traversal,
                               pipeline=pipeline)
    # > 126 MB so that node / triangle lines cached in L2 by the previous step are gone
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step():
        """one mcpt(): clear the rank's frame, render its tiles, combine the frames on rank 0"""
        frame.zero_()
        if shared:
            opts.shared_counter = shared.begin_frame(stream.cuda_stream)
        scene.render_device(view.camera, view.sensor, view.params, samples_sqrt, frame.data_ptr(), view.t0, view.t1, opts)
        multi.reduce_frame(frame)

    for _ in range(args.warmup):
        l2_flush.fill_(1)
        step()
    barrier()

    # ---- timed region 1: device-resident throughput (the metric `value`) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ms = []
    barrier()
    for _ in range(args.steps):
        l2_flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        step_ms.append((e0, e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in step_ms)
    t_ms = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    ms_per_step = total_ms / args.steps
    samples_per_step = width * height * samples_sqrt * samples_sqrt
    value = samples_per_step / (ms_per_step * 1e3)

    # ---- kernel-only duration and ray count of this rank's render kernel (for the roofline) ----
    l2_flush.fill_(1)
    frame.zero_()
    torch.cuda.synchronize(device)
    if shared:
        opts.shared_counter = shared.begin_frame(stream.cuda_stream)
    st = scene.render_device(view.camera, view.sensor, view.params, samples_sqrt, frame.data_ptr(), view.t0, view.t1,
                             opts, want_stats=True)
    multi.reduce_frame(frame)       # (keeps the frames of the ranks in step: see multi.SharedWork)
    kernel_ms, rays_rank, samples_rank = st.kernel_ms, st.rays, st.samples
    launches_per_step = int(st.launches)
    c = torch.tensor([float(rays_rank), float(samples_rank)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    rays_total = float(c[0].item())

    # ---- timed region 2: end to end through the C-ABI call with a HOST frame buffer ----
    e2e = None
    if world == 1:
        e2e_opts = capi.make_opts(traversal=args.traversal, substreams=args.substreams, pipeline=pipeline)
        host_np = host_frame.numpy()
        for _ in range(2):
            scene.render(view.camera, view.sensor, view.params, samples_sqrt, view.t0, view.t1, e2e_opts, host_np)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            scene.render(view.camera, view.sensor, view.params, samples_sqrt, view.t0, view.t1, e2e_opts, host_np)
        e2e_s = (time.perf_counter() - t0) / args.steps
        h2d = C.sizeof(abi.wpt_camera_desc) + C.sizeof(abi.wpt_sensor_desc) + C.sizeof(abi.wpt_params) + C.sizeof(abi.wpt_render_opts)
        e2e = {"value": samples_per_step / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(host_np.nbytes), "checksum": float(host_np.mean())}
    else:
        # multi-GPU: the user-facing call is the same step plus the device->host read of the combined frame on rank 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
            if rank == 0:
                host_frame.copy_(frame, non_blocking=False)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        h2d = C.sizeof(abi.wpt_camera_desc) + C.sizeof(abi.wpt_sensor_desc) + C.sizeof(abi.wpt_params) + C.sizeof(abi.wpt_render_opts)
        e2e = {"value": samples_per_step / float(ts.item()) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(host_frame.numel() * 4), "checksum": float(host_frame.mean()) if rank == 0 else 0.0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the render kernel ----
    work = load_json(os.path.join(ROOT, "profiles", "r1_algorithmic_work.json"), {}).get(scene_name)
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), None)
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks else (6650.0, "fallback")
    roofline = None
    if work:
        algo_bytes = rays_rank * work["bytes_per_query"]
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        # DRAM bytes per closest-hit query from the committed ncu capture (all kernels of the step), scaled to
        # this launch's query count like `achieved`; None for workloads without a capture
        measured = load_json(os.path.join(ROOT, "profiles", "r2_issue_work.json"), {}).get(scene_name)
        traffic = None
        if measured and measured.get("dram_bytes_per_query") and launches_per_step > 2:
            traffic = rays_rank * measured["dram_bytes_per_query"]
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "traffic_source": "committed ncu capture (profiles/r2_issue_work.json: dram bytes of all launches of one frame / "
                                      "queries), scaled to this launch's query count; not measured in this run",
                    "peak_source": peak_src,
                    "kernel": "renderKernel (megakernel)" if launches_per_step <= 2 else
                    "one mcpt() = %d launches of waveTracePersistentKernel (dominant, ~65 %% of the time), waveSortKernel, "
                    "waveMaterialKernel, waveGenerateKernel; timed as a whole" % launches_per_step,
                    "kernel_ms": kernel_ms,
                    "queries_per_launch": rays_rank, "bytes_per_query": work["bytes_per_query"],
                    "nodes_per_query": work["nodes_per_query"], "leaves_per_query": work["leaves_per_query"],
                    "mrays_per_s": rays_rank / kernel_ms / 1e3}

    # ---- issue-slot roofline: the pipeline is bound by instruction issue, not by HBM ----
    issue = None
    iw = load_json(os.path.join(ROOT, "profiles", "r2_issue_work.json"), {}).get(scene_name)
    if iw and clocks and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        peak_issue = sms * 4 * clocks["sm_mhz"] * 1e6                      # warp instructions per second, 4 schedulers per SM
        achieved_issue = rays_rank * iw["warp_instr_per_query"] / (kernel_ms * 1e-3)
        issue = {"bound": "issue", "achieved": achieved_issue / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s",
                 "frac": achieved_issue / peak_issue, "warp_instr_per_query": iw["warp_instr_per_query"],
                 "lanes_per_instr": iw["lanes_per_instr"], "thread_slot_frac": achieved_issue / peak_issue * iw["lanes_per_instr"] / 32.0,
                 "source": iw.get("source")}

    # ---- frame check: bit pattern of the frame; on several GPUs it must equal the frame ONE GPU renders ----
    import zlib
    frame_check = None
    if rank == 0:
        combined = frame.cpu().numpy()
        frame_check = {"crc32": "%08x" % (zlib.crc32(combined.tobytes()) & 0xFFFFFFFF)}
        if world > 1:
            solo = torch.zeros_like(frame)
            solo_opts = multi.rank_opts(0, 1, substreams=args.substreams, stream=stream.cuda_stream, traversal=args.traversal,
                                        pipeline=pipeline)
            scene.render_device(view.camera, view.sensor, view.params, samples_sqrt, solo.data_ptr(), view.t0, view.t1, solo_opts)
            torch.cuda.synchronize(device)
            solo_np = solo.cpu().numpy()
            frame_check["single_gpu_crc32"] = "%08x" % (zlib.crc32(solo_np.tobytes()) & 0xFFFFFFFF)
            frame_check["equals_single_gpu"] = bool(np.array_equal(solo_np.view(np.uint32), combined.view(np.uint32)))

    # ---- the other configs BASELINE.json names, one short measurement each (N = 1 only) ----
    configs = None
    if world == 1 and not args.no_configs and not args.inprocess:
        configs = {}
        for w in ("cornell", "rtiow", "tof", "media"):
            if w == args.workload:
                continue
            try:
                configs[w] = measure_config(w, device, stream, args)
            except Exception as ex:
                configs[w] = {"error": str(ex)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            v, threads, seconds, text = reference_timing(args.workload, args.cpu_samples_sqrt)
            cpu = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "reference", "sample": text,
                   "seconds": seconds}
        except Exception as ex:     # the checker is optional for the product arm
            cpu = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ex}

    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": devices_used if args.inprocess else world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s %dx%d %dspp" % (args.workload, width, height, samples_sqrt ** 2),
                   "triangles": int(info.triangles), "hitables": int(info.hitables),
                   "wide_nodes": int(info.wide_nodes), "device_mb": info.device_bytes / 1e6,
                   "parallelism": ("one process, %d GPU(s): blocks of work items from one device-side counter, frames added "
                                   "device to device" % devices_used) if args.inprocess else
                   ("%d ranks draw blocks of 256 work items from one counter in rank 0's GPU memory (CUDA IPC over NVLink), "
                    "one NCCL reduce per frame" % world) if shared else
                   "tiles interleaved over %d GPU(s), one NCCL reduce per frame" % world,
                   "cache": "256 MiB buffer written between timed steps (L2 flush); scene 250 MB > 126 MB L2",
                   "traversal": "wide4" if args.traversal == 0 else "reference-order",
                   "pipeline": "megakernel" if launches_per_step <= 2 else "wavefront",
                   "scene_build_s": build_scene_s, "bvh_build_s": info.build_seconds, "upload_s": info.upload_seconds},
        "mrays_per_s": rays_total / (ms_per_step * 1e3), "rays_per_sample": rays_total / samples_per_step,
        "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks, "roofline": roofline,
        "issue_roofline": issue, "frame": frame_check, "configs": configs, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if frame_check and frame_check.get("equals_single_gpu") is False:
        sys.stderr.write("bench.py: the %d-GPU frame differs from the single-GPU frame\n" % world)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
