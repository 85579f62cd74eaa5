"""GPU tests (-m gpu) of Scene::updateBVH() on the device (wurblpt_b200/csrc/bvh_device.cu, WPT_BUILDER_DEVICE_LBVH):
closest hits must be those of the reference (the tree only decides how fast they are found; exact distance ties are
counted), renders must agree, and the build must be fast."""
import time

import numpy as np
import pytest

from wurblpt_b200 import abi
from util import Golden, GOLDEN_SCENES, box_downsample, compare_hits, rel_rmse

pytestmark = pytest.mark.gpu

DEVICE = abi.WPT_BUILDER_DEVICE_LBVH


@pytest.fixture(scope="module")
def gpu(capi):
    if capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (there is no CPU fallback)")
    return capi


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_device_built_tree_finds_the_reference_hits(gpu, name):
    g = Golden(name)
    s = gpu.Scene(g.desc, device=0).set_builder(DEVICE).update_bvh(g.t0, g.t1)
    h, w = g.z["hit_ids"].shape
    ids, dist = s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance)
    c = compare_hits(ids, dist, g.z["hit_ids"], g.z["hit_dist"])
    print(name, "device-built tree vs reference:", c)
    same = (ids == g.z["hit_ids"]) & (ids >= 0)
    if g.camera.surround_mode != 0:      # sin/cos in the camera: rays differ by rounding (test_gpu_parity.py)
        assert c["wrong"] + c["near"] + c["ties"] <= c["total"] // 1000, c
        assert np.allclose(dist[same], g.z["hit_dist"][same], rtol=2e-5, atol=0.0)
    elif g.desc.n_media > 0:
        # media are served by the host builder: same tree as always
        assert c["wrong"] == 0, c
    else:
        assert c["wrong"] == 0, c
        assert c["ties"] + c["near"] <= max(4, c["total"] // 500), c
        assert np.array_equal(dist[same].view(np.uint32), g.z["hit_dist"][same].view(np.uint32))
    if g.desc.n_media == 0 and g.desc.n_components >= 2:
        with pytest.raises(gpu.WptError):       # no binary tree in the reference's layout: the parity walk is refused
            s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance, gpu.make_opts(traversal=abi.WPT_TRAVERSAL_REFERENCE))
    s.close()


@pytest.mark.parametrize("name", ["cornell", "rtiow", "mixed", "tof", "synth"])
def test_device_and_host_built_trees_answer_alike(gpu, name):
    """200 000 incoherent rays: ids and distance bits equal to the host-built tree's (ties excepted and counted),
    through the per-lane walk and through the pooled trace kernel."""
    g = Golden(name)
    host = gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)
    dev = gpu.Scene(g.desc, device=0).set_builder(DEVICE).update_bvh(g.t0, g.t1)
    boxes, _, _ = host.binary_nodes()
    lo, hi = np.maximum(boxes[0, :3], -25.0), np.minimum(boxes[0, 3:], 25.0)
    rng = np.random.default_rng(4321)
    n = 200000
    o = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d = (d / np.sqrt((d ** 2).sum(axis=1, keepdims=True))).astype(np.float32)
    hid, hdist = host.trace_rays(o, d, g.t0, 1e-4)
    modes = [abi.WPT_TRAVERSAL_WIDE] + ([abi.WPT_TRAVERSAL_POOL, abi.WPT_TRAVERSAL_PERSISTENT, abi.WPT_TRAVERSAL_QUAD] if g.desc.n_animations == 0 else [])
    for mode in modes:
        did, ddist = dev.trace_rays(o, d, g.t0, 1e-4, gpu.make_opts(traversal=mode))
        c = compare_hits(did, ddist, hid, hdist)
        # Where two surfaces coincide (a box standing on the floor) the reference's storage order decides which one a
        # ray "hits"; the host-built tree reproduces that order, Morton order cannot. Such hits have the same distance
        # up to the rounding of two different triangle tests: allowed within 2e-5 relative + 1e-6 absolute, and counted.
        diff = did != hid
        both = diff & (did >= 0) & (hid >= 0)
        coincident = both & (np.abs(ddist - hdist) <= 2e-5 * np.abs(hdist) + 1e-6)
        print(name, mode, c, "coincident surfaces:", int(coincident.sum()))
        odd = diff & ~coincident
        for k in np.flatnonzero(odd)[:8]:
            print("  odd ray %d: device tree id %d a %.9g, host tree id %d a %.9g, o %s d %s" % (
                k, did[k], ddist[k], hid[k], hdist[k], o[k], d[k]))
        assert int(odd.sum()) == 0, (name, mode, c)
        assert int(diff.sum()) <= n // 500, (name, mode, c)
        same = (did == hid) & (hid >= 0)
        assert np.array_equal(ddist[same].view(np.uint32), hdist[same].view(np.uint32))
    host.close()
    dev.close()


@pytest.mark.parametrize("name", ["cornell", "rtiow", "mis", "tof"])
def test_render_with_device_built_tree(gpu, name):
    """Same closest hits => same paths: the frame of the device-built tree equals the host-built tree's except
    where an exact tie was settled differently."""
    g = Golden(name)
    sensor = g.sensor
    sensor.height, sensor.width = g.z["render"].shape[:2]
    spp = int(g.z["render_spp_sqrt"])
    frames = []
    for builder in (abi.WPT_BUILDER_HOST_SAH, DEVICE):
        s = gpu.Scene(g.desc, device=0).set_builder(builder).update_bvh(g.t0, g.t1)
        f, st = s.render(g.camera, sensor, g.params, spp, g.t0, g.t1, gpu.make_opts(substreams=1))
        frames.append(f)
        s.close()
    differing = (frames[0].view(np.uint32) != frames[1].view(np.uint32)).any(axis=2).mean()
    print(name, "pixels that differ between the builders: %.4f" % differing)
    assert differing <= 0.02, (name, differing)
    assert abs(frames[0].mean() - frames[1].mean()) <= 0.005 * frames[0].mean()
    assert rel_rmse(box_downsample(frames[1], 8), box_downsample(frames[0], 8), floor=1e-3) <= 0.02


def test_device_build_of_one_million_triangles(gpu, capi):
    """C3's scene: build time, and what the tree costs the walk (wide nodes visited per primary ray) next to the
    host-built SAH tree."""
    from wurblpt_b200 import scenes
    built = scenes.BuiltScene("synth", 1.0)
    view = scenes.View("synth")
    host = capi.Scene(built.desc, device=0).update_bvh(view.t0, view.t1)
    dev = capi.Scene(built.desc, device=0).set_builder(DEVICE)
    dev.update_bvh(view.t0, view.t1)        # first call: allocations
    t = time.perf_counter()
    dev.update_bvh(view.t0, view.t1)
    rebuild_s = time.perf_counter() - t
    built.close()
    hi, di = host.info(), dev.info()
    assert di.triangles == hi.triangles == 1008120
    w, h = 960, 540
    hid, hdist, hc = host.primary_hits(view.camera, w, h, view.t0, 1e-5, count=True)
    did, ddist, dc = dev.primary_hits(view.camera, w, h, view.t0, 1e-5, count=True)
    c = compare_hits(did, ddist, hid, hdist)
    print("1M triangles: host build %.3f s, device rebuild %.4f s (of which host bounds + upload); wide nodes %d vs %d, depth %d vs %d; "
          "nodes/ray %.2f vs %.2f, leaves/ray %.2f vs %.2f; %s" % (
              hi.build_seconds, rebuild_s, hi.wide_nodes, di.wide_nodes, hi.wide_depth, di.wide_depth,
              hc["nodes"] / (w * h), dc["nodes"] / (w * h), hc["leaves"] / (w * h), dc["leaves"] / (w * h), c))
    assert c["wrong"] == 0 and c["ties"] + c["near"] <= 50, c
    assert rebuild_s < 0.15, rebuild_s
    assert dc["nodes"] <= 2.0 * hc["nodes"]
    host.close()
    dev.close()
