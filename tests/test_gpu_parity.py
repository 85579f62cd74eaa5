"""GPU parity tests (-m gpu): the sm_100a kernels, called through the C ABI, against the oracle.

Integer / index results (closest-hit ids) must be exact apart from counted exact-distance ties;
hit distances are compared bit for bit; radiance images are Monte-Carlo estimates and are compared
statistically with the tolerance written next to each check.
"""
import numpy as np
import pytest

from wurblpt_b200 import abi
from util import Golden, GOLDEN_SCENES, GOLDEN, box_downsample, compare_hits, read_wptarray, rel_rmse

pytestmark = pytest.mark.gpu

REFERENCE = abi.WPT_TRAVERSAL_REFERENCE
WIDE = abi.WPT_TRAVERSAL_WIDE
POOL = abi.WPT_TRAVERSAL_POOL
PERSISTENT = abi.WPT_TRAVERSAL_PERSISTENT
QUAD = abi.WPT_TRAVERSAL_QUAD


@pytest.fixture(scope="module")
def gpu(capi):
    if capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (there is no CPU fallback)")
    return capi


def device_scene(gpu, g):
    return gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)


def test_prng_known_answers(gpu):
    with np.load(GOLDEN + "/prng_kat.npz") as z:
        out = gpu.prng_kat(z["seeds"], z["values"].shape[1])
        assert np.array_equal(out.view(np.uint32), z["values"].view(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_primary_hits_reference_walk_is_bit_exact(gpu, name):
    """Same tree, same walk order, same arithmetic: ids AND distances must equal the reference's."""
    g = Golden(name)
    s = device_scene(gpu, g)
    h, w = g.z["hit_ids"].shape
    ids, dist = s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance, gpu.make_opts(traversal=REFERENCE))
    c = compare_hits(ids, dist, g.z["hit_ids"], g.z["hit_dist"])
    hit = g.z["hit_ids"] >= 0
    if g.camera.surround_mode != 0:
        # 180/360 degree cameras build the ray direction with sin/cos (camera.hpp:160-169); CUDA's sinf/cosf
        # and glibc's differ by up to 2 ulp, so the rays are not bit-identical: same closest hits wherever the
        # ray is not within rounding of a silhouette (allowed: 0.1 % of the rays), distances within 2e-5 relative
        assert c["wrong"] + c["near"] + c["ties"] <= c["total"] // 1000, c
        same = hit & (ids == g.z["hit_ids"])
        assert np.allclose(dist[same], g.z["hit_dist"][same], rtol=2e-5, atol=0.0)
        return
    assert c["wrong"] == 0 and c["near"] == 0 and c["ties"] == 0, c
    assert np.array_equal(dist[hit].view(np.uint32), g.z["hit_dist"][hit].view(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_primary_hits_wide_walk(gpu, name):
    """The ordered 4-wide walk: identical closest hits; exact ties are counted, never wrong."""
    g = Golden(name)
    s = device_scene(gpu, g)
    h, w = g.z["hit_ids"].shape
    ids, dist = s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance, gpu.make_opts(traversal=WIDE))
    c = compare_hits(ids, dist, g.z["hit_ids"], g.z["hit_dist"])
    print(name, "wide walk vs reference:", c)
    same = ids == g.z["hit_ids"]
    hit = same & (ids >= 0)
    if g.camera.surround_mode != 0:      # sin/cos in the camera, see the reference-walk test
        assert c["wrong"] + c["near"] + c["ties"] <= c["total"] // 1000, c
        assert np.allclose(dist[hit], g.z["hit_dist"][hit], rtol=2e-5, atol=0.0)
        return
    assert c["wrong"] == 0, c
    assert c["ties"] + c["near"] <= max(4, c["total"] // 500), c
    assert np.array_equal(dist[hit].view(np.uint32), g.z["hit_dist"][hit].view(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_SCENES)
@pytest.mark.parametrize("mode", [PERSISTENT, POOL, QUAD])
def test_pipeline_trace_kernels_answer_like_the_wide_walk(gpu, name, mode):
    """The trace kernels of the wavefront pipeline - waveTracePersistentKernel (lanes refill inside the walk loop, leaf
    rounds, triangles of a leaf round pooled and tested one per lane), waveTracePoolKernel (walk state in shared
    memory, node and leaf passes from queues) and waveTraceQuadKernel (four lanes per query: one child of a wide node /
    one triangle of a leaf each) - run the same tests in the same order per query as the per-lane wide
    walk of the query kernel: ids and distance bits must be identical, ties included."""
    g = Golden(name)
    s = device_scene(gpu, g)
    h, w = g.z["hit_ids"].shape
    try:
        ids, dist = s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance, gpu.make_opts(traversal=mode))
    except gpu.WptError as e:
        assert e.code == abi.WPT_E_UNSUPPORTED      # media / animated hitables: kernel variants with Prng and animation cache
        assert g.desc.n_media > 0 or g.desc.n_animations > 0, name
        return
    wids, wdist = s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance, gpu.make_opts(traversal=WIDE))
    assert np.array_equal(ids, wids)
    assert np.array_equal(dist.view(np.uint32), wdist.view(np.uint32))
    # and ragged counts of incoherent rays (refills, drain, partial rounds)
    rng = np.random.default_rng(99)
    for n in (1, 31, 33, 4097, 100003):
        boxes, _, _ = s.binary_nodes()
        lo, hi = np.maximum(boxes[0, :3], -25.0), np.minimum(boxes[0, 3:], 25.0)
        o = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d = (d / np.sqrt((d ** 2).sum(axis=1, keepdims=True))).astype(np.float32)
        a_ids, a_dist = s.trace_rays(o, d, g.t0, 1e-4, gpu.make_opts(traversal=mode))
        b_ids, b_dist = s.trace_rays(o, d, g.t0, 1e-4, gpu.make_opts(traversal=WIDE))
        assert np.array_equal(a_ids, b_ids), (name, n)
        assert np.array_equal(a_dist.view(np.uint32), b_dist.view(np.uint32)), (name, n)


@pytest.mark.parametrize("name,scale", [("cornell", 1.0), ("rtiow", 1.0), ("mixed", 1.0), ("media", 1.0), ("synth", 0.05)])
def test_secondary_ray_closest_hits_against_live_reference(gpu, ref, name, scale):
    """BVH::hit() parity on an incoherent ray set: random origins inside the scene box, random directions."""
    r = ref.RefScene(name, scale, 0)
    r.update_bvh()
    s = gpu.Scene(r.desc, device=0).update_bvh()
    boxes, _, _ = s.binary_nodes()
    lo, hi = boxes[0, :3], boxes[0, 3:]
    span = np.minimum(hi - lo, 50.0)
    mid = 0.5 * (np.maximum(lo, -25.0) + np.minimum(hi, 25.0))
    rng = np.random.default_rng(1234)
    n = 200000
    o = (mid + (rng.random((n, 3)) - 0.5) * span * 0.9).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # re-normalise in float32 the way a Ray must be (ray.hpp:47)
    d = (d / np.sqrt((d.astype(np.float32) ** 2).sum(axis=1, keepdims=True))).astype(np.float32)
    has_media = r.desc.n_media > 0
    rid, ra = r.trace_rays(o, d, 0.0, 1e-4)
    for mode in (REFERENCE, WIDE) + (() if has_media else (POOL, PERSISTENT, QUAD)):
        ids, dist = s.trace_rays(o, d, 0.0, 1e-4, gpu.make_opts(traversal=mode))
        if has_media:
            # media hits are random free-flight samples (hitable_medium.hpp:89): compare surfaces only
            med = np.frombuffer(bytes(r.desc.components[0:0]), dtype=np.uint8)  # no-op, keeps flake8 quiet
            surf = np.ones(n, dtype=bool)
            n_prims = 0
            comps = r.desc.components
            kinds = []
            for ci in range(r.desc.n_components):
                c = comps[ci]
                cnt = r.desc.meshes[r.desc.instances[c.ref].mesh].triangle_count if c.type == 0 else 1
                kinds += [c.type] * cnt
            kinds = np.array(kinds)
            is_med = np.zeros(n, dtype=bool)
            is_med[rid >= 0] |= kinds[rid[rid >= 0]] == 2
            is_med[ids >= 0] |= kinds[ids[ids >= 0]] == 2
            keep = ~is_med
            c = compare_hits(ids[keep], dist[keep], rid[keep], ra[keep])
            # a ray may be stopped in a medium by one implementation and pass through in the other;
            # those were removed, the rest must agree
            assert c["wrong"] <= 0.002 * c["total"], (mode, c)
            continue
        c = compare_hits(ids, dist, rid, ra)
        assert c["wrong"] == 0, (mode, c)
        if mode == REFERENCE:
            assert c["ties"] == 0 and c["near"] == 0, c
        else:
            assert c["ties"] + c["near"] <= max(2, n // 5000), c


def test_reference_walk_counters_match_instrumented_reference(gpu):
    """Node visits / leaf tests of the reference-order walk are the reference's own:
    cornell primary rays visit 71-node tree deterministically; just check the counters are sane and
    that the wide walk visits fewer nodes."""
    g = Golden("cornell")
    s = device_scene(gpu, g)
    _, _, cref = s.primary_hits(g.camera, 256, 256, 0.0, 1e-5, gpu.make_opts(traversal=REFERENCE), count=True)
    _, _, cwide = s.primary_hits(g.camera, 256, 256, 0.0, 1e-5, gpu.make_opts(traversal=WIDE), count=True)
    n = 256 * 256
    assert 5 < cref["nodes"] / n < 71
    assert 0.5 < cref["leaves"] / n < 36
    assert cwide["nodes"] < cref["nodes"]
    # wide-tree leaves hold up to four triangles, so it may test a few more primitives than the reference
    assert cwide["leaves"] <= 4 * cref["leaves"]


# ---------------------------------------------------------------- radiance

def render_golden_view(gpu, g, opts=None, spp_sqrt=None):
    s = device_scene(gpu, g)
    sensor = g.sensor
    sensor.height, sensor.width = g.z["render"].shape[:2]
    spp = int(g.z["render_spp_sqrt"]) if spp_sqrt is None else spp_sqrt
    frame, stats = s.render(g.camera, sensor, g.params, spp, g.t0, g.t1, opts)
    return frame, stats


# tolerance: both images are Monte-Carlo estimates at the golden's (low) sample count. After an
# 8x8 box filter the per-pixel noise drops by ~8x; the bound below is ~4 sigma of what two
# independent reference renders with different seeds differ by (measured here: 2-4 %).
RENDER_TOL = {"cornell": 0.06, "rtiow": 0.06, "mis": 0.10, "mis_nohot": 0.25, "media": 0.10, "tof": 0.05,
              "furnace": 0.01, "mixed": 0.12, "synth": 0.25, "synth360": 0.25,
              "imported": 0.02}    # (the OBJ/MTL scene: two reference renders of it differ by 0.1 %)


@pytest.mark.parametrize("name", GOLDEN_SCENES)
@pytest.mark.parametrize("mode", [WIDE, REFERENCE])
def test_render_matches_reference_render(gpu, name, mode):
    g = Golden(name)
    frame, stats = render_golden_view(gpu, g, gpu.make_opts(traversal=mode, substreams=1))
    ref_img = g.z["render"]
    assert np.isfinite(frame).all()
    assert stats.samples == frame.shape[0] * frame.shape[1] * int(g.z["render_spp_sqrt"]) ** 2
    assert stats.rays >= stats.samples
    # global energy: channel means within 3 % (5 % for the high-variance scenes)
    mean_tol = 0.05 if name in ("mis_nohot", "synth", "synth360", "mixed") else 0.03
    m_dev = frame.reshape(-1, 3).mean(axis=0)
    m_ref = ref_img.reshape(-1, 3).mean(axis=0)
    assert np.all(np.abs(m_dev - m_ref) <= mean_tol * np.maximum(m_ref, 1e-3)), (m_dev, m_ref)
    # structure: relative RMSE of 8x8 box-filtered images
    e = rel_rmse(box_downsample(frame, 8), box_downsample(ref_img, 8), floor=1e-2 * float(m_ref.mean()) ** 2 + 1e-12)
    assert e <= RENDER_TOL[name], (name, e)


def test_furnace_is_analytic(gpu):
    """wurblpt-furnace-test: Lambertian sphere (albedo 0.42) in a constant L=1 environment."""
    g = Golden("furnace")
    frame, _ = render_golden_view(gpu, g, gpu.make_opts(substreams=1))
    ids = g.z["hit_ids"]
    assert frame.shape[:2] == ids.shape
    bg = frame[ids < 0]
    assert np.all(bg == 1.0)
    sphere = frame[ids >= 0]
    interior = sphere[np.abs(sphere[:, 0] - 0.42) < 2e-3]
    assert len(interior) > 0.9 * len(sphere)
    assert np.abs(interior - 0.42).max() < 1e-4
    assert sphere.min() > 0.39          # silhouette pixels lose a few grazing samples


def test_mis_and_material_sampling_agree_in_expectation(gpu):
    """wurblpt-mis-test: same scene with and without hot spots converges to the same image."""
    a = Golden("mis")
    b = Golden("mis_nohot")
    fa, _ = render_golden_view(gpu, a, spp_sqrt=24)
    fb, _ = render_golden_view(gpu, b, spp_sqrt=24)
    ma, mb = fa.mean(), fb.mean()
    assert abs(ma - mb) <= 0.04 * mb, (ma, mb)
    # MIS has the lower variance: its distance to the blurred consensus is smaller
    cons = box_downsample(0.5 * (fa + fb), 6)
    assert rel_rmse(box_downsample(fa, 6), cons) <= rel_rmse(box_downsample(fb, 6), cons) + 0.02


def test_render_is_deterministic(gpu):
    g = Golden("cornell")
    f1, _ = render_golden_view(gpu, g, gpu.make_opts(substreams=4))
    f2, _ = render_golden_view(gpu, g, gpu.make_opts(substreams=4))
    assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32))


def test_substreams_do_not_change_the_estimate(gpu):
    g = Golden("cornell")
    f1, _ = render_golden_view(gpu, g, gpu.make_opts(substreams=1))
    f8, s8 = render_golden_view(gpu, g, gpu.make_opts(substreams=8))
    assert s8.samples == f8.shape[0] * f8.shape[1] * 256
    assert abs(f1.mean() - f8.mean()) <= 0.02 * f1.mean()
    assert rel_rmse(box_downsample(f8, 8), box_downsample(f1, 8), floor=1e-3) <= 0.06


def test_tile_partition_is_an_exact_split(gpu):
    """Rendering tiles {0,2,4,..} and {1,3,5,..} separately and adding the frames gives the single
    render bit for bit (each pixel is written by exactly one part; x + 0 = x)."""
    g = Golden("cornell")
    whole, sw = render_golden_view(gpu, g, gpu.make_opts(substreams=2))
    parts = [render_golden_view(gpu, g, gpu.make_opts(substreams=2, tile_first=k, tile_stride=3)) for k in range(3)]
    total = parts[0][0] + parts[1][0] + parts[2][0]
    assert np.array_equal(total.view(np.uint32), whole.view(np.uint32))
    assert sum(p[1].tiles for p in parts) == sw.tiles
    assert sum(p[1].samples for p in parts) == sw.samples
    for k in range(3):
        others = sum(parts[j][0] for j in range(3) if j != k)
        assert np.all((parts[k][0] == 0) | (others == 0))


def test_tof_phase_images_give_plausible_distance(gpu):
    """C5: four phase images -> distance via the AMCW formula (sensor_tof_amcw.hpp:171-213); the
    background wall is 2 m in front of the camera."""
    g = Golden("tof")
    s = device_scene(gpu, g)
    sensor = g.sensor
    sensor.width, sensor.height = 88, 72
    D = []
    for phase in range(4):
        sensor.phase_index = phase
        f, _ = s.render(g.camera, sensor, g.params, 8, g.t0, g.t1)
        D.append(f[..., 0] - f[..., 1])
    phase_shift = np.arctan2(D[3] - D[1], D[0] - D[2])
    phase_shift[phase_shift < 0] += 2 * np.pi
    dist = (299792458.0 / sensor.modulation_frequency) * phase_shift / (4 * np.pi)
    corner = dist[:8, :8]
    assert 1.9 < np.median(corner) < 3.2


def test_reference_furnace_app_runs_unchanged(gpu, tmp_path):
    """The reference's wurblpt-furnace-test.cpp, compiled unchanged against include/wurblpt and linked
    against libwurblpt_b200.so (apps/Makefile), renders on the GPU and reproduces the analytic result."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apps", "_bin", "wurblpt-furnace-test")
    if not os.path.exists(exe):
        pytest.skip("apps/_bin not built (needs /root/reference at build time)")
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Rendering 512x512 pixels with 100 samples" in r.stderr
    img = read_wptarray(str(tmp_path / "frame.exr"))
    assert img.shape == (512, 512, 3)
    assert np.all(img[0, 0] == 1.0) and np.all(img[-1, -1] == 1.0)
    centre = img[200:312, 200:312]
    # Most interior pixels are exactly the albedo; a few lose grazing samples that re-hit the sphere from
    # inside (minHitDistance = 1e-5). The reference shows the same at this resolution: centre mean
    # 0.41958, about 8 % of the centre pixels low by up to 0.025 (measured with oracle/_ref).
    assert abs(float(np.median(centre)) - 0.42) < 1e-6
    assert 0.4190 < float(centre.mean()) < 0.4200, float(centre.mean())
    assert np.abs(centre - 0.42).max() < 0.05
    assert (np.abs(centre[..., 0] - 0.42) > 1e-4).mean() < 0.15
    sphere = img[img[..., 0] < 0.9]
    assert sphere.min() > 0.39


def test_reference_mis_app_runs_unchanged(gpu, tmp_path):
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apps", "_bin", "wurblpt-mis-test")
    if not os.path.exists(exe):
        pytest.skip("apps/_bin not built (needs /root/reference at build time)")
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    a = read_wptarray(str(tmp_path / "img-with-material-importance-sampling.tgd"))
    b = read_wptarray(str(tmp_path / "img-with-multiple-importance-sampling.tgd"))
    assert a.shape == b.shape == (540, 960, 3)
    assert abs(a.mean() - b.mean()) <= 0.05 * b.mean()


def test_more_animations_than_cache_entries_with_motion_blur(gpu):
    """The per-thread animation cache (AnimationCache, animation.hpp:52-125) has 16 entries; animations beyond the
    first 15 share the last one. A copy of the ToF scene whose hitables use copies number 17.. and 18.. of the two
    animations must render the same frame as the original, with an exposure interval (every thread evaluates the key
    frames at its ray's time)."""
    import ctypes as C
    from wurblpt_b200 import scene_desc as sd
    g = Golden("tof")
    arrays = {k: v.copy() for k, v in g.host.arrays.items()}
    n_anim = g.desc.n_animations
    assert n_anim >= 1
    copies = 10
    arrays["animations"] = np.tile(arrays["animations"], copies)
    inst = sd.struct_view(arrays["instances"], abi.wpt_instance)
    moved = 0
    for i in range(len(inst)):
        if inst[i].animation >= 0:
            inst[i].animation += n_anim * (copies - 1 - (i % 2))      # the last two copies, alternating
            moved += 1
    sph = sd.struct_view(arrays["spheres"], abi.wpt_sphere)
    for i in range(len(sph)):
        if sph[i].animation >= 0:
            sph[i].animation += n_anim * (copies - 1)
    assert moved > 0 and n_anim * (copies - 1) >= 16
    many = sd.HostScene(arrays)
    t0, t1 = g.t0, g.t0 + 0.02
    sensor = g.sensor
    sensor.height, sensor.width = g.z["render"].shape[:2]
    frames = []
    for desc in (g.desc, many.desc):
        s = gpu.Scene(desc, device=0).update_bvh(t0, t1)
        f, _ = s.render(g.camera, sensor, g.params, 6, t0, t1, gpu.make_opts(substreams=1))
        frames.append(f)
        s.close()
    assert np.isfinite(frames[0]).all() and frames[0].mean() > 0
    assert np.array_equal(frames[0].view(np.uint32), frames[1].view(np.uint32))


@pytest.mark.gpu
def test_trace_rays_refuses_rays_the_reference_would_assert_on(gpu):
    """Ray's constructor asserts a finite, unit-length direction (ray.hpp:47-49); wpt_trace_rays returns WPT_E_INVALID
    for such input instead of walking the tree with it (the boxes of unused child slots are only out of reach of
    unit directions)."""
    g = Golden("cornell")
    s = device_scene(gpu, g)
    o = np.zeros((4, 3), dtype=np.float32)
    d = np.tile(np.array([0.0, 0.0, -1.0], dtype=np.float32), (4, 1))
    ids, dist = s.trace_rays(o, d, g.t0, 1e-4)
    assert ids.shape == (4,)
    for bad in (np.array([10.0, 10.0, 10.0]), np.array([0.0, 0.0, 0.0]), np.array([np.nan, 0.0, 1.0])):
        d2 = d.copy()
        d2[2] = bad
        with pytest.raises(gpu.WptError) as e:
            s.trace_rays(o, d2, g.t0, 1e-4)
        assert e.value.code == abi.WPT_E_INVALID
    o2 = o.copy()
    o2[1, 0] = np.inf
    with pytest.raises(gpu.WptError) as e:
        s.trace_rays(o2, d, g.t0, 1e-4)
    assert e.value.code == abi.WPT_E_INVALID
