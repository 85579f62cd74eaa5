"""GPU parity tests (-m gpu) for getGroundTruth() (reference wurblpt.hpp:643-759): the device kernel
behind wpt_get_ground_truth() against the compiled reference (live, and committed fixtures).

The images are deterministic (pixel-centre rays, no sampling). Triangle hit records use the same
un-contracted FP32 arithmetic as the reference and must match bit for bit; sphere hit records and
surround cameras go through atan2/asin/sin/cos, where CUDA's libm and glibc differ by a few ulp, so
those are compared with the tolerance written below.
"""
import os
import subprocess

import numpy as np
import pytest

from wurblpt_b200 import abi
from util import GOLDEN, Golden, box_downsample, read_png, read_wptarray, rel_rmse

pytestmark = pytest.mark.gpu

REFERENCE = abi.WPT_TRAVERSAL_REFERENCE
WIDE = abi.WPT_TRAVERSAL_WIDE
RTOL, ATOL = 1e-4, 2e-5     # transcendental functions (see the module docstring)


@pytest.fixture(scope="module")
def gpu(capi):
    if capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (there is no CPU fallback)")
    return capi


def compare_images(got, want, exact, allowed_mismatch=0):
    """Every image of `got` against `want`. Pixels where the two disagree about the hitable's material are
    silhouette pixels decided by a rounding difference in the camera ray; at most `allowed_mismatch` of them."""
    same = got["materials"] == want["materials"]
    assert (~same).sum() <= allowed_mismatch, int((~same).sum())
    for name, comps in zip(abi.WPT_GT_NAMES, abi.WPT_GT_COMPONENTS):
        if name not in got:
            continue
        a, b = got[name][same], want[name][same]
        assert a.shape[-1] == comps
        if exact:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
        else:
            assert np.allclose(a, b, rtol=RTOL, atol=ATOL), (name, float(np.abs(a - b).max()))


@pytest.mark.parametrize("name,scale,exact", [
    ("cornell", 1.0, True), ("tof", 1.0, True), ("synth", 0.004, True),
    ("rtiow", 1.0, False), ("mixed", 1.0, False), ("media", 1.0, False), ("synth360", 0.004, False)])
@pytest.mark.parametrize("mode", [REFERENCE, WIDE])
def test_ground_truth_against_live_reference(gpu, ref, name, scale, exact, mode):
    view = ref.View(name)
    t0, tp, tn = (1.0, 0.96, 1.04) if name == "tof" else (view.t0, view.t0, view.t0)
    r = ref.RefScene(name, scale, 0)
    r.update_bvh(t0, t0)
    s = gpu.Scene(r.desc, device=0).update_bvh(t0, t0)
    w, h = 160, 120
    min_hit = view.params.min_hit_distance
    want = r.ground_truth(view.camera, w, h, t0, tp, tn, min_hit)
    bits = abi.WPT_GT_ALL
    if view.camera.surround_mode != 0 or view.camera.stereoscopic_distance > 0:
        bits &= ~abi.WPT_GT_PIXEL_OFFSETS        # the reference asserts (camera.hpp:207-208)
    got = s.ground_truth(view.camera, w, h, t0, tp, tn, min_hit, bits, opts=gpu.make_opts(traversal=mode))
    assert (want["materials"] >= 0).any()
    # exact ties between two hitables may be settled differently by the ordered walk (counted)
    allowed = 0 if (exact and mode == REFERENCE) else w * h // 500
    compare_images(got, want, exact and mode == REFERENCE, allowed)
    if name == "tof":
        assert np.abs(got["world_offset_to_next"]).max() > 1e-3      # the animated quad moves
        assert np.abs(got["pixel_offset_to_prev"]).max() > 1e-2


@pytest.mark.parametrize("name", ["tof", "mixed"])
def test_ground_truth_against_committed_fixture(gpu, name):
    g = Golden(name)
    with np.load(os.path.join(GOLDEN, "gt_" + name + ".npz")) as z:
        want = {k: z[k] for k in z.files}
    t0, tp, tn = [float(v) for v in want["times"]]
    h, w = want["materials"].shape
    s = gpu.Scene(g.desc, device=0).update_bvh(t0, t0)
    got = s.ground_truth(g.camera, w, h, t0, tp, tn, g.params.min_hit_distance,
                         opts=gpu.make_opts(traversal=REFERENCE))
    compare_images(got, want, exact=(name == "tof"))


def test_ground_truth_selects_images_and_maps_material_ids(gpu):
    g = Golden("cornell")
    s = gpu.Scene(g.desc, device=0).update_bvh(0.0, 0.0)
    bits = (1 << 0) | (1 << 11) | abi.WPT_GT_MATERIALS
    ids = np.arange(g.desc.n_materials, dtype=np.int32) + 100
    got = s.ground_truth(g.camera, 64, 48, bits=bits, material_ids=ids)
    assert sorted(got) == ["camera_distances", "materials", "world_positions"]
    hit = got["materials"] >= 0
    assert hit.any() and (got["materials"][hit] >= 100).all()
    assert (got["camera_distances"][hit] > 0).all() and (got["camera_distances"][~hit] == 0).all()
    with pytest.raises(gpu.WptError):
        s.ground_truth(g.camera, 64, 48, t0=1.0, t_prev=2.0)        # t_prev <= t0 is required


def test_reference_tof_example_app_runs_unchanged(gpu, tmp_path):
    """wurblpt-tof-example.cpp, compiled unchanged against include/wurblpt: 4 phase images per frame, then
    getGroundTruth() with flow images (wurblpt-tof-example.cpp:125-151). The app renders 5 s of video; the
    first frames are enough here."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apps", "_bin", "wurblpt-tof-example")
    if not os.path.exists(exe):
        pytest.skip("apps/_bin/wurblpt-tof-example not built")
    log = ""
    try:
        log = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120).stderr
    except subprocess.TimeoutExpired as e:
        log = (e.stderr or b"").decode(errors="replace") if isinstance(e.stderr, bytes) else (e.stderr or "")
    assert (tmp_path / "frame0000-groundtruth-distances.tgd").exists(), log[-3000:]
    dist = read_wptarray(str(tmp_path / "frame0000-groundtruth-distances.tgd"))
    result = read_wptarray(str(tmp_path / "frame0000-results.tgd"))
    flow = read_wptarray(str(tmp_path / "frame0000-groundtruth-psflow-forward.tgd"))
    assert dist.shape[:2] == (288, 352) and flow.shape[2] == 2
    # the simulated ToF distance agrees with the ground truth distance on the diffuse background wall
    measured = result[..., 0]
    ok = (dist[..., 0] > 0) & np.isfinite(measured) & (measured > 0)
    rel = np.abs(measured[ok] - dist[..., 0][ok]) / dist[..., 0][ok]
    assert np.median(rel) < 0.05
    assert np.abs(flow).max() > 0.01


def test_reference_material_playground_app_runs_unchanged(gpu, tmp_path):
    """wurblpt-material-playground.cpp, unchanged: the camera and two mesh instances follow user-defined Animation
    subclasses (not key frames), which the API mirror samples for every frame's updateBVH(t0, t0). The app
    renders 125 frames of 800x600 at 225 spp; the first few are enough here."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apps", "_bin",
                       "wurblpt-material-playground")
    if not os.path.exists(exe):
        pytest.skip("apps/_bin/wurblpt-material-playground not built")
    log = ""
    try:
        log = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=25).stderr
    except subprocess.TimeoutExpired as e:
        log = (e.stderr or b"").decode(errors="replace") if isinstance(e.stderr, bytes) else (e.stderr or "")
    frames = sorted(p.name for p in tmp_path.glob("frame-*.pnm"))
    assert len(frames) >= 3, log[-3000:]

    def read_ppm(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"P6"
            w, h = [int(v) for v in f.readline().split()]
            assert f.readline().strip() == b"255"
            return np.frombuffer(f.read(), dtype=np.uint8).reshape(h, w, 3)

    a, b = read_ppm(tmp_path / frames[0]), read_ppm(tmp_path / frames[2])
    assert a.shape == (600, 800, 3)
    assert 5 < a.mean() < 250 and a.std() > 5          # an image, not a constant
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).mean() > 0.5     # the camera and the objects moved
    # against the same app rendered by the reference itself (tests/golden/make_app_golden.py): 8-bit sRGB frames,
    # rows stored top to bottom in the file, 8x8 box filter
    with np.load(os.path.join(GOLDEN, "app_material_playground.npz")) as z:
        for k, frame in ((0, a), (2, b)):
            ref_small = z["frame%d" % k].astype(np.float32)
            dev_small = box_downsample(frame.astype(np.float32), int(z["box"]))
            err = rel_rmse(dev_small, ref_small, floor=4.0)
            m_dev, m_ref = dev_small.reshape(-1, 3).mean(axis=0), ref_small.reshape(-1, 3).mean(axis=0)
            print("material-playground frame %d: relRMSE %.4f, means %s vs %s" % (k, err, m_dev, m_ref))
            assert err <= 0.05, (k, err)
            assert np.all(np.abs(m_dev - m_ref) <= 0.02 * m_ref), (k, m_dev, m_ref)


def test_reference_toomuch_app_runs_unchanged(gpu, tmp_path):
    """wurblpt-toomuch.cpp, unchanged, frame 0: eleven user-defined Animation subclasses (lights, objects, camera)
    with motion blur, a user-defined time-dependent Texture on the lights (sampled into images), mirror sphere,
    glass pillars, ModPhong objects, 21 hot spots."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apps", "_bin", "wurblpt-toomuch")
    if not os.path.exists(exe):
        pytest.skip("apps/_bin/wurblpt-toomuch not built")
    r = subprocess.run([exe, "0"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Rendering 400x300 pixels with 100 samples" in r.stderr
    hdr = read_wptarray(str(tmp_path / "toomuch-0000.tgd"))
    assert hdr.shape == (300, 400, 3) and np.isfinite(hdr).all()
    assert hdr.mean() > 0.01 and (hdr > 0).mean() > 0.9      # a lit room, not a black frame
    # against frame 0 of the same app rendered by the reference itself (tests/golden/make_app_golden.py): HDR frame,
    # 4x4 box filter; both are 100-spp estimates of a scene lit by 21 small moving lights
    with np.load(os.path.join(GOLDEN, "app_toomuch.npz")) as z:
        ref_small = z["frame0"]
        dev_small = box_downsample(hdr, int(z["box"]))
    m_dev, m_ref = dev_small.reshape(-1, 3).mean(axis=0), ref_small.reshape(-1, 3).mean(axis=0)
    err = rel_rmse(dev_small, ref_small, floor=1e-2 * float(m_ref.mean()) ** 2)
    print("toomuch frame 0: relRMSE %.4f, means %s vs %s" % (err, m_dev, m_ref))
    assert np.all(np.abs(m_dev - m_ref) <= 0.04 * m_ref), (m_dev, m_ref)
    assert err <= 0.25, err


def test_reference_noise_textures_app_runs_unchanged(gpu, tmp_path):
    """wurblpt-noise-textures.cpp, unchanged: value / gradient / Worley / Perlin noise textures combined by
    user-defined Texture classes that switch with the render time; ten renders at t = 0..9 after one updateBVH()."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apps", "_bin", "wurblpt-noise-textures")
    if not os.path.exists(exe):
        pytest.skip("apps/_bin/wurblpt-noise-textures not built")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=200)
    assert r.returncode == 0, r.stderr[-3000:]
    images = [read_png(str(tmp_path / ("noise-example-%d.png" % t))) for t in range(10)]
    assert all(im.shape == (600, 800, 3) and im.dtype == np.uint8 for im in images)
    assert all(im.std() > 3 for im in images)
    # the textures are functions of time: consecutive renders differ by more than Monte-Carlo noise
    d01 = np.abs(images[0].astype(np.int32) - images[1].astype(np.int32)).mean()
    assert d01 > 1.0, d01
    # against the ten images of the same app rendered by the reference itself (tests/golden/make_app_golden.py):
    # 8-bit sRGB, 8x8 box filter. The noise textures are evaluated on the host with the reference's arithmetic and
    # sampled into images of the size the reference gives them (texture.hpp:100-105).
    with np.load(os.path.join(GOLDEN, "app_noise_textures.npz")) as z:
        for t in range(10):
            ref_small = z["t%d" % t].astype(np.float32)
            dev_small = box_downsample(images[t].astype(np.float32), int(z["box"]))
            err = rel_rmse(dev_small, ref_small, floor=4.0)
            m_dev, m_ref = dev_small.reshape(-1, 3).mean(axis=0), ref_small.reshape(-1, 3).mean(axis=0)
            print("noise-textures t=%d: relRMSE %.4f, means %s vs %s" % (t, err, m_dev, m_ref))
            assert err <= 0.06, (t, err)
            assert np.all(np.abs(m_dev - m_ref) <= 0.02 * m_ref + 0.5), (t, m_dev, m_ref)
