"""GPU parity tests (-m gpu) for cameras and sensors the default views do not use: the three lens-distortion
models (optics.hpp:226-309), Camera::Surround_180 (camera.hpp:160-169) and SensorRGB with its gates set
(sensor_rgb.hpp:63-80). Fixtures: tests/golden/view_*.npz, generated from the compiled reference by
tests/golden/make_view_golden.py."""
import numpy as np
import pytest

from wurblpt_b200 import abi
from util import GoldenView, VIEW_GOLDENS, box_downsample, compare_hits, rel_rmse

pytestmark = pytest.mark.gpu

REFERENCE = abi.WPT_TRAVERSAL_REFERENCE
WIDE = abi.WPT_TRAVERSAL_WIDE


@pytest.fixture(scope="module")
def gpu(capi):
    if capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (there is no CPU fallback)")
    return capi


@pytest.mark.parametrize("name", VIEW_GOLDENS)
def test_view_primary_hits(gpu, name):
    """Pixel-centre rays through the view's camera. Perspective cameras with lens distortion: undistortion is
    polynomial FP32 arithmetic (the OpenCV model iterates it), so the rays, hence ids and distances, are the
    reference's bit for bit. Surround cameras build the ray with sin/cos (CUDA's and glibc's differ by <= 2 ulp)."""
    g = GoldenView(name)
    s = gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)
    h, w = g.z["hit_ids"].shape
    for mode in (REFERENCE, WIDE):
        ids, dist = s.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance, gpu.make_opts(traversal=mode))
        c = compare_hits(ids, dist, g.z["hit_ids"], g.z["hit_dist"])
        same = (ids == g.z["hit_ids"]) & (ids >= 0)
        if g.camera.surround_mode != 0:
            assert c["wrong"] + c["near"] + c["ties"] <= max(2, c["total"] // 1000), (name, mode, c)
            assert np.allclose(dist[same], g.z["hit_dist"][same], rtol=2e-5, atol=0.0)
            continue
        assert c["wrong"] == 0, (name, mode, c)
        if mode == REFERENCE:
            assert c["ties"] == 0 and c["near"] == 0, (name, c)
        else:
            assert c["ties"] + c["near"] <= max(4, c["total"] // 500), (name, c)
        assert np.array_equal(dist[same].view(np.uint32), g.z["hit_dist"][same].view(np.uint32)), (name, mode)
    s.close()


# relative RMSE of 8x8 box-filtered images / relative difference of the channel means; the device renders with 4x
# the fixture's samples, so the bound is the noise of the reference render alone (measured: two reference renders
# of these views with different seeds differ by 3-6 % / 1-2 %). Gated sensors keep only part of the light paths.
VIEW_TOL = {"cornell_radial_planar": (0.08, 0.03), "cornell_radial_only": (0.08, 0.03), "cornell_opencv": (0.08, 0.03),
            "mixed_opencv": (0.15, 0.05), "cornell_surround180": (0.15, 0.06), "mixed_surround180_stereo": (0.20, 0.06),
            "cornell_gated": (0.12, 0.05), "mis_gated": (0.12, 0.04)}


@pytest.mark.parametrize("name", VIEW_GOLDENS)
def test_view_render_matches_reference_render(gpu, name):
    g = GoldenView(name)
    s = gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)
    sensor = g.sensor
    sensor.height, sensor.width = g.z["render"].shape[:2]
    spp = 2 * int(g.z["render_spp_sqrt"])
    frame, stats = s.render(g.camera, sensor, g.params, spp, g.t0, g.t1)
    ref_img = g.z["render"]
    assert np.isfinite(frame).all()
    assert stats.samples == frame.shape[0] * frame.shape[1] * spp * spp
    tol_rmse, tol_mean = VIEW_TOL[name]
    m_dev = frame.reshape(-1, 3).mean(axis=0)
    m_ref = ref_img.reshape(-1, 3).mean(axis=0)
    assert np.all(np.abs(m_dev - m_ref) <= tol_mean * np.maximum(m_ref, 1e-4)), (name, m_dev, m_ref)
    e = rel_rmse(box_downsample(frame, 8), box_downsample(ref_img, 8), floor=1e-2 * float(m_ref.mean()) ** 2 + 1e-12)
    assert e <= tol_rmse, (name, e)
    s.close()


def test_gates_change_the_image(gpu):
    """The gated fixture must differ from the ungated render of the same view (the gates are exercised), and
    an always-open gate given explicitly must equal the ungated render bit for bit."""
    g = GoldenView("cornell_gated")
    s = gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)
    sensor = g.sensor
    sensor.height, sensor.width = 64, 64
    gated, _ = s.render(g.camera, sensor, g.params, 8, g.t0, g.t1, gpu.make_opts(substreams=1))
    fmax = float(np.finfo(np.float32).max)
    sensor.min_path_len, sensor.max_path_len, sensor.min_dist_to_light, sensor.max_dist_to_light = 0.0, fmax, 0.0, fmax
    open_, _ = s.render(g.camera, sensor, g.params, 8, g.t0, g.t1, gpu.make_opts(substreams=1))
    assert gated.mean() < 0.6 * open_.mean()
    # a gate that passes everything but is "set" (min > 0 would be needed to force the gated kernel variant)
    sensor.min_path_len = 1e-30
    forced, _ = s.render(g.camera, sensor, g.params, 8, g.t0, g.t1, gpu.make_opts(substreams=1))
    assert abs(forced.mean() - open_.mean()) <= 0.03 * open_.mean()
    s.close()
