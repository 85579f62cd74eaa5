"""CPU tests pinning the oracle: the compiled reference must reproduce the committed known answers
(and the values quoted in SURVEY.md / BASELINE.md), and scenes/configs.hpp must describe the same
scenes as the reference example apps' own createScene() functions."""
import numpy as np
import pytest

from util import Golden, GOLDEN

# generated from libwurblpt/prng.hpp during the survey
SURVEY_KAT = {0: [0.136347055, 0.812755167, 0.70104903, 0.946349502],
              42: [0.222297013, 0.64561826, 0.663711309, 0.301135063]}


def test_reference_prng_matches_survey_values(ref):
    v = ref.prng_kat([0, 42], 4)
    assert np.allclose(v[0], SURVEY_KAT[0], rtol=0, atol=5e-9)
    assert np.allclose(v[1], SURVEY_KAT[42], rtol=0, atol=5e-9)


def test_reference_prng_matches_golden(ref):
    with np.load(GOLDEN + "/prng_kat.npz") as z:
        assert np.array_equal(ref.prng_kat(z["seeds"], z["values"].shape[1]), z["values"])


def test_golden_prng_matches_survey_values():
    with np.load(GOLDEN + "/prng_kat.npz") as z:
        seeds = list(z["seeds"])
        assert np.allclose(z["values"][seeds.index(0)][:4], SURVEY_KAT[0], atol=5e-9)
        assert np.allclose(z["values"][seeds.index(42)][:4], SURVEY_KAT[42], atol=5e-9)


@pytest.mark.parametrize("name", ["cornell", "rtiow", "tof", "media", "mis", "mis_nohot"])
def test_config_recipes_equal_reference_apps(ref, name):
    """scenes/configs.hpp (variant 0) flattens to the same bytes as the reference app (variant 1)."""
    from wurblpt_b200 import scene_desc as sd
    a = sd.desc_to_arrays(ref.RefScene(name, 1.0, 0).desc)
    b = sd.desc_to_arrays(ref.RefScene(name, 1.0, 1).desc)
    assert sd.diff_descs(a, b) == []


@pytest.mark.parametrize("name", ["cornell", "mis", "furnace", "tof"])
def test_reference_reproduces_golden_primary_hits(ref, name):
    g = Golden(name)
    r = ref.RefScene(name, g.scale, 0)
    r.update_bvh(g.t0, g.t1)
    h, w = g.z["hit_ids"].shape
    ids, dist = r.primary_hits(g.camera, w, h, g.t0, g.params.min_hit_distance)
    assert np.array_equal(ids, g.z["hit_ids"])
    assert np.array_equal(dist.view(np.uint32), g.z["hit_dist"].view(np.uint32))


def test_reference_render_is_deterministic_and_matches_golden(ref):
    g = Golden("cornell")
    r = ref.RefScene("cornell", 1.0, 0)
    r.update_bvh()
    s = g.sensor
    s.height, s.width = g.z["render"].shape[:2]
    frame, _ = r.render(g.camera, s, g.params, int(g.z["render_spp_sqrt"]))
    assert np.array_equal(frame.view(np.uint32), g.z["render"].view(np.uint32))


def test_furnace_golden_is_analytic():
    """reference wurblpt-furnace-test: background exactly 1, sphere interior = albedo 0.42"""
    g = Golden("furnace")
    img = g.z["render"]
    ids = None
    bg = img[(img[..., 0] > 0.99)]
    assert np.all(bg == 1.0)
    interior = img[(img[..., 0] < 0.5)]
    assert len(interior) > 100
    assert np.median(interior) == pytest.approx(0.42, abs=1e-5)
    assert interior.min() > 0.39
