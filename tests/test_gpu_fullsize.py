"""GPU tests at BASELINE.json's full sizes, through properties that do not need a stored image:
exact agreement of primary-ray ids with the live reference on the 1M-triangle scene at 1920x1080,
tile-partition additivity and sample accounting at full resolution."""
import numpy as np
import pytest

from wurblpt_b200 import abi
from util import compare_hits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth(capi, ref):
    if capi.device_count() == 0:
        pytest.fail("no CUDA device")
    r = ref.RefScene("synth", 1.0, 0)
    r.update_bvh()
    s = capi.Scene(r.desc, device=0).update_bvh()
    return r, s, ref.View("synth")


def test_one_million_triangles(synth):
    r, s, view = synth
    info = s.info()
    assert 0.99e6 <= info.triangles <= 1.01e6, info.triangles
    assert info.binary_nodes == 2 * info.hitables - 1


def test_primary_ids_1080p_against_reference(synth, capi):
    r, s, view = synth
    w, h = 1920, 1080
    rid, ra = r.primary_hits(view.camera, w, h, 0.0, view.params.min_hit_distance)
    ids, dist = s.primary_hits(view.camera, w, h, 0.0, view.params.min_hit_distance,
                               capi.make_opts(traversal=abi.WPT_TRAVERSAL_REFERENCE))
    c = compare_hits(ids, dist, rid, ra)
    assert c["wrong"] == 0 and c["ties"] == 0 and c["near"] == 0, c
    assert c["dist_bit_equal"] == c["same"], c
    ids, dist = s.primary_hits(view.camera, w, h, 0.0, view.params.min_hit_distance,
                               capi.make_opts(traversal=abi.WPT_TRAVERSAL_WIDE))
    c = compare_hits(ids, dist, rid, ra)
    print("wide walk vs reference at 1920x1080:", c)
    assert c["wrong"] == 0, c
    assert c["ties"] + c["near"] <= 200, c


def test_full_resolution_partition_and_accounting(synth, capi):
    r, s, view = synth
    sensor = view.sensor
    whole, sw = s.render(view.camera, sensor, view.params, 2)
    assert sw.samples == 1920 * 1080 * 4
    a, sa = s.render(view.camera, sensor, view.params, 2, opts=capi.make_opts(tile_first=0, tile_stride=2))
    b, sb = s.render(view.camera, sensor, view.params, 2, opts=capi.make_opts(tile_first=1, tile_stride=2))
    assert np.array_equal((a + b).view(np.uint32), whole.view(np.uint32))
    assert sa.samples + sb.samples == sw.samples
    assert np.isfinite(whole).all() and whole.mean() > 0
