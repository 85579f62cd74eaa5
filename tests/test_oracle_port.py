"""CPU tests pinning the oracle's own restatement (oracle/port.cpp) against outputs of the reference:
the committed fixtures (generated from the compiled reference) and, where it is built, the live
reference. Everything here is integer / index work and must be bit-exact."""
import numpy as np
import pytest

from util import Golden, GOLDEN

STATIC_SCENES = ["cornell", "rtiow", "mis", "mis_nohot", "furnace", "mixed", "synth"]


@pytest.fixture(scope="module")
def port():
    from oracle import portapi
    if not portapi.available():
        pytest.skip("oracle/_build/liboracle_port.so not built (make -C oracle port)")
    return portapi


def test_port_prng_matches_golden(port):
    with np.load(GOLDEN + "/prng_kat.npz") as z:
        assert np.array_equal(port.prng_kat(z["seeds"], z["values"].shape[1]).view(np.uint32), z["values"].view(np.uint32))


@pytest.mark.parametrize("name", STATIC_SCENES)
def test_port_tree_equals_reference_tree(port, name):
    g = Golden(name)
    boxes, link, leaf = port.bvh_nodes(g.desc)
    assert np.array_equal(link, g.z["bvh_link"])
    assert np.array_equal(leaf, g.z["bvh_leaf"])
    assert np.array_equal(boxes.view(np.uint32), g.z["bvh_boxes"].view(np.uint32))


@pytest.mark.parametrize("name", STATIC_SCENES)
def test_port_primary_hits_equal_reference(port, name):
    g = Golden(name)
    h, w = g.z["hit_ids"].shape
    ids, dist = port.primary_hits(g.desc, g.camera, w, h, g.params.min_hit_distance)
    assert np.array_equal(ids, g.z["hit_ids"])
    assert np.array_equal(dist.view(np.uint32), g.z["hit_dist"].view(np.uint32))


def test_port_refuses_what_it_does_not_cover(port):
    for name in ("media", "tof"):
        g = Golden(name)
        with pytest.raises(NotImplementedError):
            port.bvh_nodes(g.desc)


@pytest.mark.parametrize("name,scale", [("cornell", 1.0), ("rtiow", 1.0), ("synth", 0.02)])
def test_port_incoherent_rays_equal_live_reference(port, ref, name, scale):
    r = ref.RefScene(name, scale, 0)
    r.update_bvh()
    rng = np.random.default_rng(99)
    n = 50000
    boxes, _, _ = ref.bvh_nodes(r)
    lo, hi = np.maximum(boxes[0, :3], -30.0), np.minimum(boxes[0, 3:], 30.0)
    o = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d = (d / np.sqrt((d * d).sum(axis=1, keepdims=True))).astype(np.float32)
    rid, ra = r.trace_rays(o, d, 0.0, 1e-4)
    pid, pa = port.trace_rays(r.desc, o, d, 1e-4)
    assert np.array_equal(rid, pid)
    assert np.array_equal(ra.view(np.uint32), pa.view(np.uint32))
