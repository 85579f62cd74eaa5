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


def _boundary_rays(corners, rng, n, targets=None):
    """Rays around a boundary mesh: origins inside and outside its box; towards `targets` (exact points) or, without
    them, half of the rays in random directions and half towards random points of the box."""
    pts = corners.reshape(-1, 3)
    centre, ext = 0.5 * (pts.min(0) + pts.max(0)), 0.5 * (pts.max(0) - pts.min(0))

    def unit(v):
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    if targets is None:
        o = (centre + (rng.random((n, 3)) * 2 - 1) * ext * 2.0).astype(np.float32)
        d = unit(rng.normal(size=(n, 3)))
        aim = centre + (rng.random((n // 2, 3)) * 2 - 1) * ext * 0.9
        d[:n // 2] = unit(aim - o[:n // 2])
    else:
        o = (centre + unit(rng.normal(size=(n, 3))) * np.linalg.norm(ext) * rng.uniform(0.0, 3.0, (n, 1))).astype(np.float32)
        d = unit(targets[rng.integers(0, len(targets), n)] - o)
    return o, d.astype(np.float32)


def test_medium_boundary_replay_equals_the_nested_walks():
    """HitableMedium::hit asks the boundary's BVH twice (hitable_medium.hpp:75-79). For a small triangle boundary the
    device tests every boundary triangle once on the whole line and replays both queries over the triangles that were
    hit (device_trace.cuh: mediumRoundCoop, mediumTest). oracle/port.cpp computes both forms with the restated
    reference arithmetic: on random rays the two distances are the reference's bit for bit."""
    from oracle import portapi
    if not portapi.available():
        pytest.skip("oracle/_build/liboracle_port.so not built")
    g = Golden("media")
    medium = [i for i in range(g.desc.n_media) if g.desc.media[i].boundary_count > 1]
    assert len(medium) == 1
    rng = np.random.default_rng(7)
    _, _, corners = portapi.medium_boundary(g.desc, medium[0], np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32))
    assert corners.shape == (20, 3, 3)          # the icosahedron around the dense medium
    o, d = _boundary_rays(corners, rng, 500_000)
    nested, replay, _ = portapi.medium_boundary(g.desc, medium[0], o, d, g.params.min_hit_distance)
    assert np.isfinite(nested).all(1).sum() > 100_000       # plenty of rays cross the boundary twice
    assert np.array_equal(nested.view(np.uint32), replay.view(np.uint32))
    # rays through the middle of a triangle: likewise
    o, d = _boundary_rays(corners, rng, 100_000, corners.mean(1))
    nested, replay, _ = portapi.medium_boundary(g.desc, medium[0], o, d, g.params.min_hit_distance)
    assert np.array_equal(nested.view(np.uint32), replay.view(np.uint32))


def test_medium_boundary_replay_on_rays_through_vertices_and_edges():
    """Rays aimed EXACTLY at a vertex or an edge of the boundary hit two or more triangles at (nearly) the same distance
    and graze the boxes of the nested tree. There the two forms may differ: by which of the coincident triangles is
    accepted (distances a few ulp apart), or because a box test of the reference's nested walk culls a crossing that the
    watertight triangle test finds - never the other way round. Counted here so that a change of either form shows."""
    from oracle import portapi
    if not portapi.available():
        pytest.skip("oracle/_build/liboracle_port.so not built")
    g = Golden("media")
    medium = [i for i in range(g.desc.n_media) if g.desc.media[i].boundary_count > 1][0]
    rng = np.random.default_rng(11)
    _, _, corners = portapi.medium_boundary(g.desc, medium, np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32))
    verts = np.unique(corners.reshape(-1, 3), axis=0)
    mids = np.unique(np.array([0.5 * (t[a] + t[b]) for t in corners for a, b in ((0, 1), (1, 2), (2, 0))]), axis=0)
    assert len(verts) == 12 and len(mids) == 30
    o, d = _boundary_rays(corners, rng, 300_000, np.concatenate([verts, mids]))
    nested, replay, _ = portapi.medium_boundary(g.desc, medium, o, d, g.params.min_hit_distance)
    differ = ~(nested.view(np.uint32) == replay.view(np.uint32)).all(1)
    assert differ.mean() < 0.03
    # every crossing the nested walk reports is one the replay reports too, to rounding (when a box test culls the FIRST
    # crossing, the nested walk's first distance is the replay's second) ...
    for k in range(2):
        have = np.isfinite(nested[:, k])
        with np.errstate(invalid="ignore"):
            gap = np.abs(replay - nested[:, k:k + 1]) / np.maximum(np.abs(nested[:, k:k + 1]), 1e-30)
        gap = np.where(np.isfinite(replay), gap, np.inf).min(1)
        assert (gap[have] <= 1e-5).all()
    # ... and the replay never reports fewer crossings than the nested walk
    assert (np.isfinite(replay).sum(1) >= np.isfinite(nested).sum(1)).all()
