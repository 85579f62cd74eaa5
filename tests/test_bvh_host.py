"""CPU tests of Scene::updateBVH's host side: the SAH tree must be the reference's tree, node for
node and bit for bit (bvh.hpp:93-246), against committed fixtures and against the live reference."""
import numpy as np
import pytest

from util import Golden, GOLDEN_SCENES


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_tree_equals_golden_reference_tree(capi, name):
    g = Golden(name)
    s = capi.Scene(g.desc, device=-1).update_bvh(g.t0, g.t1)
    boxes, link, leaf = s.binary_nodes()
    assert np.array_equal(link, g.z["bvh_link"])
    assert np.array_equal(leaf, g.z["bvh_leaf"])
    assert np.array_equal(boxes.view(np.uint32), g.z["bvh_boxes"].view(np.uint32))
    info = s.info()
    assert info.binary_nodes == len(link)
    assert info.wide_nodes <= max(1, info.binary_nodes // 2 + 1)
    assert info.wide_depth <= info.binary_depth


@pytest.mark.parametrize("name,scale,times", [
    ("tof", 1.0, (1.0, 1.25)), ("tof", 1.0, (2.4, 2.6)), ("synth", 0.05, (0.0, 0.0)), ("rtiow", 1.0, (0.0, 0.0))])
def test_tree_equals_live_reference_tree(capi, ref, name, scale, times):
    r = ref.RefScene(name, scale, 0)
    r.update_bvh(*times)
    rb = ref.bvh_nodes(r)
    s = capi.Scene(r.desc, device=-1).update_bvh(*times)
    mb = s.binary_nodes()
    assert np.array_equal(rb[1], mb[1]) and np.array_equal(rb[2], mb[2])
    assert np.array_equal(rb[0].view(np.uint32), mb[0].view(np.uint32))


def test_every_hitable_appears_exactly_once(capi):
    g = Golden("mixed")
    s = capi.Scene(g.desc, device=-1).update_bvh()
    _, link, leaf = s.binary_nodes()
    ids = np.sort(leaf[leaf >= 0])
    assert np.array_equal(ids, np.arange(len(ids)))
    inner = leaf < 0
    assert inner.sum() == len(ids) - 1
    assert (link[inner] > np.nonzero(inner)[0] + 1).all()


def test_empty_scene_builds(capi):
    g = Golden("furnace")
    arrays = dict(g.host.arrays)
    arrays["components"] = np.zeros(0, dtype=np.uint8)
    arrays["spheres"] = np.zeros(0, dtype=np.uint8)
    from wurblpt_b200 import scene_desc as sd
    s = capi.Scene(sd.HostScene(arrays).desc, device=-1).update_bvh()
    assert s.info().hitables == 0


def test_coincident_triangles_build_a_valid_tree(capi):
    """Exact ties in the SAH sweep (every triangle of a mesh instanced twice at the same place): the builder must
    terminate, keep one hitable per leaf and list every hitable once; this is also the input on which the device
    has to settle exact-distance ties the reference's way."""
    g = Golden("cornell")
    arrays = dict(g.host.arrays)
    from wurblpt_b200 import scene_desc as sd
    comps = arrays["components"].copy()
    arrays["components"] = np.concatenate([comps, comps])      # every component a second time
    s = capi.Scene(sd.HostScene(arrays).desc, device=-1).update_bvh()
    n = s.info().hitables
    assert n == 2 * capi.Scene(g.desc, device=-1).update_bvh().info().hitables
    boxes, link, leaf = s.binary_nodes()
    assert len(leaf) == 2 * n - 1
    assert np.array_equal(np.sort(leaf[leaf >= 0]), np.arange(n))
    assert np.isfinite(boxes).all() and (boxes[:, :3] <= boxes[:, 3:]).all()


def test_single_hitable_tree(capi):
    g = Golden("furnace")       # one sphere
    s = capi.Scene(g.desc, device=-1).update_bvh()
    boxes, link, leaf = s.binary_nodes()
    assert s.info().hitables == 1 and len(leaf) == 1 and leaf[0] == 0


def test_small_triangle_boundaries_of_media_are_marked_for_the_warp_wide_test(capi):
    """HitableMedium::hit (hitable_medium.hpp:69-98) asks the boundary's BVH twice. A boundary that is a mesh of up to 32
    triangles gets consecutive triangle slots in the nested tree's storage order and is tested by a whole warp of the
    trace kernel; other boundaries keep the nested walk. The media fixture has one of each: the icosahedron around the
    dense medium and the sphere around the thin one."""
    g = Golden("media")
    assert g.desc.n_media == 2
    s = capi.Scene(g.desc, device=-1).update_bvh(g.t0, g.t1)
    info = s.info()
    assert info.media == 2
    assert info.warp_tested_media == 1
    # scenes without media: nothing to mark
    cornell = Golden("cornell")     # (keeps the arrays behind the description alive)
    c = capi.Scene(cornell.desc, device=-1).update_bvh()
    assert c.info().warp_tested_media == 0
