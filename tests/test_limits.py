"""CPU tests of the documented limits: what the device code cannot do is refused with a clear error
when the scene is ingested - never rendered wrongly - and what it handles in a special way is recognised on the host."""
import ctypes as C

import numpy as np
import pytest

from util import Golden
from wurblpt_b200 import abi
from wurblpt_b200 import scene_desc as sd


def _copy_arrays(g):
    return {k: v.copy() for k, v in g.host.arrays.items()}


def _append(buf, items):
    extra = np.frombuffer(b"".join(bytes(i) for i in items), dtype=np.uint8)
    return np.concatenate([buf, extra])


def _prim_ids_of_media(arrays):
    """Hitable ids (component order, one per triangle / sphere / medium: scene_host.cpp) of the medium components."""
    comps = sd.struct_view(arrays["components"], abi.wpt_component)
    inst = sd.struct_view(arrays["instances"], abi.wpt_instance)
    meshes = sd.struct_view(arrays["meshes"], abi.wpt_mesh)
    ids, next_id = [], 0
    for c in comps:
        if c.type == abi.WPT_COMP_MESH_INSTANCE:
            next_id += meshes[inst[c.ref].mesh].triangle_count
        else:
            if c.type == abi.WPT_COMP_MEDIUM:
                ids.append(next_id)
            next_id += 1
    return ids, next_id


def test_nested_two_sided_materials_are_refused(capi):
    """MaterialTwoSided (material.hpp:193-236) of a two-sided material would need recursion on the device."""
    g = Golden("cornell")
    arrays = _copy_arrays(g)
    n = arrays["materials"].nbytes // C.sizeof(abi.wpt_material)
    assert n >= 2
    inner = abi.wpt_material()
    inner.type = abi.WPT_MAT_TWOSIDED
    inner.normal_tex = -1
    for k in range(5):
        inner.tex[k] = -1
    inner.front, inner.back = 0, 1
    outer = abi.wpt_material()
    outer.type = abi.WPT_MAT_TWOSIDED
    outer.normal_tex = -1
    for k in range(5):
        outer.tex[k] = -1
    outer.front, outer.back = n, -1        # front side = the two-sided material appended before it
    # one level of two-sided materials is fine ...
    ok = dict(arrays)
    ok["materials"] = _append(arrays["materials"], [inner])
    capi.Scene(sd.HostScene(ok).desc, device=-1).update_bvh()
    # ... two levels are not
    bad = dict(arrays)
    bad["materials"] = _append(arrays["materials"], [inner, outer])
    host = sd.HostScene(bad)
    with pytest.raises(capi.WptError) as e:
        capi.Scene(host.desc, device=-1)
    assert e.value.code == abi.WPT_E_UNSUPPORTED
    assert "nested two-sided" in str(e.value)


def test_two_sided_material_with_invalid_sides_is_refused(capi):
    g = Golden("cornell")
    arrays = _copy_arrays(g)
    n = arrays["materials"].nbytes // C.sizeof(abi.wpt_material)
    m = abi.wpt_material()
    m.type = abi.WPT_MAT_TWOSIDED
    m.normal_tex = -1
    for k in range(5):
        m.tex[k] = -1
    m.front, m.back = n + 7, -1
    arrays["materials"] = _append(arrays["materials"], [m])
    host = sd.HostScene(arrays)
    with pytest.raises(capi.WptError) as e:
        capi.Scene(host.desc, device=-1)
    assert e.value.code == abi.WPT_E_INVALID


def test_media_bounded_by_media_are_refused(capi):
    """HitableMedium takes any hitables as its boundary (hitable_medium.hpp:43-52); a medium as the boundary of a medium
    would nest the nested walk."""
    g = Golden("media")
    arrays = _copy_arrays(g)
    medium_ids, n_prims = _prim_ids_of_media(arrays)
    assert len(medium_ids) == 2
    boundary = arrays["boundary_hitables"].view(np.uint32).copy()
    media = sd.struct_view(arrays["media"], abi.wpt_medium)
    first = media[0].first_boundary
    original = int(boundary[first])
    assert original < n_prims
    # a hitable that does not exist
    boundary[first] = n_prims + 5
    arrays["boundary_hitables"] = boundary.view(np.uint8)
    host = sd.HostScene(arrays)
    with pytest.raises(capi.WptError) as e:
        capi.Scene(host.desc, device=-1)
    assert e.value.code == abi.WPT_E_INVALID and "boundary" in str(e.value)
    # the other medium
    boundary[first] = medium_ids[1]
    arrays["boundary_hitables"] = boundary.view(np.uint8)
    host = sd.HostScene(arrays)
    with pytest.raises(capi.WptError) as e:
        capi.Scene(host.desc, device=-1)
    assert e.value.code == abi.WPT_E_UNSUPPORTED and "boundary" in str(e.value)
    # and the scene as it was builds
    boundary[first] = original
    arrays["boundary_hitables"] = boundary.view(np.uint8)
    host = sd.HostScene(arrays)
    assert capi.Scene(host.desc, device=-1).update_bvh(g.t0, g.t1).info().media == 2


def test_boundary_shared_with_the_top_level_keeps_the_nested_walk(capi):
    """The warp-wide boundary test needs the boundary triangles in consecutive slots of their own; a boundary mesh that
    is also a visible hitable of the scene has its slots assigned by the top-level tree, so the medium keeps the nested
    tree walk (and still builds)."""
    g = Golden("media")
    arrays = _copy_arrays(g)
    comps = sd.struct_view(arrays["components"], abi.wpt_component)
    inst = sd.struct_view(arrays["instances"], abi.wpt_instance)
    meshes = sd.struct_view(arrays["meshes"], abi.wpt_mesh)
    media = sd.struct_view(arrays["media"], abi.wpt_medium)
    boundary = arrays["boundary_hitables"].view(np.uint32)
    # which component holds the boundary triangles of the mesh-bounded medium?
    made_visible = False
    next_id = 0
    for c in comps:
        count = meshes[inst[c.ref].mesh].triangle_count if c.type == abi.WPT_COMP_MESH_INSTANCE else 1
        ids = set(range(next_id, next_id + count))
        for m in media:
            b = set(int(x) for x in boundary[m.first_boundary:m.first_boundary + m.boundary_count])
            if c.type == abi.WPT_COMP_MESH_INSTANCE and b and b <= ids and not (c.flags & abi.WPT_COMPF_HITABLE):
                c.flags |= abi.WPT_COMPF_HITABLE
                made_visible = True
        next_id += count
    assert made_visible
    host = sd.HostScene(arrays)
    info = capi.Scene(host.desc, device=-1).update_bvh(g.t0, g.t1).info()
    assert info.media == 2 and info.warp_tested_media == 0
