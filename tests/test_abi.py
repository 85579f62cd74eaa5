"""CPU tests: the C-ABI library loads, exports what include/wurblpt_b200.h declares, its struct
layouts agree with the ctypes mirror, and it refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from wurblpt_b200 import abi
from util import Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "wurblpt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wpt_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    names = declared_functions()
    assert len(names) >= 12
    lib = capi.lib()
    for n in names:
        assert hasattr(lib, n), "libwurblpt_b200.so does not export %s" % n
    assert set(names) == set(capi.EXPORTS)


def test_struct_layouts_match_library(capi):
    lib = capi.lib()
    for i, S in enumerate(abi.STRUCTS):
        assert lib.wpt_abi_sizeof(i) == C.sizeof(S), S.__name__
    assert lib.wpt_abi_sizeof(len(abi.STRUCTS)) == 0


def test_version_string(capi):
    assert b"sm_100a" in capi.lib().wpt_version()


def test_no_cpu_fallback(capi):
    """Without a device, rendering must fail loudly, never fall back to the host."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    g = Golden("cornell")
    with pytest.raises(capi.WptError) as e:
        capi.Scene(g.desc, device=0)
    assert e.value.code == abi.WPT_E_NO_DEVICE
    s = capi.Scene(g.desc, device=-1).update_bvh()
    with pytest.raises(capi.WptError) as e:
        s.render(g.camera, g.sensor, g.params, 2)
    assert e.value.code == abi.WPT_E_NO_DEVICE
    with pytest.raises(capi.WptError) as e:
        s.primary_hits(g.camera, 8, 8)
    assert e.value.code == abi.WPT_E_NO_DEVICE
    with pytest.raises(capi.WptError) as e:
        s.ground_truth(g.camera, 8, 8)
    assert e.value.code == abi.WPT_E_NO_DEVICE
    with pytest.raises(capi.WptError):
        capi.prng_kat([1, 2], 4)


def test_malformed_descriptions_are_rejected(capi):
    g = Golden("cornell")
    d = g.host.desc
    d.abi_version = 99
    with pytest.raises(capi.WptError) as e:
        capi.Scene(d, device=-1)
    assert e.value.code == abi.WPT_E_INVALID
    d.abi_version = abi.WPT_ABI_VERSION

    g = Golden("cornell")
    comps = np.frombuffer(g.host.arrays["components"], dtype=np.int32).copy()
    comps[1] = 10 ** 6              # first component's ref out of range
    g.host.arrays["components"] = comps.view(np.uint8)
    from wurblpt_b200 import scene_desc as sd
    bad = sd.HostScene(g.host.arrays)
    with pytest.raises(capi.WptError) as e:
        capi.Scene(bad.desc, device=-1)
    assert e.value.code == abi.WPT_E_INVALID

    g = Golden("cornell")
    idx = np.frombuffer(g.host.arrays["indices"], dtype=np.uint32).copy()
    idx[0] = 999
    g.host.arrays["indices"] = idx.view(np.uint8)
    bad = sd.HostScene(g.host.arrays)
    with pytest.raises(capi.WptError):
        capi.Scene(bad.desc, device=-1)


def test_render_before_update_bvh_is_a_state_error(capi):
    if capi.device_count() == 0:
        pytest.skip("needs a GPU to get past device selection")
    g = Golden("cornell")
    s = capi.Scene(g.desc, device=0)
    with pytest.raises(capi.WptError) as e:
        s.render(g.camera, g.sensor, g.params, 2)
    assert e.value.code == abi.WPT_E_STATE
