"""Benchmark / test scenes built by the C++ API mirror (scenes/configs.hpp -> lib/libwpt_scenes.so).

Product code: no dependency on oracle/. Gives flat scene descriptions and default views by name:
cornell, rtiow, synth, synth360, tof, media, furnace, mis, mis_nohot, mixed.
"""
import ctypes as C
import os

from . import abi, capi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libwpt_scenes.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        capi.lib()      # libwpt_scenes.so links against libwurblpt_b200.so
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not found: run `make -C scenes`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.wpts_last_error.restype = C.c_char_p
        L.wpts_scene_create.argtypes = [C.c_char_p, C.c_float]
        L.wpts_scene_desc.restype = C.POINTER(abi.wpt_scene_desc)
        L.wpts_scene_desc.argtypes = [C.c_int]
        L.wpts_scene_destroy.argtypes = [C.c_int]
        L.wpts_default_view.argtypes = [C.c_char_p, C.POINTER(abi.wpt_camera_desc), C.POINTER(abi.wpt_sensor_desc),
                                        C.POINTER(abi.wpt_params), C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib = L
    return _lib


class BuiltScene:
    """A scene constructed through the API mirror; .desc is valid until close()."""

    def __init__(self, name, scale=1.0):
        self.name = name
        self.h = lib().wpts_scene_create(name.encode(), scale)
        if self.h < 0:
            raise RuntimeError(lib().wpts_last_error().decode())

    @property
    def desc(self):
        return lib().wpts_scene_desc(self.h).contents

    def close(self):
        if self.h >= 0:
            lib().wpts_scene_destroy(self.h)
            self.h = -1


class View:
    """Default camera / sensor / parameters of a named config."""

    def __init__(self, name):
        self.camera = abi.wpt_camera_desc()
        self.sensor = abi.wpt_sensor_desc()
        self.params = abi.wpt_params()
        s, t0, t1 = C.c_uint32(), C.c_float(), C.c_float()
        rc = lib().wpts_default_view(name.encode(), C.byref(self.camera), C.byref(self.sensor), C.byref(self.params),
                                     C.byref(s), C.byref(t0), C.byref(t1))
        if rc != 0:
            raise RuntimeError("no default view for %s" % name)
        self._keyframes = None
        if self.camera.n_keyframes:
            self._keyframes = (abi.wpt_keyframe * self.camera.n_keyframes)()
            C.memmove(self._keyframes, self.camera.keyframes, C.sizeof(self._keyframes))
            self.camera.keyframes = C.cast(self._keyframes, C.POINTER(abi.wpt_keyframe))
        self.samples_sqrt, self.t0, self.t1 = s.value, t0.value, t1.value
