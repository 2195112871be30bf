"""Drop-in check: adapter/sva_functions.cpp (the reference-side binding, compiled against the reference's OWN headers) driven like
the reference's main() — getCameraPairs / Camera / improveWithDisparity / getAbsDiff under their reference names, the loop nest as
one svaMatchLiteral call — must reproduce what the reference's main() produced (tests/golden/main_*.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

from stereovisionarray_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "adapter", "_build", "libsva_adapter_test.so")


def test_adapter_reproduces_reference_driver():
    if not os.path.exists(SO):
        pytest.skip("adapter test library not built (needs /root/reference at build time)")
    lib = C.CDLL(SO)
    lib.adapter_last_error.restype = C.c_char_p
    for name in ("main_120x160_s7", "main_100x176_s9"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        h, w, seed = int(g["h"]), int(g["w"]), int(g["seed"])
        sc = synth.make_literal_scene(h, w, seed)
        imgs = [np.ascontiguousarray(i) for i in sc["images"]]
        ptrs = (C.c_void_p * 25)(*[i.ctypes.data for i in imgs])
        mask = np.ascontiguousarray(g["mask"])
        disp = np.zeros((h, w), np.uint8); imp = np.zeros((h, w), np.uint8); sad = C.c_double()
        rc = lib.adapter_driver(ptrs, w, h, mask.ctypes.data_as(C.c_void_p), disp.ctypes.data_as(C.c_void_p), imp.ctypes.data_as(C.c_void_p), C.byref(sad))
        assert rc == 0, lib.adapter_last_error()
        assert np.array_equal(disp, g["disparity"])
        assert np.array_equal(imp, g["improved"])
        assert sad.value == float(np.abs(imgs[12][10:50, 10:50].astype(int) - imgs[11][11:51, 12:52].astype(int)).sum())


def test_adapter_depth_consumers():
    """shiftPerspective2 / DepthMapToPoints3D / Points3DToDepthMap / getGroups through the C++ binding == the reference's fixtures"""
    if not os.path.exists(SO):
        pytest.skip("adapter test library not built (needs /root/reference at build time)")
    lib = C.CDLL(SO)
    lib.adapter_last_error.restype = C.c_char_p
    lib.adapter_depth_consumers.restype = C.c_longlong
    g = np.load(os.path.join(ROOT, "tests", "golden", "depth_consumers.npz"))
    depth = np.ascontiguousarray(g["depth"])
    h, w = depth.shape
    shifted = np.zeros((h, w)); cloud = np.zeros((h * w, 3)); dmap = np.zeros((h, w)); sizes = np.zeros(16, np.int32)
    n = lib.adapter_depth_consumers(depth.ctypes.data_as(C.c_void_p), w, h, 12, 11, shifted.ctypes.data_as(C.c_void_p), cloud.ctypes.data_as(C.c_void_p),
                                    dmap.ctypes.data_as(C.c_void_p), sizes.ctypes.data_as(C.c_void_p))
    assert n >= 0, lib.adapter_last_error()
    assert np.array_equal(shifted, g["sp2"][0])          # pair (12, 11) is the first fixture pair
    assert np.array_equal(cloud[:n], g["cloud"])
    assert np.array_equal(dmap, g["maps"][1])            # the cloud of camera 12 seen from camera 11
    assert list(sizes[:13]) == list(g["group_sizes"])
