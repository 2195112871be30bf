"""The C-ABI library loads and exports every symbol include/sva_c_api.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sva_c_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sva_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stereovisionarray_b200 import build
    build.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "stereovisionarray_b200", "libsva_b200.so"))
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from stereovisionarray_b200._lib import EXPORTS
    assert sorted(EXPORTS) == names, (sorted(set(names) - set(EXPORTS)), sorted(set(EXPORTS) - set(names)))
    assert lib.sva_api_version() == 1


def test_struct_layouts_match_the_header():
    from stereovisionarray_b200 import abi
    assert ctypes.sizeof(abi.SvaCamera) == 40            # pos3D (24) + f (8) + pixel_size (8) — include/Camera.h member order
    assert ctypes.sizeof(abi.SvaImageU8) == 24
    assert ctypes.sizeof(abi.SvaParams) == 4 * (6 + 64 + 8 + 8)
    assert abi.MID_LEFT == 8 and abi.TO_CENTER_SMALL == 7 and abi.CROSS == 5  # enum pairType order, include/functions.h:8-19


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from stereovisionarray_b200.pipeline import DepthContext
    with pytest.raises(RuntimeError):
        DepthContext(0)


def test_host_shims_without_gpu():
    """scalar helpers are host-side and need no device"""
    import numpy as np
    from stereovisionarray_b200 import reference_api as api, synth
    g = np.load(os.path.join(ROOT, "tests", "golden", "scalar_helpers.npz"))
    cams = [api.Camera(f, pos, ps) for pos, f, ps in synth.reference_cameras(640)]
    for c, p, ray, pr in list(zip(g["cam_idx"], g["px"], g["rays"], g["proj"]))[:64]:
        r = cams[c[0]].inv_project((int(p[0]), int(p[1])))
        assert np.array_equal(np.array(r), ray)
        assert cams[c[1]].project([cams[c[0]].pos3D[i] + r[i] * 0.5 for i in range(3)]) == tuple(pr[0])
    off = 0
    for e, n in zip(g["ends"], g["line_len"]):
        assert np.array_equal(np.array(api.bresenham((int(e[0]), int(e[1])), (int(e[2]), int(e[3]))), np.int32).reshape(-1, 2), g["line_pts"][off:off + n])
        off += n
    for t in range(10):
        assert np.array_equal(np.array(api.getCameraPairs(cams, t), np.int32).reshape(-1, 2), g["t%d" % t])
    pairs, offs = api.gridPairs(3, 3, 4, api.TO_CENTER_SMALL)
    assert offs == [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)] and pairs[0] == [4, 0]
    pairs, offs = api.gridPairs(4, 4, 5, api.TO_CENTER)
    assert len(pairs) == 15 and (2, 2) in offs and (-1, -1) in offs
