"""Literal mode + the stages either side of the matcher: CUDA (through the C ABI) vs the oracle and vs the fixtures the reference's
own code produced (tests/golden).  Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

from stereovisionarray_b200 import abi, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def api():
    from stereovisionarray_b200 import reference_api
    return reference_api


def _cams(api, w):
    return [api.Camera(f, pos, ps) for pos, f, ps in synth.reference_cameras(w)]


def _abi_cams(w):
    return [abi.camera(*c) for c in synth.reference_cameras(w)]


def test_reference_driver_fixtures(api):
    """bit-exact against what the reference's main() produced (disparity :49-95, depth :98-100, improveWithDisparity :114)"""
    for name in ("main_120x160_s7", "main_100x176_s9"):
        g = np.load(os.path.join(G, name + ".npz"))
        h, w, seed = int(g["h"]), int(g["w"]), int(g["seed"])
        sc = synth.make_literal_scene(h, w, seed)
        cams = _cams(api, w)
        disp = api.matchLiteral(sc["images"], cams, [(12, 11)], g["mask"], 20, 0.5, 1.0)
        assert np.array_equal(disp, g["disparity"])
        base = float(np.sqrt(sum((a - b) ** 2 for a, b in zip(cams[12].pos3D, cams[11].pos3D))))
        assert np.array_equal(api.disparityToDepth(disp, base, synth.REF_F, synth.REF_SENSOR / w), g["depth"])
        imp = api.improveWithDisparity(disp, sc["images"][12], [sc["images"][11]], [(cams[12], cams[11])], 21, g["mask"])
        assert np.array_equal(imp, g["improved"])


@pytest.mark.parametrize("pairs,mask_on,size", [([(12, 11)], True, (120, 160)), ([(12, 7)], False, (128, 120)), ([(12, 6), (12, 13)], True, (120, 168)),
                                                 ([(12, 18), (12, 11), (12, 17)], False, (112, 144)), ([(12, 11)], False, (240, 320))])
def test_match_literal_vs_oracle(api, oracle, pairs, mask_on, size):
    h, w = size
    sc = synth.make_literal_scene(h, w, 31 + len(pairs), pair=pairs[-1])
    mask = sc["mask"] if mask_on else None
    got = api.matchLiteral(sc["images"], _cams(api, w), pairs, mask, 20)
    exp = oracle.match_literal(sc["images"], _abi_cams(w), pairs, mask, 20, n_threads=8)
    assert exp.max() > 0
    assert np.array_equal(got, exp)


def test_warp_refine_fixtures(api):
    g = np.load(os.path.join(G, "warp_refine.npz"))
    cams = _cams(api, g["center"].shape[1])
    for (a, b), exp in zip(g["warp_pairs"], g["warps"]):
        assert np.array_equal(api.shiftPerspectiveWithDisparity(cams[a], cams[b], g["disp"], g["other"]), exp)
    for (a, b), exp in zip(g["imp_pairs"], g["imps"]):
        out = api.improveWithDisparity(np.clip(g["disp"], 5, 14), g["center"], [g["other"]], [(cams[a], cams[b])], 21, g["mask"])
        assert np.array_equal(out, exp)
    # the reference throws when a masked pixel's window leaves the image
    with pytest.raises(api.CvException):
        api.improveWithDisparity(g["disp"], g["center"], [g["other"]], [(cams[11], cams[12])], 21, np.full(g["center"].shape, 255, np.uint8))


def test_scalar_shims_and_abs_diff(api, oracle):
    g = np.load(os.path.join(G, "scalar_helpers.npz"))
    cams = _cams(api, 640)
    for c, p, ray, pr in zip(g["cam_idx"], g["px"], g["rays"], g["proj"]):
        r = api.Camera.inv_project(cams[c[0]], (int(p[0]), int(p[1])))
        assert np.array_equal(np.array(r), ray)
        for j, t in enumerate((0.5, 1.0)):
            assert cams[c[1]].project([cams[c[0]].pos3D[i] + r[i] * t for i in range(3)]) == tuple(pr[j])
    off = 0
    for e, n in zip(g["ends"], g["line_len"]):
        pts = api.bresenham((int(e[0]), int(e[1])), (int(e[2]), int(e[3])))
        assert np.array_equal(np.array(pts, np.int32).reshape(-1, 2), g["line_pts"][off:off + n])
        off += n
    for t in range(10):
        assert np.array_equal(np.array(api.getCameraPairs(cams, t), np.int32).reshape(-1, 2), g["t%d" % t])
    for cnum in range(25):
        assert np.array_equal(np.array(api.getCameraPairs(cams, 5, cnum), np.int32).reshape(-1, 2), g["t5_c%d" % cnum])
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (300, 400), dtype=np.uint8); b = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    for (y0, x0, hh, ww) in [(0, 0, 40, 40), (10, 17, 1, 1), (5, 3, 20, 20), (0, 0, 300, 400), (100, 101, 37, 255)]:
        ra, rb = a[y0:y0 + hh, x0:x0 + ww], b[y0 + 0:y0 + hh, x0:x0 + ww]
        assert api.getAbsDiff(ra, rb) == oracle.abs_diff(ra, rb)
    assert api.getAbsDiff(np.zeros((40, 40), np.uint8), np.full((40, 40), 255, np.uint8)) == 1600 * 255


def test_depth_consumers_f2_f3(api, oracle):
    """shiftPerspective2 / Points3DToDepthMap / DepthMapToPoints3D / getGroups on the GPU path: bit-exact f64 against the reference's
    fixtures (tests/golden/depth_consumers.npz) and against the oracle on a larger scene with heavy overwriting"""
    g = np.load(os.path.join(G, "depth_consumers.npz"))
    depth = g["depth"]
    h, w = depth.shape
    cams = _cams(api, w)
    for (a, b), exp in zip(g["sp2_pairs"], g["sp2"]):
        assert np.array_equal(api.shiftPerspective2(cams[a], cams[b], depth), exp)
    cloud = api.DepthMapToPoints3D(depth, cams[12], (w, h))
    assert np.array_equal(cloud, g["cloud"])
    for c, exp in zip(g["maps_cams"], g["maps"]):
        assert np.array_equal(api.Points3DToDepthMap(cloud, cams[c], (w, h)), exp)
    assert np.array_equal(api.Points3DToDepthMap(cloud, cams[12], (w // 2, h // 2)), g["half_map"])
    groups = api.getGroups(cams, "CHESS")
    assert [len(x) for x in groups] == list(g["group_sizes"]) and np.array_equal(np.array([p for x in groups for p in x], np.int32), g["group_pairs"])
    assert api.getGroups(cams, "OTHER") == []
    # larger, against the oracle: 640x480 depth, clouds re-projected at a quarter of the resolution (16 points per pixel compete)
    h, w = 480, 640
    depth = synth.make_depth_scene(h, w, 9)
    cams, ocams = _cams(api, w), _abi_cams(w)
    for a, b in [(12, 11), (12, 8), (2, 22)]:
        assert np.array_equal(api.shiftPerspective2(cams[a], cams[b], depth), oracle.shift_perspective2(ocams[a], ocams[b], depth))
    cloud = api.DepthMapToPoints3D(depth, cams[7], (w, h))
    assert np.array_equal(cloud, oracle.depth_map_to_points3d(depth, ocams[7], w, h)) and len(cloud) > 250000
    for c, res in [(7, (w, h)), (12, (w, h)), (7, (w // 4, h // 4)), (17, (w // 2, h // 2))]:
        assert np.array_equal(api.Points3DToDepthMap(cloud, cams[c], res), oracle.points3d_to_depth_map(cloud, ocams[c], res[0], res[1]))
    assert np.array_equal(api.Points3DToDepthMap(np.zeros((0, 3)), cams[7], (w, h)), np.zeros((h, w)))


def test_calculate_average_error(api):
    """cv::mean(image, mask)[0] (src/functions.cpp:348-354): the masked mean of an f64 map, against numpy and cv2"""
    import cv2
    rng = np.random.default_rng(8)
    img = rng.random((480, 640)) * 2.0 - 0.5
    mask = synth.ellipse_mask(480, 640)
    got = api.calculateAverageError(img, mask)
    assert abs(got - float(img[mask != 0].mean())) <= 1e-13 and abs(got - cv2.mean(img, mask)[0]) <= 1e-13
    assert api.calculateAverageError(img, np.zeros_like(mask)) == 0.0 == cv2.mean(img, np.zeros_like(mask))[0]
