"""Pins the CPU restatement (oracle/sva_oracle.c) against the reference's OWN sources, compiled from
/root/reference against oracle/cvshim into oracle/_ref/libsva_ref.so."""
import numpy as np
import pytest

from stereovisionarray_b200 import abi, synth


def _cams(width):
    return [abi.camera(*c) for c in synth.reference_cameras(width)]


def test_camera_project_inv_project(oracle, reference):
    rng = np.random.default_rng(1)
    cams = _cams(640)
    for _ in range(2000):
        c = cams[int(rng.integers(0, 25))]
        px = (int(rng.integers(-400, 400)), int(rng.integers(-300, 300)))
        ro, rr = oracle.camera_inv_project(c, px), reference.camera_inv_project(c, px)
        assert np.array_equal(ro, rr)  # bit-exact f64
        o = cams[int(rng.integers(0, 25))]
        for t in (0.5, 1.0, float(rng.uniform(0.3, 2.0))):
            p = [c.pos[i] + ro[i] * t for i in range(3)]
            assert oracle.camera_project(o, p) == reference.camera_project(o, p)


def test_bresenham(oracle, reference):
    rng = np.random.default_rng(2)
    cases = [((0, 0), (0, 0)), ((3, 3), (10, 3)), ((10, 3), (3, 3)), ((5, 1), (5, 9)), ((5, 9), (5, 1)), ((0, 0), (7, 7)), ((7, 7), (0, 0)), ((2, 9), (9, 2))]
    cases += [((int(a), int(b)), (int(c), int(d))) for a, b, c, d in rng.integers(-60, 60, size=(500, 4))]
    for a, b in cases:
        po, pr = oracle.bresenham(a, b), reference.bresenham(a, b)
        assert po.shape == pr.shape and np.array_equal(po, pr), (a, b)


def test_camera_pairs(oracle, reference):
    for t in range(10):
        assert np.array_equal(oracle.get_camera_pairs(25, t), reference.get_camera_pairs(25, t)), t
        for cam in range(25):
            assert np.array_equal(oracle.get_camera_pairs(25, t, cam), reference.get_camera_pairs(25, t, cam)), (t, cam)


def test_abs_diff(oracle, reference):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(90, 120), dtype=np.uint8)
    b = rng.integers(0, 256, size=(90, 120), dtype=np.uint8)
    for _ in range(50):
        w, h = int(rng.integers(1, 41)), int(rng.integers(1, 41))
        ax, ay, bx, by = (int(rng.integers(0, 120 - w)), int(rng.integers(0, 90 - h)), int(rng.integers(0, 120 - w)), int(rng.integers(0, 90 - h)))
        r = reference.abs_diff_roi(a, ax, ay, b, bx, by, w, h)
        o = oracle.abs_diff(a[ay:ay + h, ax:ax + w], b[by:by + h, bx:bx + w])
        assert r == o == float(np.abs(a[ay:ay + h, ax:ax + w].astype(int) - b[by:by + h, bx:bx + w].astype(int)).sum())
    # saturating 255/0 extremes: abs(A-B) must be absdiff, not a saturated subtract
    z, f = np.zeros((40, 40), np.uint8), np.full((40, 40), 255, np.uint8)
    assert reference.abs_diff_roi(z, 0, 0, f, 0, 0, 40, 40) == oracle.abs_diff(z, f) == 1600 * 255


def test_shift_perspective(oracle, reference):
    rng = np.random.default_rng(4)
    cams = _cams(160)
    img = rng.integers(0, 256, size=(120, 160), dtype=np.uint8)
    disp = rng.integers(0, 40, size=(120, 160), dtype=np.uint8)
    for a, b in [(12, 11), (11, 12), (12, 7), (12, 6), (12, 18), (3, 21)]:
        assert np.array_equal(oracle.shift_perspective_with_disparity(cams[a], cams[b], disp, img),
                              reference.shift_perspective_with_disparity(cams[a], cams[b], disp, img)), (a, b)


def _interior_mask(h, w, margin):
    m = synth.ellipse_mask(h, w)
    m[:margin] = 0; m[-margin:] = 0; m[:, :margin] = 0; m[:, -margin:] = 0
    return m


def test_improve_with_disparity(oracle, reference):
    rng = np.random.default_rng(5)
    h, w = 96, 128
    cams = _cams(w)
    center = synth.texture(h, w, 50)
    disp = rng.integers(5, 14, size=(h, w), dtype=np.uint8)
    mask = _interior_mask(h, w, 16)
    for pairs in ([(11, 12)], [(12, 11)], [(7, 12), (11, 12)], [(6, 12)], [(12, 12)]):
        images = [synth.texture(h, w, 60 + i) for i in range(len(pairs))]
        cp = [(cams[a], cams[b]) for a, b in pairs]
        rco, o = oracle.improve_with_disparity(disp, center, images, cp, mask, 21)
        rcr, r = reference.improve_with_disparity(disp, center, images, cp, mask, 21)
        assert rco == 0 and rcr == 0
        assert np.array_equal(o, r), pairs
    # ROI overrun: the reference throws (cv::Exception from Mat::operator()(Rect)); the oracle reports SVA_ERR_ROI
    full = np.full((h, w), 255, np.uint8)
    rco, _ = oracle.improve_with_disparity(disp, center, [center], [(cams[11], cams[12])], full, 21)
    rcr, _ = reference.improve_with_disparity(disp, center, [center], [(cams[11], cams[12])], full, 21)
    assert rco == abi.SVA_ERR_ROI and rcr != 0


@pytest.mark.parametrize("h,w,seed", [(120, 160, 7), (96, 200, 8)])
def test_reference_main_vs_oracle(oracle, reference, h, w, seed):
    """The reference driver end to end (main: resize x0.5 -> loop nest :49-95 -> depth :98-100 -> improveWithDisparity :114)
    against the oracle's literal mode on the same images."""
    sc = synth.make_literal_scene(h, w, seed)
    mask = _interior_mask(h, w, 26)
    up = [np.repeat(np.repeat(i, 2, axis=0), 2, axis=1) for i in sc["images"]]  # x0.5 block mean of a 2x replica is exact
    disp_r, depth_r, imp_r = reference.main_run(up, mask)
    cams = _cams(w)
    disp_o = oracle.match_literal(sc["images"], cams, [(12, 11)], mask, 20, 0.5, 1.0)
    assert disp_r.max() > 0
    assert np.array_equal(disp_o, disp_r)
    base = float(np.sqrt(sum((cams[12].pos[i] - cams[11].pos[i]) ** 2 for i in range(3))))
    depth_o = oracle.disparity_to_depth(disp_o, base, synth.REF_F, synth.REF_SENSOR / w)
    assert np.array_equal(depth_o, depth_r)  # inf == inf where disparity is 0
    rc, imp_o = oracle.improve_with_disparity(disp_o, sc["images"][12], [sc["images"][11]], [(cams[12], cams[11])], mask, 21)
    assert rc == 0 and np.array_equal(imp_o, imp_r)


def test_depth_consumers_f2(oracle, reference):
    """shiftPerspective2, Points3DToDepthMap, DepthMapToPoints3D (src/functions.cpp:79-146): bit-exact f64, same overwrite order"""
    cams = _cams(160)
    for seed, (h, w) in enumerate([(90, 160), (64, 75)]):
        depth = synth.make_depth_scene(h, w, seed)
        for a, b in [(12, 11), (12, 13), (12, 7), (12, 18), (6, 12), (0, 24)]:
            assert np.array_equal(oracle.shift_perspective2(cams[a], cams[b], depth), reference.shift_perspective2(cams[a], cams[b], depth)), (a, b)
        for c in (12, 6):
            po, pr = oracle.depth_map_to_points3d(depth, cams[c], w, h), reference.depth_map_to_points3d(depth, cams[c], w, h)
            assert po.shape == pr.shape and len(po) > 100 and np.array_equal(po, pr)
            for c2 in (12, 11, 17):  # re-project the cloud into other cameras: many points per pixel, the last one wins
                assert np.array_equal(oracle.points3d_to_depth_map(po, cams[c2], w, h), reference.points3d_to_depth_map(po, cams[c2], w, h))
            assert np.array_equal(oracle.points3d_to_depth_map(po, cams[c], w // 2, h // 2), reference.points3d_to_depth_map(po, cams[c], w // 2, h // 2))


def test_groups_f3(oracle, reference):
    go, gr = oracle.get_groups(25, "CHESS"), reference.get_groups(25, "CHESS")
    assert len(go) == len(gr) == 13
    for a, b in zip(go, gr):
        assert np.array_equal(a, b)
    assert oracle.get_groups(25, "OTHER") == [] and reference.get_groups(25, "OTHER") == []
