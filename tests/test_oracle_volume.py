"""Volume-mode oracle (frozen spec) — self-consistency and cross-checks against cv2 primitives.
These stages have no counterpart in the reference (parity unpinned by the reference)."""
import numpy as np
import pytest

from stereovisionarray_b200 import abi, synth

OFF8 = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


def _scene(h, w, D, offsets, seed, **kw):
    sc = synth.make_scene(h, w, D, offsets, seed)
    p = abi.make_params(w, h, D, offsets, **kw)
    return sc, p


def test_ad_volume_matches_numpy(oracle):
    sc, p = _scene(40, 56, 16, OFF8, 11, win_half=4)
    A = oracle.ad_volume(p, sc["ref"], sc["others"])
    h, w, D = A.shape
    exp = np.zeros((h, w, D), np.int64)
    R = sc["ref"].astype(np.int64)
    for d in range(D):
        for (gx, gy), img in zip(OFF8, sc["others"]):
            sh = np.zeros((h, w), np.int64)  # sh[y,x] = img[y - gy*d, x - gx*d], 0 outside
            ys, xs = np.mgrid[0:h, 0:w]
            sy, sx = ys - gy * d, xs - gx * d
            ok = (sy >= 0) & (sy < h) & (sx >= 0) & (sx < w)
            sh[ok] = img[sy[ok], sx[ok]]
            exp[:, :, d] += np.abs(R - sh)
    assert np.array_equal(A, exp)
    # pair-range partial sums add up (what the pair-sharded multi-GPU reduce relies on)
    parts = sum(oracle.ad_volume(p, sc["ref"], sc["others"], b, e).astype(np.int64) for b, e in [(0, 3), (3, 5), (5, 8)])
    assert np.array_equal(parts, exp)


def test_box_cost_vs_cv2_and_bruteforce(oracle):
    cv2 = pytest.importorskip("cv2")
    k = 5
    sc, p = _scene(48, 64, 16, [(-1, 0), (0, 1)], 12, win_half=k, cost_shift=3)
    A = oracle.ad_volume(p, sc["ref"], sc["others"])
    C16, C32 = oracle.box_cost(p, A, raw=True)
    h, w, D = A.shape
    for d in range(D):
        box = cv2.boxFilter(A[:, :, d].astype(np.float64), cv2.CV_64F, (2 * k, 2 * k), anchor=(k, k), normalize=False, borderType=cv2.BORDER_CONSTANT)
        valid = C32[:, :, d] != abi.SVA_COST_INVALID_U32
        assert valid.any()
        assert np.array_equal(C32[:, :, d][valid], box.astype(np.uint32)[valid])
        assert np.array_equal(C16[:, :, d][valid], np.minimum(p.cost_cap, box.astype(np.uint32)[valid] >> 3))
        assert (C16[:, :, d][~valid] == p.cost_cap).all()
    rng = np.random.default_rng(0)
    for _ in range(200):
        y, x, d = int(rng.integers(0, h)), int(rng.integers(0, w)), int(rng.integers(0, D))
        assert oracle.raw_cost_cell(p, sc["ref"], sc["others"], y, x, d) == int(C32[y, x, d])
        assert oracle.cell_valid(p, y, x, d) == (int(C32[y, x, d]) != abi.SVA_COST_INVALID_U32)


def test_sgm_properties(oracle):
    rng = np.random.default_rng(13)
    h, w, D = 20, 28, 16
    p = abi.make_params(w, h, D, [(-1, 0)], win_half=2, p1=0, p2=0, n_paths=8)
    Cv = rng.integers(0, 4096, size=(h, w, D), dtype=np.uint16)
    # P1 = P2 = 0  =>  L_r = C for every path  =>  S = n_paths * C
    assert np.array_equal(oracle.sgm_aggregate(p, Cv), (8 * Cv.astype(np.int64)).astype(np.uint16))
    p4 = abi.make_params(w, h, D, [(-1, 0)], win_half=2, p1=0, p2=0, n_paths=4)
    assert np.array_equal(oracle.sgm_aggregate(p4, Cv), 4 * Cv)
    # uniform cost => L_r = C (min term cancels), for any penalties
    pu = abi.make_params(w, h, D, [(-1, 0)], win_half=2, p1=7, p2=100, n_paths=8)
    U = np.full((h, w, D), 37, np.uint16)
    assert np.array_equal(oracle.sgm_aggregate(pu, U), 8 * U)
    # sum of the single paths == aggregate; bound L <= C + P2
    S = np.zeros((h, w, D), np.int64)
    for i in range(8):
        L = oracle.sgm_single_path(pu, Cv, i).astype(np.int64)
        assert (L <= Cv.astype(np.int64) + 100).all() and (L >= Cv).all() is not None
        S += L
    assert np.array_equal(S.astype(np.uint16), oracle.sgm_aggregate(pu, Cv))
    # mirror symmetry: path (+1,0) on the x-flipped volume == x-flip of path (-1,0)
    Lr = oracle.sgm_single_path(pu, Cv, 2)
    Ll = oracle.sgm_single_path(pu, np.ascontiguousarray(Cv[:, ::-1]), 3)
    assert np.array_equal(Lr, Ll[:, ::-1])


def test_sgm_single_path_python_reference(oracle):
    """independent pure-Python statement of the recurrence on a tiny volume (all 8 directions)"""
    rng = np.random.default_rng(14)
    h, w, D, P1, P2 = 6, 7, 8, 5, 40
    p = abi.make_params(w, h, D, [(-1, 0)], win_half=1, p1=P1, p2=P2, n_paths=8)
    Cv = rng.integers(0, 300, size=(h, w, D), dtype=np.uint16)
    dirs = [(0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (-1, 1), (1, -1), (-1, -1)]
    for di, (dx, dy) in enumerate(dirs):
        L = np.zeros((h, w, D), np.int64)
        ys = range(h) if dy >= 0 else range(h - 1, -1, -1)
        xs = range(w) if dx >= 0 else range(w - 1, -1, -1)
        for y in ys:
            for x in xs:
                px, py = x - dx, y - dy
                if not (0 <= px < w and 0 <= py < h):
                    L[y, x] = Cv[y, x]
                    continue
                q = L[py, px]
                m = q.min()
                for d in range(D):
                    c = [q[d], m + P2]
                    if d > 0: c.append(q[d - 1] + P1)
                    if d < D - 1: c.append(q[d + 1] + P1)
                    L[y, x, d] = Cv[y, x, d] + min(c) - m
        assert np.array_equal(oracle.sgm_single_path(p, Cv, di), L), (dx, dy)


def test_wta_lr_subpixel(oracle):
    rng = np.random.default_rng(15)
    h, w, D, k = 24, 40, 16, 3
    p = abi.make_params(w, h, D, [(-1, 0)], win_half=k, n_paths=4, lr_gx=-1, lr_max_diff=1, subpixel=1)
    S = rng.integers(100, 5000, size=(h, w, D), dtype=np.uint16)
    mask = synth.ellipse_mask(h, w)
    disp, sub = oracle.wta(p, S, mask)
    for y in range(h):
        for x in range(w):
            d = int(np.argmin(S[y, x]))  # numpy argmin = first minimum
            ok = k <= x < w - k and k <= y < h - k and mask[y, x] != 0 and oracle.cell_valid(p, y, x, d)
            if ok:
                xo = x + d  # lr_gx = -1: x' = x - gx*d
                if not (0 <= xo < w):
                    ok = False
                else:
                    cand = [(int(S[y, xo - dd, dd]), dd) for dd in range(D) if 0 <= xo - dd < w]
                    do = min(cand)[1]
                    ok = abs(d - do) <= 1
            if not ok:
                assert disp[y, x] == abi.SVA_DISP_INVALID and sub[y, x] == -1.0
                continue
            assert disp[y, x] == d
            e = np.float32(d)
            if 0 < d < D - 1:
                den = int(S[y, x, d - 1]) - 2 * int(S[y, x, d]) + int(S[y, x, d + 1])
                if den > 0:
                    e = np.float32(d) + np.float32(int(S[y, x, d - 1]) - int(S[y, x, d + 1])) / np.float32(2 * den)
            assert sub[y, x] == e
    assert (disp != abi.SVA_DISP_INVALID).sum() > 20


def test_pipeline_recovers_ground_truth(oracle):
    """end to end on a synthetic scene: most valid pixels land within 1 px of the known disparity"""
    offsets = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    sc, p = _scene(96, 128, 32, offsets, 16, win_half=4, n_paths=8, lr_gx=-1)
    disp, sub = oracle.depth_from_array(p, sc["ref"], sc["others"])
    ok = disp != abi.SVA_DISP_INVALID
    # judge only pixels whose ground-truth cell is a valid cell of the volume (all pair windows inside the image)
    gtv = np.array([[oracle.cell_valid(p, y, x, int(sc["gt"][y, x])) for x in range(128)] for y in range(96)])
    ok &= gtv
    assert ok.mean() > 0.15
    err = np.abs(disp[ok].astype(int) - sc["gt"][ok])
    assert (err <= 1).mean() > 0.9
    assert np.abs(sub[ok] - disp[ok]).max() <= 0.5 + 1e-6


def test_literal_equals_volume_wta_on_a_rectified_pair(oracle):
    """Where the literal epipolar segment stays on row y, the reference loop nest and the volume formulation pick the same winner:
    literal disparity == min_disp + argmin_d C_raw(y, x, d) restricted to the candidate range."""
    h, w = 96, 160
    sc = synth.make_literal_scene(h, w, 21)
    cams = [abi.camera(*c) for c in synth.reference_cameras(w)]
    k = 20
    lit = oracle.match_literal(sc["images"], cams, [(12, 11)], None, k)
    D = 32
    p = abi.make_params(w, h, D, [(-1, 0)], win_half=k, min_disp=0)
    A = oracle.ad_volume(p, sc["images"][12], [sc["images"][11]])
    _, C32 = oracle.box_cost(p, A, raw=True)
    hx, hy = w // 2, h // 2
    checked = 0
    for y in range(k, h - k):
        for x in range(k, w - k):
            ray = oracle.camera_inv_project(cams[12], (x - hx, y - hy))
            a = oracle.camera_project(cams[11], [cams[12].pos[i] + ray[i] * 0.5 for i in range(3)])
            b = oracle.camera_project(cams[11], [cams[12].pos[i] + ray[i] * 1.0 for i in range(3)])
            a, b = (a[0] + hx, a[1] + hy), (b[0] + hx, b[1] + hy)
            if a[1] != y or b[1] != y or not all(k <= v[0] <= w - k for v in (a, b)):
                continue
            lo, hi = min(a[0], b[0]) - x, max(a[0], b[0]) - x
            if lo < 0 or hi >= D:
                continue
            seg = C32[y, x, lo:hi + 1].astype(np.int64)
            assert lit[y, x] == lo + int(np.argmin(seg))
            checked += 1
    assert checked > 500


def test_sgm_rows_blocks_equal_the_whole_path(oracle):
    """orc_sgm_rows (the restatement of the multi-GPU row-block sweep): any cut of the rows into blocks, with L of the previous row carried
    across, adds up to exactly the whole-frame path — for every row-sweeping direction, blocks of one row included"""
    sc, p = _scene(29, 36, 16, [(-1, 0), (1, 1)], 23, win_half=3, n_paths=8)
    Cv = oracle.box_cost(p, oracle.ad_volume(p, sc["ref"], sc["others"]))
    rng = np.random.default_rng(4)
    for di in (0, 1, 4, 5, 6, 7):
        whole = oracle.sgm_single_path(p, Cv, di)
        cuts = sorted(set(rng.integers(1, p.height, 4).tolist()) | {1, p.height - 1})
        blocks = list(zip([0] + cuts, cuts + [p.height]))
        down = di in (0, 4, 5)
        S = np.zeros_like(whole)
        state = None
        for y0, y1 in (blocks if down else reversed(blocks)):
            out = np.zeros((p.width, p.num_disp), np.uint16)
            oracle.sgm_rows(p, Cv, di, y0, y1 - y0, state, S, out)
            state = out
        assert np.array_equal(S, whole), "direction %d" % di
    # a block that continues a sweep refuses to run without the state
    with pytest.raises(AssertionError):
        oracle.sgm_rows(p, Cv, 0, 5, 4, None, np.zeros_like(whole), None)


def _census_numpy(img):
    """independent restatement of the census signature: 9 x 7 window, bit = neighbour < centre, outside = 0, row-major without the centre"""
    h, w = img.shape
    pad = np.zeros((h + 6, w + 8), np.int32)
    pad[3:3 + h, 4:4 + w] = img
    c = img.astype(np.int32)
    t = np.zeros((h, w), np.uint64)
    bit = 0
    for dy in range(-3, 4):
        for dx in range(-4, 5):
            if dx == 0 and dy == 0:
                continue
            nb = pad[3 + dy:3 + dy + h, 4 + dx:4 + dx + w]
            t |= (nb < c).astype(np.uint64) << np.uint64(bit)
            bit += 1
    assert bit == 62
    return t


def test_census_cost_mode(oracle):
    """census mode (north_star names it, the reference has none: parity unpinned by the reference): signatures and the Hamming volume against
    independent numpy restatements; the rest of the pipeline is shared with SAD.  Unlike SAD, the cost ignores a gain / offset between cameras:
    with the other views darkened, census still recovers the ground truth where SAD does not."""
    h, w, D = 48, 72, 24
    offs = [(-1, 0), (1, 0), (0, -1), (1, 1)]
    sc = synth.make_scene(h, w, D, offs, 21)
    p = abi.make_params(w, h, D, offs, win_half=3, n_paths=8, lr_gx=-1, cost_mode=abi.COST_CENSUS)
    assert p.reserved[0] == abi.COST_CENSUS and p.cost_shift == 2  # 36 * 62 * 4 = 8928: two shifts bring the largest window sum under the cap of 4095
    tr = oracle.census_transform(sc["ref"])
    assert np.array_equal(tr, _census_numpy(sc["ref"]))
    assert int(tr.max()) < (1 << 62)
    A = oracle.ad_volume(p, sc["ref"], sc["others"])
    to = [_census_numpy(o) for o in sc["others"]]
    exp = np.zeros((h, w, D), np.int64)
    ys, xs = np.mgrid[0:h, 0:w]
    for d in range(D):
        for (gx, gy), t in zip(offs, to):
            sh = np.zeros((h, w), np.uint64)
            sy, sx = ys - gy * d, xs - gx * d
            ok = (sy >= 0) & (sy < h) & (sx >= 0) & (sx < w)
            sh[ok] = t[sy[ok], sx[ok]]
            x = tr ^ sh
            exp[:, :, d] += np.array([bin(int(v)).count("1") for v in x.ravel()]).reshape(h, w)
    assert np.array_equal(A.astype(np.int64), exp)
    assert int(A.max()) <= 62 * len(offs)
    # pair-range partials add up (what pair sharding relies on), as for SAD
    assert np.array_equal(oracle.ad_volume(p, sc["ref"], sc["others"], 0, 1) + oracle.ad_volume(p, sc["ref"], sc["others"], 1, 4), A)
    # robustness to a radiometric difference between the cameras: darkened other views cost SAD half of its correct pixels, census none
    h, w, D = 96, 144, 32
    sc = synth.make_scene(h, w, D, offs, 22)
    dark = [np.clip(o.astype(np.int32) * 6 // 10 + 9, 0, 255).astype(np.uint8) for o in sc["others"]]
    res = {}
    for name, views in (("clean", sc["others"]), ("dark", dark)):
        for mode in (abi.COST_SAD, abi.COST_CENSUS):
            pm = abi.make_params(w, h, D, offs, win_half=3, n_paths=8, lr_gx=0, cost_mode=mode)
            disp, _ = oracle.depth_from_array(pm, sc["ref"], views)
            valid = disp != abi.SVA_DISP_INVALID
            res[name, mode] = float((np.abs(disp[valid].astype(np.int32) - sc["gt"][valid]) <= 1).mean())
    assert abs(res["dark", abi.COST_CENSUS] - res["clean", abi.COST_CENSUS]) < 0.02, res
    assert res["dark", abi.COST_SAD] < res["clean", abi.COST_SAD] - 0.2, res
    assert res["clean", abi.COST_CENSUS] > res["clean", abi.COST_SAD] - 0.05, res
