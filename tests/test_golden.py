"""Oracle (CPU) against the committed fixtures that the reference's own sources produced (tests/golden/make_golden.py).
Runs without /root/reference, so it also guards the oracle on the GPU box."""
import os

import numpy as np

from stereovisionarray_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_scalar_helpers(oracle):
    g = np.load(os.path.join(G, "scalar_helpers.npz"))
    cams = [abi.camera(*c) for c in synth.reference_cameras(640)]
    for c, p, ray, pr in zip(g["cam_idx"], g["px"], g["rays"], g["proj"]):
        r = oracle.camera_inv_project(cams[c[0]], (int(p[0]), int(p[1])))
        assert np.array_equal(r, ray)
        for j, t in enumerate((0.5, 1.0)):
            assert oracle.camera_project(cams[c[1]], [cams[c[0]].pos[i] + r[i] * t for i in range(3)]) == tuple(pr[j])
    off = 0
    for e, n in zip(g["ends"], g["line_len"]):
        pts = oracle.bresenham((int(e[0]), int(e[1])), (int(e[2]), int(e[3])))
        assert np.array_equal(pts, g["line_pts"][off:off + n])
        off += n
    for t in range(10):
        assert np.array_equal(oracle.get_camera_pairs(25, t), g["t%d" % t])
    for c in range(25):
        assert np.array_equal(oracle.get_camera_pairs(25, 5, c), g["t5_c%d" % c])


def test_reference_driver_fixtures(oracle):
    for name in ("main_120x160_s7", "main_100x176_s9"):
        g = np.load(os.path.join(G, name + ".npz"))
        h, w, seed = int(g["h"]), int(g["w"]), int(g["seed"])
        sc = synth.make_literal_scene(h, w, seed)
        assert np.array_equal(sc["images"][12], g["img12"]) and np.array_equal(sc["images"][11], g["img11"])  # synth is deterministic
        cams = [abi.camera(*c) for c in synth.reference_cameras(w)]
        disp = oracle.match_literal(sc["images"], cams, [(12, 11)], g["mask"], 20, 0.5, 1.0)
        assert np.array_equal(disp, g["disparity"])
        base = float(np.sqrt(sum((cams[12].pos[i] - cams[11].pos[i]) ** 2 for i in range(3))))
        assert np.array_equal(oracle.disparity_to_depth(disp, base, synth.REF_F, synth.REF_SENSOR / w), g["depth"])
        rc, imp = oracle.improve_with_disparity(disp, sc["images"][12], [sc["images"][11]], [(cams[12], cams[11])], g["mask"], 21)
        assert rc == 0 and np.array_equal(imp, g["improved"])


def test_warp_refine_fixtures(oracle):
    g = np.load(os.path.join(G, "warp_refine.npz"))
    cams = [abi.camera(*c) for c in synth.reference_cameras(g["center"].shape[1])]
    for (a, b), exp in zip(g["warp_pairs"], g["warps"]):
        assert np.array_equal(oracle.shift_perspective_with_disparity(cams[a], cams[b], g["disp"], g["other"]), exp)
    for (a, b), exp in zip(g["imp_pairs"], g["imps"]):
        rc, out = oracle.improve_with_disparity(np.clip(g["disp"], 5, 14), g["center"], [g["other"]], [(cams[a], cams[b])], g["mask"], 21)
        assert rc == 0 and np.array_equal(out, exp)


def test_depth_consumer_fixtures(oracle):
    """f2 / f3: shiftPerspective2, DepthMapToPoints3D, Points3DToDepthMap, getGroups against what the reference's code produced"""
    g = np.load(os.path.join(G, "depth_consumers.npz"))
    depth = g["depth"]
    h, w = depth.shape
    assert np.array_equal(depth, synth.make_depth_scene(h, w, 5))  # synth is deterministic
    cams = [abi.camera(*c) for c in synth.reference_cameras(w)]
    for (a, b), exp in zip(g["sp2_pairs"], g["sp2"]):
        assert np.array_equal(oracle.shift_perspective2(cams[a], cams[b], depth), exp)
    cloud = oracle.depth_map_to_points3d(depth, cams[12], w, h)
    assert np.array_equal(cloud, g["cloud"])
    for c, exp in zip(g["maps_cams"], g["maps"]):
        assert np.array_equal(oracle.points3d_to_depth_map(cloud, cams[c], w, h), exp)
    assert np.array_equal(oracle.points3d_to_depth_map(cloud, cams[12], w // 2, h // 2), g["half_map"])
    groups = oracle.get_groups(25, "CHESS")
    assert [len(x) for x in groups] == list(g["group_sizes"]) and np.array_equal(np.concatenate(groups), g["group_pairs"])


def test_c3_digest_inputs_are_what_the_generator_makes():
    """tests/golden/c3_full_oracle.json holds digests of the ORACLE's maps of the full c3 frame (made by tests/golden/make_c3_hash.py; the
    GPU tests re-derive them from a live oracle run).  Here, on CPU: the synthetic generator still produces the inputs those digests belong to."""
    import hashlib
    import json
    from stereovisionarray_b200 import configs
    with open(os.path.join(ROOT, "tests", "golden", "c3_full_oracle.json")) as f:
        gold = json.load(f)
    sc = configs.scene("c3")
    assert hashlib.sha256(np.ascontiguousarray(np.stack([sc["ref"]] + list(sc["others"]))).tobytes()).hexdigest() == gold["inputs_sha256"]
    assert gold["valid_pixels"] > 0.8 * 3840 * 2160 and len(gold["disp_sha256"]) == 64
