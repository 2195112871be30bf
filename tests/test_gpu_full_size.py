"""Parity at BASELINE.json's FULL sizes (needs a B200: -m gpu): the oracle's OpenMP pipeline finishes a whole c0 / c1 / c2 / c4 frame in
seconds, so the integer disparity map and the sub-pixel map of the one-call path are compared bit for bit on the full frame; c3 (2.1 G
cells, 15 pairs: about a minute and 13 GB of host memory for the oracle) is compared at its real size too — that is the only size at which
the ranged SGM launches and the D = 256 block layout at 3840 columns run — and at a quarter of its resolution.  Plus two size-independent properties
of the pipeline at c1 size: zero penalties make the aggregation a multiple of the cost volume, and the AD volume is additive over
camera pairs."""
import numpy as np
import pytest

from stereovisionarray_b200 import abi, configs, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from stereovisionarray_b200.pipeline import DepthContext
    c = DepthContext(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["c0", "c1", "c2", "c4"])
def test_full_frame_bit_exact(ctx, oracle, name):
    p = configs.params(name)
    sc = configs.scene(name)
    disp, sub = ctx.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    disp_o, sub_o = oracle.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    assert np.array_equal(disp, disp_o), "integer disparity differs on the full %s frame" % name
    assert float(np.max(np.abs(sub - sub_o))) <= 0.05 and np.array_equal(sub, sub_o)
    valid = disp != abi.SVA_DISP_INVALID
    assert valid.mean() > 0.05
    # the synthetic scene has known ground truth: most accepted pixels are within one disparity of it
    gt = sc["gt"].astype(np.int32)
    assert (np.abs(disp[valid].astype(np.int32) - gt[valid]) <= 1).mean() > 0.8


def test_c3_pairs_and_range_at_quarter_resolution(ctx, oracle):
    c = configs.CONFIGS["c3"]
    off = configs.offsets("c3")
    h, w, D = c["height"] // 4, c["width"] // 4, c["num_disp"]
    sc = synth.make_scene(h, w, D, off, 3000)
    p = abi.make_params(w, h, D, off, win_half=20, n_paths=8, lr_gx=-1, subpixel=1)
    disp, sub = ctx.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    disp_o, sub_o = oracle.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    assert np.array_equal(disp, disp_o) and np.array_equal(sub, sub_o)


def test_c3_full_resolution_bit_exact(ctx, oracle):
    """3840x2160, D = 256, 15 pairs, 8 paths on ONE GPU against a live oracle run; the oracle's maps must also reproduce the committed digests
    (tests/golden/c3_full_oracle.json, made by tests/golden/make_c3_hash.py) that the multi-GPU checks compare against."""
    import json
    import os
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_c3_hash", os.path.join(os.path.dirname(__file__), "golden", "make_c3_hash.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    summarize = mod.summarize
    p = configs.params("c3")
    sc = configs.scene("c3")
    disp, sub = ctx.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    disp_o, sub_o = oracle.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    assert np.array_equal(disp, disp_o), "integer disparity differs on the full c3 frame"
    assert float(np.max(np.abs(sub - sub_o))) <= 0.05  # north_star's tolerance
    assert np.array_equal(sub, sub_o)                  # in practice bit-equal
    valid = disp != abi.SVA_DISP_INVALID
    assert valid.mean() > 0.05
    assert (np.abs(disp[valid].astype(np.int32) - sc["gt"][valid]) <= 1).mean() > 0.8
    with open(os.path.join(os.path.dirname(__file__), "golden", "c3_full_oracle.json")) as f:
        gold = json.load(f)
    assert summarize(sc, disp_o, sub_o) == gold


def test_zero_penalties_and_pair_linearity_at_c1_size(ctx):
    p = configs.params("c1")
    sc = configs.scene("c1")
    # P1 = P2 = 0: every path cost is L = C + min_k L_prev - min_k L_prev = C, so S = n_paths * C exactly
    p0 = configs.params("c1", p1=0, p2=0)
    ctx.set_debug(1, 0)
    ctx.upload(p0, sc["ref"], sc["others"], sc["mask"])
    ctx.run(abi.STAGE_ALL)
    C = ctx.download_cost()
    S = ctx.download_sgm()
    assert np.array_equal(S, C * np.uint16(p0.n_paths))
    ctx.set_debug(0, 0)
    del C, S
    # the AD volume is a sum over camera pairs: partials over disjoint pair ranges add up to the full volume (what pair sharding relies on)
    ctx.upload(p, sc["ref"], sc["others"], sc["mask"])
    ctx.run(abi.STAGE_AD)
    full = ctx.download_ad()
    acc = np.zeros_like(full)
    for b, e in [(0, 3), (3, 4), (4, 8)]:
        ctx.set_pair_range(b, e)
        ctx.run(abi.STAGE_AD)
        acc += ctx.download_ad()
    assert np.array_equal(acc, full)
    assert int(full.max()) <= 255 * p.n_pairs
