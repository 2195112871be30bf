"""Multi-GPU entry points of the C ABI (csrc/sva_dist.cu) on real hardware: -m gpu.

  * one GPU: the peer-direct row-block pipeline with the ranks emulated by several contexts of one process (sva_rows_connect_local) — same
    kernels, same flags, same hand-off stores as across GPUs — and the NCCL-from-C path with a world of one;
  * two GPUs (skipped cleanly on a one-GPU box): tools/check_sharded.py under torchrun, every scheme against the ORACLE."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from stereovisionarray_b200 import abi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OFF8 = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


@pytest.mark.parametrize("h,w,D,G,offs,k", [(67, 96, 64, 3, OFF8, 4), (90, 140, 128, 4, OFF8, 5), (64, 70, 256, 2, OFF8, 6), (58, 3840, 192, 2, OFF8[:3], 4), (43, 200, 64, 8, OFF8, 3)])
def test_rows_direct_pipeline_with_local_ranks(oracle, h, w, D, G, offs, k):
    """G contexts of this process play the ranks: every rank's march stores its path-line state straight into the next context's link
    memory and the flag kernels sequence the hops; three frames go through the same links (the flags count frames), the first two without
    any host synchronisation in between.  The assembled maps equal the oracle's, bit for bit."""
    from stereovisionarray_b200.pipeline import DepthContext
    frames = [synth.make_scene(h, w, D, offs, 1200 + D + i, face=(i == 1)) for i in range(3)]
    p = abi.make_params(w, h, D, offs, win_half=k, n_paths=8, lr_gx=-1)
    ctxs = [DepthContext(0) for _ in range(G)]
    try:
        for r, c in enumerate(ctxs):
            c.upload(p, frames[0]["ref"], frames[0]["others"], frames[0]["mask"])
            c.rows_open(p, r, G)
        for r, c in enumerate(ctxs):
            c.rows_connect_local(ctxs[r - 1] if r > 0 else None, ctxs[r + 1] if r < G - 1 else None)
        blocks = [c.rows_block() for c in ctxs]
        assert blocks[0][0] == 0 and sum(n for _, n in blocks) == h and all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))
        # allocate every workspace before the first hand-off: one host thread enqueues all the ranks here, and a cudaFree (a device-wide
        # synchronisation) behind a wait kernel whose producer is not enqueued yet would sit out the time-out
        for c, (y0, n) in zip(ctxs, blocks):
            c.rows_begin(y0, n); c.run(abi.STAGE_AD); c.run(abi.STAGE_BOX); c.sgm_rows(2, y0, n); c.wta_rows(None, y0, n); c.synchronize()
        for i, f in enumerate(frames):
            for c in ctxs:
                c.upload(p, f["ref"], f["others"], f["mask"])
            order = list(range(G)) if i != 1 else list(reversed(range(G)))  # the enqueue order must not matter
            for r in order:
                ctxs[r].rows_run()
            if i == 0:
                for r in order:  # same frame again at once: acks keep a producer from overwriting a state that is still unread
                    ctxs[r].rows_run()
            parts = [c.rows_download() for c in ctxs]
            disp = np.concatenate([d for d, _ in parts]); sub = np.concatenate([s for _, s in parts])
            disp_o, sub_o = oracle.depth_from_array(p, f["ref"], f["others"], f["mask"])
            assert np.array_equal(disp, disp_o), "frame %d: integer disparity" % i
            assert np.array_equal(sub, sub_o), "frame %d: sub-pixel" % i
    finally:
        for c in ctxs:
            c.close()


def test_rows_direct_times_out_instead_of_hanging(monkeypatch):
    """a neighbour that never delivers: the wait kernel gives up after SVA_ROWS_TIMEOUT_MS and the download reports SVA_ERR_COMM"""
    from stereovisionarray_b200._lib import SvaError
    from stereovisionarray_b200.pipeline import DepthContext
    monkeypatch.setenv("SVA_ROWS_TIMEOUT_MS", "200")
    h, w, D = 48, 64, 64
    sc = synth.make_scene(h, w, D, OFF8, 5)
    p = abi.make_params(w, h, D, OFF8, win_half=3, n_paths=8, lr_gx=-1)
    a, b = DepthContext(0), DepthContext(0)
    try:
        for r, c in enumerate((a, b)):
            c.upload(p, sc["ref"], sc["others"])
            c.rows_open(p, r, 2)
        a.rows_connect_local(None, b); b.rows_connect_local(a, None)
        b.rows_run()  # rank 0 never runs: rank 1's down-sweep state never arrives
        with pytest.raises(SvaError) as ei:
            b.rows_download()
        assert ei.value.code == abi.SVA_ERR_COMM
    finally:
        a.close(); b.close()


def test_nccl_from_c_with_a_world_of_one(oracle):
    """sva_comm_* and the pair-sharded entry point through the library's own NCCL binding (dlopen): world = 1 reduces onto itself"""
    from stereovisionarray_b200.pipeline import DepthContext
    h, w, D = 60, 88, 32
    off15 = [(gx, gy) for gy in range(-1, 3) for gx in range(-1, 3) if (gx, gy) != (0, 0)]
    sc = synth.make_scene(h, w, D, off15, 9)
    p = abi.make_params(w, h, D, off15, win_half=4, n_paths=8, lr_gx=-1)
    c = DepthContext(0)
    try:
        c.comm_init(DepthContext.comm_unique_id(), 0, 1)
        c.comm_barrier()
        disp, sub = c.depth_pair_sharded(p, sc["ref"], sc["others"], None, 0)
        disp_o, sub_o = oracle.depth_from_array(p, sc["ref"], sc["others"])
        assert np.array_equal(disp, disp_o) and np.array_equal(sub, sub_o)
        c.comm_destroy()
    finally:
        c.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gpus_every_scheme_against_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if "vs oracle" in l]
    assert len(lines) >= 9 and all("bit-exact" in l for l in lines), r.stdout


@pytest.mark.parametrize("P", [2, 3, 4])
def test_rows_direct_frames_in_flight_on_one_stream(oracle, P):
    """several frames in flight without concurrent kernels: every rank has P contexts (own links, own volumes) on ONE stream.  P = 2: the host
    enqueues phase 0 of frame f before phase 1 of frame f - 1 (sva_rows_run_phase); P = 3: the three-part form (sva_rows_run_part) — part 0 of
    frame f, part 1 of frame f - 1, part 2 of frame f - 2; P = 4 = G + 1: the order skewed by chain position (iteration i of rank r: part 0 of
    frame i, part 1 of frame i - 1 - min(r, G-1-r), part 2 of frame i - 1 - max(r, G-1-r)) — every state a sweep needs was produced by the
    neighbour one iteration earlier.  G = 3 local ranks, five (P = 4: nine) frames through the pipeline."""
    from stereovisionarray_b200.pipeline import DepthContext
    h, w, D, G = 70, 120, 64, 3
    inputs = [synth.make_scene(h, w, D, OFF8, 1700 + i, face=(i == 1)) for i in range(P)]
    p = abi.make_params(w, h, D, OFF8, win_half=4, n_paths=8, lr_gx=-1)
    ctxs = [[DepthContext(0) for _ in range(P)] for _ in range(G)]
    try:
        for r in range(G):
            for j in range(1, P):
                ctxs[r][j].set_stream(ctxs[r][0].get_stream())
            for j in range(P):
                ctxs[r][j].upload(p, inputs[j]["ref"], inputs[j]["others"], inputs[j]["mask"])
                ctxs[r][j].rows_open(p, r, G)
        for r in range(G):
            for j in range(P):
                ctxs[r][j].rows_connect_local(ctxs[r - 1][j] if r > 0 else None, ctxs[r + 1][j] if r < G - 1 else None)
        for r in range(G):  # allocate every workspace before anything can wait on a neighbour
            for c in ctxs[r]:
                y0, n = c.rows_block()
                c.rows_begin(y0, n); c.run(abi.STAGE_AD); c.run(abi.STAGE_BOX); c.sgm_rows(2, y0, n); c.wta_rows(None, y0, n); c.synchronize()
        frames = 9 if P == 4 else 5
        for f in range(frames + P - 1):
            for r in range(G):
                if P == 4:
                    lag1, lag2 = 1 + min(r, G - 1 - r), 1 + max(r, G - 1 - r)
                    if f < frames:
                        ctxs[r][f % P].rows_run_part(0)
                    if 0 <= f - lag1 < frames:
                        ctxs[r][(f - lag1) % P].rows_run_part(1)
                    if 0 <= f - lag2 < frames:
                        ctxs[r][(f - lag2) % P].rows_run_part(2)
                elif P == 2:
                    if f < frames:
                        ctxs[r][f % P].rows_run_phase(0)
                    if f - P + 1 >= 0:
                        ctxs[r][(f - P + 1) % P].rows_run_phase(1)
                else:
                    if f < frames:
                        ctxs[r][f % P].rows_run_part(0)
                    if 0 <= f - 1 < frames:
                        ctxs[r][(f - 1) % P].rows_run_part(1)
                    if f - 2 >= 0:
                        ctxs[r][(f - 2) % P].rows_run_part(2)
        for j in range(P):
            parts = [ctxs[r][j].rows_download() for r in range(G)]
            disp = np.concatenate([d for d, _ in parts]); sub = np.concatenate([s for _, s in parts])
            disp_o, sub_o = oracle.depth_from_array(p, inputs[j]["ref"], inputs[j]["others"], inputs[j]["mask"])
            assert np.array_equal(disp, disp_o) and np.array_equal(sub, sub_o), "frame set %d" % j
        from stereovisionarray_b200._lib import SvaError
        with pytest.raises(SvaError, match="order 0, 1"):
            ctxs[0][0].rows_run_phase(1)
    finally:
        for row in ctxs:
            for c in reversed(row):  # the borrowers first: row[0] owns the stream they run on
                c.close()
