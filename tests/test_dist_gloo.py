"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups.  The compute backend here is the ORACLE (tests may use it);
what is under test is the sharding + packed-int32 reduce of stereovisionarray_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereovisionarray_b200 import abi, dist as sdist, synth

OFF15 = [(gx, gy) for gy in range(-1, 3) for gx in range(-1, 3) if (gx, gy) != (0, 0)]


def test_partitions():
    for n, w in [(64, 1), (64, 2), (64, 4), (64, 8), (10, 4), (3, 8), (15, 8)]:
        rs = [sdist.frame_range(n, w, r) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1
    assert [e - b for b, e in sdist.pair_ranges(15, 8)] == [2, 2, 2, 2, 2, 2, 2, 1]
    assert sdist.pair_ranges(3, 8)[3:] == [(3, 3)] * 5
    assert sdist.packed_no_carry(15) and sdist.packed_no_carry(32) and not sdist.packed_no_carry(129)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        orc = Oracle()
        h, w, D = 40, 56, 16
        sc = synth.make_scene(h, w, D, OFF15, 77)  # same seed on every rank: the frame is replicated
        p = abi.make_params(w, h, D, OFF15, win_half=3, n_paths=4, lr_gx=-1)
        b, e = sdist.pair_ranges(p.n_pairs, world)[rank]
        part = orc.ad_volume(p, sc["ref"], sc["others"], b, e) if e > b else np.zeros((h, w, D), np.uint16)
        t = torch.from_numpy(sdist.numpy_pack(part).copy())
        sdist.reduce_packed_u16(t, 0)
        # frames: every rank processes its own frames, nothing is exchanged; gather only the checksums for the test
        fb, fe = sdist.frame_range(5, world, rank)
        sums = []
        for f in range(fb, fe):
            scf = synth.make_scene(32, 40, 16, [(-1, 0)], 4000 + f)
            pf = abi.make_params(40, 32, 16, [(-1, 0)], win_half=2, n_paths=4, lr_gx=-1)
            d, _ = orc.depth_from_array(pf, scf["ref"], scf["others"])
            sums.append((f, int(d.astype(np.int64).sum())))
        gathered = [None] * world
        dist.all_gather_object(gathered, sums)
        if rank == 0:
            full = orc.ad_volume(p, sc["ref"], sc["others"])
            A = t.numpy().view(np.uint16).reshape(h, w, D)
            ok_reduce = bool(np.array_equal(A, full))
            disp, sub = orc.wta(p, orc.sgm_aggregate(p, orc.box_cost(p, A)))
            disp1, sub1 = orc.depth_from_array(p, sc["ref"], sc["others"])
            ok_pipe = bool(np.array_equal(disp, disp1) and np.array_equal(sub, sub1))
            flat = sorted(x for g in gathered for x in g)
            exp = []
            for f in range(5):
                scf = synth.make_scene(32, 40, 16, [(-1, 0)], 4000 + f)
                pf = abi.make_params(40, 32, 16, [(-1, 0)], win_half=2, n_paths=4, lr_gx=-1)
                exp.append((f, int(orc.depth_from_array(pf, scf["ref"], scf["others"])[0].astype(np.int64).sum())))
            q.put((ok_reduce, ok_pipe, flat == exp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pair_sharded_reduce_and_frame_partition(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok_reduce, ok_pipe, ok_frames = q.get(timeout=10)
    assert ok_reduce, "packed int32 reduce of the AD partials != full AD volume"
    assert ok_pipe, "pipeline from the reduced volume != single-process pipeline"
    assert ok_frames, "frame-partitioned results != serial results"


def test_slice_direction_row_plans():
    """host-side plans of dist.slice_sharded_depth (pure Python; the device path is covered by tests/test_gpu_volume.py and, over real
    GPUs, by tools/check_sharded.py under torchrun)"""
    from stereovisionarray_b200 import abi, dist as sdist
    for world in (1, 2, 4, 8):
        masks = sdist.direction_masks(8, world)
        assert len(masks) == world and sum(masks) == 0xFF and all(a & b == 0 for i, a in enumerate(masks) for b in masks[i + 1:])
    assert sdist.direction_masks(4, 8) == [1, 2, 4, 8, 0, 0, 0, 0]
    rows_per, blocks = sdist.row_blocks(2160, 8)
    assert rows_per == 270 and blocks[0] == (0, 270) and blocks[-1] == (1890, 2160)
    rows_per, blocks = sdist.row_blocks(70, 4)
    assert rows_per == 18 and blocks == [(0, 18), (18, 36), (36, 54), (54, 70)]
    assert sdist.row_blocks(10, 8)[1][-1] == (10, 10)  # empty trailing blocks
    p = abi.make_params(64, 48, 256, [(-1, 0), (1, 1)], win_half=4, min_disp=3)
    ps = sdist.slice_params(p, 5, 8)
    assert (ps.num_disp, ps.min_disp, ps.n_pairs, ps.cost_shift, ps.p2) == (32, 3 + 5 * 32, 2, p.cost_shift, p.p2)
    assert (p.num_disp, p.min_disp) == (256, 3)  # the original is untouched
    import pytest
    with pytest.raises(ValueError):
        sdist.slice_params(p, 0, 3)


class _ToyRowsBackend:
    """Stands in for DepthContext in the row-block pipeline: the same calls (rows_begin / run / sgm_rows / wta_rows, state buffers by
    address), a toy recurrence instead of SGM.  A sweep's state after row y is L(y) = (L(y-1) * (slot + 2) + c(y)) mod 251 per element,
    and S(y) accumulates every L(y): the result depends on every hop delivering the right predecessor's state, in the right order."""

    def __init__(self, p):
        self.p = p
        self.n = 3 * p.width * p.num_disp
        self.S = np.zeros(p.height, np.int64)
        self.log = []

    def _buf(self, ptr):
        import ctypes
        return np.ctypeslib.as_array((ctypes.c_int16 * self.n).from_address(ptr))

    def rows_begin(self, y0, rows):
        self.S[y0:y0 + rows] = 0

    def run(self, stage):
        self.log.append(stage)

    def sgm_rows(self, group, y0, rows, state_in=0, state_out=0):
        if group == 2:
            self.S[y0:y0 + rows] += 7
            return
        mult = np.repeat(np.arange(3) + 2, self.n // 3).astype(np.int64)
        first = (y0 == 0) if group == 0 else (y0 + rows == self.p.height)
        L = np.zeros(self.n, np.int64) if first else self._buf(state_in).astype(np.int64)
        ys = range(y0, y0 + rows) if group == 0 else range(y0 + rows - 1, y0 - 1, -1)
        for y in ys:
            L = (L * mult + (y * 13 + group * 5 + 1)) % 251
            self.S[y] += int(L.sum())
        if state_out:
            self._buf(state_out)[:] = L.astype(np.int16)

    def wta_rows(self, ptr, y0, rows):
        self.log.append(("wta", y0, rows))


def _rows_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = abi.make_params(12, 23, 8, [(-1, 0)], win_half=2)
        be = _ToyRowsBackend(p)
        keep = {}
        for _ in range(2):  # twice: the state buffers are reused across frames
            y0, y1 = sdist.row_sharded_compute(be, p, rank, world, None, keep, device="cpu")
        mine = torch.zeros(p.height, dtype=torch.int64)
        mine[y0:y1] = torch.from_numpy(be.S[y0:y1])
        dist.all_reduce(mine)
        if rank == 0:
            one = _ToyRowsBackend(p)
            one.rows_begin(0, p.height)
            one.sgm_rows(2, 0, p.height)
            one.sgm_rows(0, 0, p.height)
            one.sgm_rows(1, 0, p.height)
            q.put(bool(np.array_equal(mine.numpy(), one.S)) and be.log[-1] == ("wta", y0, y1 - y0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_row_block_pipeline_hops(world):
    """dist.row_sharded_compute over gloo: every block continues the sweeps from the state its neighbour handed over (down 0 -> G-1,
    up G-1 -> 0, both in flight at once), and the assembled result equals the single-process run"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10)


def test_row_pipeline_schedule():
    for world in (1, 2, 3, 4, 8):
        plans = [sdist.row_pipeline_steps(r, world) for r in range(world)]
        assert all(len(pl) == max(0, world - 1) for pl in plans)
        for s in range(world - 1):
            sends = [(r, plans[r][s][0]) for r in range(world) if plans[r][s][0] is not None] + [(r, plans[r][s][2]) for r in range(world) if plans[r][s][2] is not None]
            recvs = [(plans[r][s][1], r) for r in range(world) if plans[r][s][1] is not None] + [(plans[r][s][3], r) for r in range(world) if plans[r][s][3] is not None]
            assert sorted(sends) == sorted(recvs) == sorted([(s, s + 1), (world - 1 - s, world - 2 - s)])
    with pytest.raises(ValueError):
        sdist.row_sharded_compute(None, abi.make_params(16, 10, 8, [(-1, 0)], win_half=2), 0, 16)


class _OracleRowsBackend:
    """DepthContext's row-block calls served by the CPU oracle (tests may use it): real SGM through dist.row_sharded_compute.  The state
    crossing a block boundary is the oracle's own — L of the previous row by image column, three directions — in the same 3 * W * D u16."""
    GROUPS = {0: (0, 4, 5), 1: (1, 6, 7)}

    def __init__(self, orc, p, Cv):
        self.orc, self.p, self.C = orc, p, Cv
        self.S = np.zeros((p.height, p.width, p.num_disp), np.uint16)
        self.wta = []

    def _state(self, ptr):
        import ctypes
        n = 3 * self.p.width * self.p.num_disp
        return np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(ptr)).reshape(3, self.p.width, self.p.num_disp)

    def rows_begin(self, y0, rows):
        self.S[y0:y0 + rows] = 0

    def run(self, stage):
        pass  # the cost volume is given

    def sgm_rows(self, group, y0, rows, state_in=0, state_out=0):
        p = self.p
        if group == 2:
            for di in (2, 3):
                self.S[y0:y0 + rows] += self.orc.sgm_single_path(p, self.C, di)[y0:y0 + rows]  # horizontal paths are row-local
            return
        first = (y0 == 0) if group == 0 else (y0 + rows == p.height)
        for slot, di in enumerate(self.GROUPS[group]):
            pin = None if first else np.ascontiguousarray(self._state(state_in)[slot])
            pout = np.zeros((p.width, p.num_disp), np.uint16)
            self.orc.sgm_rows(p, self.C, di, y0, rows, pin, self.S, pout)
            if state_out:
                self._state(state_out)[slot] = pout

    def wta_rows(self, ptr, y0, rows):
        self.wta.append((y0, rows))


def _oracle_rows_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        orc = Oracle()
        h, w, D = 37, 44, 16
        sc = synth.make_scene(h, w, D, [(-1, 0), (1, 1)], 91)
        p = abi.make_params(w, h, D, [(-1, 0), (1, 1)], win_half=3, n_paths=8, lr_gx=-1)
        Cv = orc.box_cost(p, orc.ad_volume(p, sc["ref"], sc["others"]))
        be = _OracleRowsBackend(orc, p, Cv)
        y0, y1 = sdist.row_sharded_compute(be, p, rank, world, None, {}, device="cpu")
        mine = torch.zeros(be.S.shape, dtype=torch.int32)
        mine[y0:y1] = torch.from_numpy(be.S[y0:y1].astype(np.int32))
        dist.all_reduce(mine)
        if rank == 0:
            q.put(bool(np.array_equal(mine.numpy().astype(np.uint16), orc.sgm_aggregate(p, Cv))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_block_pipeline_with_the_oracle_recurrence(world):
    """the row-block scheme with the real SGM recurrence (oracle) in every process: the blocks' aggregation volumes assemble to the
    whole-frame oracle aggregation — the scheme is exact, independent of the CUDA kernels"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_oracle_rows_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10)


def test_rows_skewed_order_is_a_software_pipeline_across_the_ranks():
    """the enqueue order of dist.rows_skewed_order (no GPU): every frame's parts run in order on every rank, a sweep runs on a rank exactly one
    iteration after it ran on the rank it comes from, and a context (frame f % (world + 1)) is never reused before its frame is complete"""
    for world in range(1, 9):
        frames = 3 * world + 2
        P = world + 1
        orders = [sdist.rows_skewed_order(r, world, frames) for r in range(world)]
        when = [{} for _ in range(world)]  # (part, frame) -> (iteration, position inside it)
        for r in range(world):
            assert len(orders[r]) == frames + world
            for i, step in enumerate(orders[r]):
                for k, item in enumerate(step):
                    assert item not in when[r], "enqueued twice"
                    when[r][item] = (i, k)
            for f in range(frames):
                assert when[r][(0, f)] < when[r][(1, f)] < when[r][(2, f)], "parts of a frame in order"
                if f + P < frames:  # the context is free again before its next frame starts
                    assert when[r][(2, f)] < when[r][(0, f + P)]
        down_first = lambda r: r < (world + 1) // 2  # csrc/sva_dist.cu: the sweep that reaches the rank first is its part 1
        down = lambda r, f: when[r][(1 if down_first(r) else 2, f)][0]
        up = lambda r, f: when[r][(2 if down_first(r) else 1, f)][0]
        for f in range(frames):
            for r in range(1, world):
                assert down(r, f) == down(r - 1, f) + 1, "the down sweep moves one rank per iteration"
                assert up(r - 1, f) == up(r, f) + 1, "the up sweep moves one rank per iteration"
