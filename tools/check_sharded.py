"""torchrun check: dist.slice_sharded_depth and dist.row_sharded_depth over WORLD_SIZE GPUs == the single-GPU one-call path, bit for bit
(rank 0 compares)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereovisionarray_b200 import abi, dist as sdist, synth  # noqa: E402
from stereovisionarray_b200.pipeline import DepthContext  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
OFF15 = [(gx, gy) for gy in range(-1, 3) for gx in range(-1, 3) if (gx, gy) != (0, 0)]
ok = True
for (h, w, D, off, k) in [(270, 480, 256, OFF15, 20), (203, 333, 64 * world // (2 if world > 4 else 1) if False else 128, OFF15[:8], 7)]:
    if D % world or (D // world) % 8:
        continue
    sc = synth.make_scene(h, w, D, off, 77, face=True)
    p = abi.make_params(w, h, D, off, win_half=k, n_paths=8, lr_gx=-1)
    ctx = DepthContext(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    keep = {}
    for rep in range(2):  # twice: cached buffers, re-upload
        out = sdist.slice_sharded_depth(ctx, p, sc["ref"], sc["others"], sc["mask"], rank, world, None, keep)
    keep_rows = {}
    for rep in range(2):
        out_rows = sdist.row_sharded_depth(ctx, p, sc["ref"], sc["others"], sc["mask"], rank, world, None, keep_rows)
    if rank == 0:
        ref = DepthContext(local)
        d0, s0 = ref.depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
        for name, o in (("slice-sharded", out), ("row-sharded", out_rows)):
            same = np.array_equal(o[0], d0) and np.array_equal(o[1], s0)
            print("%s %dx%dx%d over %d GPUs: %s (valid %.3f)" % (name, w, h, D, world, "bit-exact" if same else "MISMATCH", float((d0 != 0xFFFF).mean())))
            ok = ok and same
        ref.close()
    ctx.close()
if len(sys.argv) > 1:  # e.g. 3840x2160x256: time the row-block pipeline on a frame of random pixels of that size (15 pairs), device time, max over ranks
    w, h, D = (int(v) for v in sys.argv[1].split("x"))
    rng = np.random.default_rng(5)
    ref_img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    others = [rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in OFF15]
    p = abi.make_params(w, h, D, OFF15, win_half=20, n_paths=8, lr_gx=-1)
    ctx = DepthContext(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.upload(p, ref_img, others, None)
    keep = {}
    for _ in range(2):
        sdist.row_sharded_compute(ctx, p, rank, world, None, keep)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        sdist.row_sharded_compute(ctx, p, rank, world, None, keep)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("row-block pipeline %dx%dx%d, 15 pairs, %d GPUs: %.2f ms per frame (random pixels)" % (w, h, D, world, float(t.item())))
    ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
