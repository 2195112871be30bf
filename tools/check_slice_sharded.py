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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
