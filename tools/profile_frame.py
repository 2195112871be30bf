"""Small driver for ncu: uploads one synthetic frame of a config and runs the resident pipeline a few times."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereovisionarray_b200 import abi, configs  # noqa: E402
from stereovisionarray_b200.pipeline import DepthContext  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c1")
ap.add_argument("--win-half", type=int, default=20)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--times", action="store_true", help="print the CUDA-event time of every launch of one more frame")
ap.add_argument("--avg", type=int, default=0, help="print the average CUDA-event time per kernel over this many more frames")
ap.add_argument("--size", default="", help="WxHxD: a 3x3-array frame of random pixels of this size instead of a named config")
a = ap.parse_args()
if a.size:
    import numpy as np
    w, h, d = (int(v) for v in a.size.split("x"))
    off8 = [(gx, gy) for gy in (-1, 0, 1) for gx in (-1, 0, 1) if (gx, gy) != (0, 0)]
    p = abi.make_params(w, h, d, off8, win_half=a.win_half, n_paths=8, lr_gx=-1)
    rng = np.random.default_rng(1)
    sc = {"ref": rng.integers(0, 256, (h, w), dtype=np.uint8), "others": [rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in off8], "mask": None}
else:
    p = configs.params(a.config, win_half=a.win_half)
    sc = configs.scene(a.config)
ctx = DepthContext(0)
ctx.upload(p, sc["ref"], sc["others"], sc["mask"])
for _ in range(a.iters):
    ctx.run(abi.STAGE_ALL)
ctx.synchronize()
if a.times:
    kt = ctx.kernel_times(abi.STAGE_ALL)
    print(" ".join("%s=%.4f" % (n, ms) for n, ms in kt), "total=%.4f" % sum(ms for _, ms in kt))
if a.avg:
    tot, per = ctx.time_detailed(abi.STAGE_ALL, a.avg)
    print(" ".join("%s=%.4f" % (n.split("/")[0][:24], sm / a.avg) for n, (sm, c) in per.items()), "total=%.4f" % (tot / a.avg))
print("ok", ctx.launches())
