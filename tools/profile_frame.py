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
a = ap.parse_args()
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
print("ok", ctx.launches())
