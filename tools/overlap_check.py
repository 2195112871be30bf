"""Throughput of N contexts (own streams) running the resident pipeline concurrently vs one context."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereovisionarray_b200 import abi, configs
from stereovisionarray_b200.pipeline import DepthContext
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
p = configs.params(name)
for n in (1, 2, 3):
    ctxs = [DepthContext(0) for _ in range(n)]
    for i, c in enumerate(ctxs):
        sc = configs.scene(name, frame=i)
        c.upload(p, sc["ref"], sc["others"], sc["mask"])
        for _ in range(3):
            c.run(abi.STAGE_ALL)
    for c in ctxs:
        c.synchronize()
    K = 30
    t0 = time.perf_counter()
    for _ in range(K):
        for c in ctxs:
            c.run(abi.STAGE_ALL)
    for c in ctxs:
        c.synchronize()
    dt = time.perf_counter() - t0
    print("contexts=%d  %.4f ms/frame" % (n, dt * 1e3 / (K * n)))
    for c in ctxs:
        c.close()
