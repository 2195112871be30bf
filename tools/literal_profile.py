"""One literal-mode frame (the reference's own algorithm on the GPU) for ncu launch lists: python tools/literal_profile.py [w h]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereovisionarray_b200 import reference_api as api, synth  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 960)
sc = synth.make_literal_scene(h, w, 7000)
cams = [api.Camera(f, pos, ps) for pos, f, ps in synth.reference_cameras(w)]
mask = np.zeros((h, w), np.uint8)
mask[20:h - 20, 20:w - 20] = 255
for i in range(2):
    t0 = time.perf_counter()
    d = api.matchLiteral(sc["images"], cams, [(12, 11)], mask, 20, 0.5, 1.0)
    t1 = time.perf_counter()
    imp = api.improveWithDisparity(d, sc["images"][12], [sc["images"][11]], [(cams[12], cams[11])], 21, mask)
    t2 = time.perf_counter()
    print("match %.2f ms  improve %.2f ms  (disp max %d)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, int(d.max())))
