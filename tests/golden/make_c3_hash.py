"""Generates tests/golden/c3_full_oracle.json: SHA-256 digests of the ORACLE's output (oracle/sva_oracle.c, the CPU restatement) on the full
c3 frame (3840x2160, D = 256, 15 pairs, 8 paths; seed per SURVEY §8d), plus digests of the input views so that a run elsewhere can tell a
different synthetic scene from a different result.  The maps themselves are 50 MB and stay out of the repository; the digests let every
multi-GPU scheme be checked against the oracle at full size in the time it takes to hash two arrays (tools/check_sharded.py --c3), and
tests/test_gpu_full_size.py re-derives them from a live oracle run on the GPU box.

    python tests/golden/make_c3_hash.py      (about 2 minutes and 14 GB of host memory on 8 cores)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402
from stereovisionarray_b200 import abi, configs  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def summarize(sc, disp, sub):
    valid = disp != abi.SVA_DISP_INVALID
    return {"config": "c3", "win_half": 20, "inputs_sha256": digest(np.stack([sc["ref"]] + list(sc["others"]))),
            "disp_sha256": digest(disp), "subpix_sha256": digest(sub), "valid_pixels": int(valid.sum()),
            "disp_sum": int(disp[valid].astype(np.int64).sum()), "gt_match": int((disp[valid] == sc["gt"][valid]).sum())}


if __name__ == "__main__":
    sc = configs.scene("c3")
    p = configs.params("c3")
    disp, sub = Oracle().depth_from_array(p, sc["ref"], sc["others"], sc["mask"])
    out = summarize(sc, disp, sub)
    with open(os.path.join(ROOT, "tests", "golden", "c3_full_oracle.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)
