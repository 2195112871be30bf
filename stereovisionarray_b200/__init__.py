"""stereovisionarray_b200 — B200 (sm_100a) implementation of the StereoVisionArray multi-camera depth hot path.

Host-side mirror of the reference's interface (include/functions.h, include/Camera.h) over the C ABI declared in
include/sva_c_api.h.  There is no CPU fallback: the compute entry points raise if libsva_b200.so or a B200 is missing.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
