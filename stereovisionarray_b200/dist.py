"""Multi-GPU sharding of the depth path on one 8xB200 box — one process per GPU, torch.distributed for plumbing.

Two natural partitions, nothing else:
  * frames of a capture batch are independent units -> rank r takes frames [r*F/G, (r+1)*F/G); NO data-path collective;
  * camera pairs of ONE frame: the cost volume is a sum over pairs (integer, associative, commutative), so each rank
    computes the AD partial of its pair range over the full [H][W][D] volume and the partials are sum-reduced onto the
    rank that owns the reference view, which then runs the box filter, SGM and WTA.  NCCL has no 16-bit integer type, so
    the u16 volume is reduced as packed int32 (two cells per word): totals are <= 255 * n_pairs <= 8160 < 2^15, so no
    carry crosses the half-word boundary and the sign bit is never set — bit-exact in any reduction order.

The compute backend is an object with upload / set_pair_range / run / ad_device_ptr ... (stereovisionarray_b200.pipeline.
DepthContext on GPUs).  The helper functions below are pure host logic and are exercised on CPU with the gloo backend."""
import numpy as np


def frame_range(n_frames, world, rank):
    """contiguous, balanced: the first (n_frames % world) ranks get one extra frame"""
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pair_ranges(n_pairs, world):
    """balanced contiguous pair ranges, e.g. 15 pairs over 8 ranks -> 2,2,2,2,2,2,2,1 (ranks beyond the pairs get empty ranges)"""
    out = []
    base, extra = divmod(n_pairs, world)
    b = 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def packed_no_carry(n_pairs):
    """the packed-int32 reduce is exact iff the full sum of a cell stays below 2^15"""
    return 255 * n_pairs < (1 << 15)


def reduce_packed_u16(t_int32, dst, group=None):
    """sum-reduce a packed-u16x2 volume (viewed as int32) onto rank `dst`; in place on dst"""
    import torch.distributed as dist
    dist.reduce(t_int32, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return t_int32


class _CudaAlias:
    """zero-copy torch view of a device buffer owned by libsva_b200.so"""

    def __init__(self, ptr, n_int32):
        self.__cuda_array_interface__ = {"shape": (n_int32,), "typestr": "<i4", "data": (ptr, False), "version": 3, "strides": None}


def pair_sharded_depth(ctx, p, ref, others, mask, rank, world, owner=0, group=None):
    """One frame, pairs sharded over `world` GPUs (config c3).  Every rank holds the frame's images; returns (disp, subpix) on the
    owner rank and None elsewhere.  ctx must run on torch's current stream (ctx.set_stream) so NCCL and the kernels are ordered."""
    import torch
    from . import abi
    if not packed_no_carry(p.n_pairs):
        raise ValueError("too many pairs for the packed-int32 reduce")
    b, e = pair_ranges(p.n_pairs, world)[rank]
    ctx.upload(p, ref, others, mask)
    ptr, nbytes = ctx.ad_device_ptr()
    vol = torch.as_tensor(_CudaAlias(ptr, nbytes // 4), device="cuda")
    if e > b:
        ctx.set_pair_range(b, e)
        ctx.run(abi.STAGE_AD)
    else:
        vol.zero_()
    if world > 1:
        reduce_packed_u16(vol, owner, group)
    if rank != owner:
        return None
    ctx.mark_ad_ready()
    ctx.run(abi.STAGE_BOX)
    ctx.run(abi.STAGE_SGM)
    return ctx.download_disparity()


def numpy_pack(a_u16):
    """[H][W][D] u16 (D even) -> int32 view, the layout the GPU path reduces"""
    return np.ascontiguousarray(a_u16).view(np.int32)
