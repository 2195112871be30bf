#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the multi-camera depth hot path.

  python bench.py --gpus N --steps K --warmup W [--config c1] [--win-half 20] [--impl reference]

A step = one pass of the hot path (K1a AD volume -> K1b box/pack -> 8-path SGM -> WTA/LR/sub-pixel) over one synthetic
frame per rank.  `value` = whole-job MDE/s with inputs resident in HBM (CUDA events on the library's stream, max over ranks);
`e2e` = the same metric through the host-buffer C-ABI call sva_depth_from_array (pinned host inputs, H2D + D2H inside the timed
region).  Volumes (>= 300 MB each at c1) exceed the 126 MB L2, so no L2 flush is needed between iterations.
`--impl reference` times the reference's own CPU code (oracle/_ref: the unmodified sources compiled against oracle/cvshim) on
bounded bands of the same resolution; rank 0 only."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from stereovisionarray_b200 import abi, configs, synth  # noqa: E402

FALLBACK_HBM_GBS = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.p:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            text = self.p.communicate(timeout=5)[0]
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        for line in text.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def algorithmic_bytes(name, p, n_cam, launches_per_step=1.0):
    """per launch, DESIGN.md §4: s_C = s_S = 2 bytes"""
    de = p.width * p.height * p.num_disp
    px = p.width * p.height
    if name in ("k_ad_planar", "k_ad_tile"):
        return n_cam * px + 2 * de
    if name in ("k_box_cost", "k_box_planar"):
        return 4 * de
    if name.startswith("k_sgm_store"):
        return 4 * de
    if name == "k_sgm_red_multi":  # bytes per launch: the frame's 6P - 4 ~ 6P bytes per cell spread over its launches
        return 6 * de * p.n_paths / max(1.0, launches_per_step)  # every direction streams C once and read-modify-writes S once
    if name == "k_wta_tile":
        return 2 * de + 6 * px
    if name == "k_wta_seg":
        return 2 * de + 6 * px   # S once, winner map + the other view's key map
    if name == "k_wta_finish":
        return 18 * px           # winner, key, mask, three S cells, u16 + f32 outputs
    if name == "k_lr_check":
        return 10 * px
    if name.startswith("k_sgm_red"):
        return 6 * de
    return None


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def run_sva(args):
    import torch
    rank, local_rank, world = dist_env()
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from stereovisionarray_b200.pipeline import DepthContext
    name = args.config
    cfg = configs.CONFIGS[name]
    p = configs.params(name, win_half=args.win_half)
    n_cam = p.n_pairs + 1
    if name == "c3":
        return run_pair_sharded(args, rank, local_rank, world)
    # capture batch: c4 has 64 frames per step split over the ranks (strong scaling); the others one frame per rank per step (weak)
    from stereovisionarray_b200 import dist as sdist
    batch = cfg["frames"]
    fb, fe = sdist.frame_range(batch, world, rank) if batch > 1 else (rank, rank + 1)
    frames_rank = fe - fb
    frames_step_total = batch if batch > 1 else world
    sc = configs.scene(name, frame=fb)  # every rank works on its own frames of the capture batch
    ctx = DepthContext(local_rank)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not use_dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident: inputs already in HBM ----
    ctx.upload(p, sc["ref"], sc["others"], sc["mask"])
    for _ in range(args.warmup):
        ctx.run(abi.STAGE_ALL)
    ctx.synchronize()
    l0 = ctx.launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    total_ms, kern = ctx.time_detailed(abi.STAGE_ALL, args.steps * frames_rank)
    barrier()
    launches = ctx.launches() - l0
    total_ms = max_over_ranks(total_ms)
    mde = configs.mde_per_frame(name)
    value = frames_step_total * mde * args.steps / (total_ms / 1e3)

    # ---- end to end: pinned host buffers -> C-ABI call -> host results ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = [pinned(sc["ref"])] + [pinned(o) for o in sc["others"]] + ([pinned(sc["mask"])] if sc["mask"] is not None else [])
    ref_h = keep[0][1]; others_h = [k[1] for k in keep[1:1 + p.n_pairs]]; mask_h = keep[-1][1] if sc["mask"] is not None else None
    others_c = abi.image_array(others_h)
    outs = []
    for _ in range(2):  # two frames are in flight: two sets of pinned result buffers
        disp_t = torch.empty((p.height, p.width), dtype=torch.uint16).pin_memory(); sub_t = torch.empty((p.height, p.width), dtype=torch.float32).pin_memory()
        outs.append((disp_t, sub_t, disp_t.numpy(), sub_t.numpy()))
    disp_h, sub_h = outs[0][2], outs[0][3]
    # (a) one synchronous call per frame: upload, kernels and download back to back
    for _ in range(max(1, args.warmup)):
        ctx.depth_from_array(p, ref_h, others_c, mask_h, disp_h, sub_h)
    barrier()
    ctx.timer_start()
    for _ in range(args.steps * frames_rank):
        ctx.depth_from_array(p, ref_h, others_c, mask_h, disp_h, sub_h)
    single_ms = ctx.timer_stop()
    # (b) the capture-stream API (what a user with a stream of frames calls): same work per frame, but frame t+1's upload and frame
    # t-1's download overlap frame t's kernels.  Timed on the device from the first upload to the end of the last download.
    for i in range(max(2, args.warmup)):
        ctx.stream_wait(ctx.stream_submit(p, ref_h, others_c, mask_h, outs[i % 2][2], outs[i % 2][3]))
    barrier()
    ctx.stream_mark()
    last = None
    for i in range(args.steps * frames_rank):
        last = ctx.stream_submit(p, ref_h, others_c, mask_h, outs[i % 2][2], outs[i % 2][3])
    e2e_ms = ctx.stream_elapsed(last)
    barrier()
    clocks = sampler.stop() if sampler else None
    e2e_ms = max_over_ranks(e2e_ms)
    single_ms = max_over_ranks(single_ms)
    e2e_value = frames_step_total * mde * args.steps / (e2e_ms / 1e3)
    h2d = frames_rank * (n_cam * p.width * p.height + (p.width * p.height if mask_h is not None else 0))
    d2h = frames_rank * p.width * p.height * (2 + 4)

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel + the per-kernel table ----
    peak, peak_src = measured_peak()
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("%s_k%d" % (name, args.win_half), {})
    except Exception:
        pass
    rows = []
    for kname, (sum_ms, cnt) in kern.items():
        ab = algorithmic_bytes(kname, p, n_cam, cnt / args.steps / frames_rank)
        avg_ms = sum_ms / max(1, cnt)
        rows.append({"kernel": kname, "launches_per_step": cnt / args.steps / frames_rank, "avg_ms": round(avg_ms, 4), "share": round(sum_ms / total_ms, 4),
                     "algorithmic_bytes": ab, "achieved_gbs": round(ab / avg_ms / 1e6, 1) if ab else None,
                     "frac": round(ab / avg_ms / 1e6 / peak, 4) if ab else None, "traffic": traffic.get(kname)})
    rows.sort(key=lambda r: -r["share"])
    dom = next(r for r in rows if r["algorithmic_bytes"])
    sgm_ms = sum(s for k, (s, c) in kern.items() if k.startswith("k_sgm")) / args.steps / frames_rank
    sgm_bytes = (6 * p.n_paths - 4) * p.width * p.height * p.num_disp if p.n_paths else 0
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": dom["traffic"], "peak_source": peak_src, "frac_of_8000": round(dom["achieved_gbs"] / 8000.0, 4),
                # actual DRAM bytes (ncu) / event-timed duration: below `achieved` when L2 sharing moves fewer bytes than the algorithmic count
                "dram_gbs": round(dom["traffic"] / dom["avg_ms"] / 1e6, 1) if dom["traffic"] else None,
                "dram_frac": round(dom["traffic"] / dom["avg_ms"] / 1e6 / peak, 4) if dom["traffic"] else None,
                "sgm_stage": {"algorithmic_bytes": sgm_bytes, "ms": round(sgm_ms, 4), "achieved_gbs": round(sgm_bytes / sgm_ms / 1e6, 1) if sgm_ms else None,
                              "frac": round(sgm_bytes / sgm_ms / 1e6 / peak, 4) if sgm_ms else None}}
    # ---- CPU baseline (rank 0, N = 1): the oracle port of the SAME pipeline on a bounded row band, all host cores ----
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle.oracle import Oracle
        orc = Oracle()
        band = min(p.height, args.cpu_band)
        scb = configs.scene(name, frame=0, height=band)
        pb = abi.make_params(p.width, band, p.num_disp, configs.offsets(name), win_half=args.win_half, n_paths=p.n_paths, lr_gx=-1, subpixel=1)
        orc.depth_from_array(pb, scb["ref"], scb["others"], scb["mask"])  # warm-up
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            orc.depth_from_array(pb, scb["ref"], scb["others"], scb["mask"])
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": round(p.width * band * p.num_disp / 1e6 / best, 2), "unit": "MDE/s", "cores": orc.num_threads(), "kind": "port",
               "sample": "oracle/sva_oracle.c volume pipeline (OpenMP) on a %dx%d row band of the %s frame, D=%d, %d pairs, best of 3" % (p.width, band, name, p.num_disp, p.n_pairs)}
    literal = None
    if world == 1 and not args.no_cpu:
        try:
            literal = literal_mode_line(p.width, p.height)
        except Exception as ex:  # reported, never fatal for the headline line
            literal = {"error": str(ex)[:200]}
    out = {
        "metric": "MDE/s", "value": round(value, 1), "unit": "MDE/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "strong" if batch > 1 else "weak", "vs_baseline": None, "dtype": "u16",
        "data": "synthetic", "frames_per_s": round(frames_step_total * args.steps / (total_ms / 1e3), 2),
        "config": {"workload": "%s: %s" % (name, cfg["desc"]), "width": p.width, "height": p.height, "num_disp": p.num_disp, "cameras": n_cam,
                   "pairs": p.n_pairs, "win_half": p.win_half, "sgm_paths": p.n_paths, "frames_per_step_per_gpu": frames_rank,
                   "partitioning": "independent frames per GPU, no data-path collective" if world > 1 else "single GPU",
                   "l2": "no flush: each volume (%.0f MB) exceeds the 126 MB L2" % (p.width * p.height * p.num_disp * 2 / 1e6)},
        "e2e": {"value": round(e2e_value, 1), "unit": "MDE/s", "ms_per_step": round(e2e_ms / args.steps, 4), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "sva_stream_submit / sva_stream_wait (C ABI, pinned host buffers, two frames in flight)",
                "single_call_ms_per_step": round(single_ms / args.steps, 4), "single_call_api": "sva_depth_from_array, one synchronous call per frame"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "literal_mode": literal, "clocks": clocks, "kernels": rows,
    }
    emit(json.dumps(out))
    if use_dist:
        dist.destroy_process_group()


def run_pair_sharded(args, rank, local_rank, world):
    """c3: ONE frame per step; the camera pairs are split over the ranks, the packed AD partials are sum-reduced (NCCL) onto rank 0,
    which runs the box filter, SGM and WTA.  Strong scaling.  Timed with CUDA events on torch's current stream (the library is told
    to run on that stream so its kernels and the NCCL reduce are ordered)."""
    import torch
    import torch.distributed as dist
    from stereovisionarray_b200 import dist as sdist
    from stereovisionarray_b200.pipeline import DepthContext
    name = "c3"
    cfg = configs.CONFIGS[name]
    p = configs.params(name, win_half=args.win_half)
    sc = configs.scene(name, frame=0)
    torch.cuda.set_device(local_rank)
    ctx = DepthContext(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    b, e = sdist.pair_ranges(p.n_pairs, world)[rank]
    slices = args.c3_scheme == "slices" and world > 1  # one GPU: nothing to exchange, the frame runs as one plain pipeline
    rows = args.c3_scheme == "rows" and world > 1
    keep = {}
    if slices:
        ctx.upload(sdist.slice_params(p, rank, world), sc["ref"], sc["others"], sc["mask"])
    elif rows:
        ctx.upload(p, sc["ref"], sc["others"], sc["mask"])
    else:
        ctx.upload(p, sc["ref"], sc["others"], sc["mask"])
        ptr, nbytes = ctx.ad_device_ptr()
        vol = torch.as_tensor(sdist._CudaAlias(ptr, nbytes // 4), device="cuda")

    def step():
        if slices:
            sdist.slice_sharded_compute(ctx, p, rank, world, None, keep)
            return
        if rows:
            sdist.row_sharded_compute(ctx, p, rank, world, None, keep)
            return
        if e > b:
            ctx.set_pair_range(b, e)
            ctx.run(abi.STAGE_AD)
        else:
            vol.zero_()
        if world > 1:
            dist.reduce(vol, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            ctx.mark_ad_ready()
            ctx.run(abi.STAGE_BOX)
            ctx.run(abi.STAGE_SGM)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = ctx.launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    launches = ctx.launches() - l0
    if rank == 0:
        mde = configs.mde_per_frame(name)
        peak, peak_src = measured_peak()
        out = {"metric": "MDE/s", "value": round(mde * args.steps / (ms / 1e3), 1), "unit": "MDE/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
               "frames_per_s": round(args.steps / (ms / 1e3), 3),
               "config": {"workload": "%s: %s" % (name, cfg["desc"]), "width": p.width, "height": p.height, "num_disp": p.num_disp, "cameras": p.n_pairs + 1,
                          "pairs": p.n_pairs, "win_half": p.win_half, "sgm_paths": p.n_paths,
                          "partitioning": ("disparity slices of %d (cost volume, no reduction) -> all-gather -> path directions %s -> reduce-scatter by row blocks -> "
                                           "row-sharded WTA; %d ranks" % (p.num_disp // world, sdist.direction_masks(p.n_paths, world), world)) if slices else
                                          ("row blocks %s end to end: cost volume and horizontal paths local, row-sweeping paths as a pipeline handing %.1f MB of path state per hop; no volume collective"
                                           % (sdist.row_blocks(p.height, world)[1], 3 * p.width * p.num_disp * 2 / 1e6)) if rows else
                                          ("single GPU: the whole frame, no exchange" if world == 1 else
                                           "pairs %s over %d ranks; packed-int32 NCCL reduce of the AD volume (%.2f GB) onto rank 0" % (sdist.pair_ranges(p.n_pairs, world), world, nbytes / 1e9)),
                          "l2": "no flush: each volume (%.0f MB) exceeds the 126 MB L2" % (p.width * p.height * p.num_disp * 2 / 1e6)},
               "e2e": None, "gpu_launches": int(launches), "roofline": {"bound": "hbm", "peak": peak, "peak_source": peak_src, "note": "see the c1 line for per-kernel rooflines"},
               "cpu_baseline": None, "clocks": clocks}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _ref_band_worker(job):
    """one band through the reference's own main() (loop nest src/CameraStereoVision.cpp:49-95); returns (seconds, pixel*candidate evaluations)"""
    w, band, seed = job
    import contextlib
    import io
    from oracle.oracle import Oracle, Reference
    r = Reference()
    sc = synth.make_literal_scene(band, w, seed)
    mask = np.zeros((band, w), np.uint8)
    mask[20:band - 20, 20:w - 20] = 255
    up = [np.repeat(np.repeat(i, 2, axis=0), 2, axis=1) for i in sc["images"]]
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the driver prints to std::cout
    try:
        t0 = time.perf_counter()
        disp, _, _ = r.main_run(up, mask)
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    return dt, literal_evals(w, band)


def literal_evals(w, h, stride=16):
    """pixel x candidate evaluations of the reference's loop nest for pair {12,11} on a w x h frame (geometry only; every `stride`-th
    column is walked and scaled)"""
    from oracle.oracle import Oracle
    o = Oracle()
    cams = [abi.camera(*c) for c in synth.reference_cameras(w)]
    hx, hy = w // 2, h // 2
    evals = 0
    for y in range(20, h - 20):
        for x in range(20, w - 20, stride):
            ray = o.camera_inv_project(cams[12], (x - hx, y - hy))
            a = o.camera_project(cams[11], [cams[12].pos[i] + ray[i] * 0.5 for i in range(3)])
            b = o.camera_project(cams[11], [cams[12].pos[i] + ray[i] * 1.0 for i in range(3)])
            a, b = (a[0] + hx, a[1] + hy), (b[0] + hx, b[1] + hy)
            if min(a[0], b[0]) < 20 or max(a[0], b[0]) > w - 20 or min(a[1], b[1]) < 20 or max(a[1], b[1]) > h - 20:
                continue
            evals += stride * (max(abs(a[0] - b[0]), abs(a[1] - b[1])) + 1)
    return evals


def literal_mode_line(w, h, reps=5):
    """the reference's OWN algorithm (SAD 40x40 over the Bresenham candidates, first-min WTA, improveWithDisparity) on the GPU through the
    C ABI, whole w x h frame, host buffers in and out — the like-for-like counterpart of `--impl reference` (same MDE definition)"""
    from stereovisionarray_b200 import reference_api as api
    sc = synth.make_literal_scene(h, w, 7000)
    cams = [api.Camera(f, pos, ps) for pos, f, ps in synth.reference_cameras(w)]
    mask = np.zeros((h, w), np.uint8)
    mask[20:h - 20, 20:w - 20] = 255
    run = lambda: api.improveWithDisparity(api.matchLiteral(sc["images"], cams, [(12, 11)], mask, 20, 0.5, 1.0), sc["images"][12], [sc["images"][11]],
                                           [(cams[12], cams[11])], 21, mask)
    run()
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    dt = (time.perf_counter() - t0) / reps
    ev = literal_evals(w, h, stride=64)
    return {"ms_per_frame": round(dt * 1e3, 3), "value": round(ev / 1e6 / dt, 1), "unit": "MDE/s (pixel x candidate evaluations, as --impl reference)",
            "api": "sva_match_literal + sva_improve_with_disparity (host buffers, wall clock)", "frame": "%dx%d, pair {12,11}, mask = interior" % (w, h)}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle.oracle import Reference
    name = args.config
    cfg = configs.CONFIGS[name]
    if not Reference.available():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsva_ref.so missing (built only where /root/reference exists)"}))
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    w, band = cfg["width"], args.ref_band
    ctxm = mp.get_context("fork")
    times = []
    with ctxm.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            res = pool.map(_ref_band_worker, [(w, band, 7000 + 100 * step + i) for i in range(cores)])
            if step >= args.warmup:  # the bands run concurrently: a step lasts as long as its slowest band (scene synthesis is not timed)
                times.append((max(t for t, _ in res), sum(e for _, e in res)))
    tot_t = sum(t for t, _ in times)
    tot_e = sum(e for _, e in times)
    v = tot_e / 1e6 / tot_t
    sample = ("UNMODIFIED reference main() (oracle/_ref): SAD 40x40 + first-min WTA + improveWithDisparity on %d independent %dx%d bands per step "
              "(one per core; the reference is single-threaded and has no multi-pair sum / SGM), pair {12,11}, MDE = pixel x candidate evaluations" % (cores, w, band))
    out = {"impl": "reference", "metric": "MDE/s", "value": round(v, 3), "unit": "MDE/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(1e3 * tot_t / max(1, len(times)), 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": "%s: %s" % (name, cfg["desc"]), "width": w, "band_rows": band},
           "cpu_baseline": {"value": round(v, 3), "unit": "MDE/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": round(v, 3), "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(json.dumps(out))


class StdoutToStderr:
    """NCCL / torchrun print banners on fd 1; the contract is ONE JSON line on stdout, so everything else goes to stderr"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_real_print = print


def emit(line):
    """the single JSON line, written to the REAL stdout even while fd 1 is redirected"""
    os.write(_STDOUT_FD, (line + "\n").encode())


_STDOUT_FD = os.dup(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c1", choices=sorted(configs.CONFIGS))
    ap.add_argument("--win-half", type=int, default=20)
    ap.add_argument("--impl", default="sva", choices=["sva", "reference"])
    ap.add_argument("--cpu-band", type=int, default=256, help="rows of the CPU-baseline sample")
    ap.add_argument("--ref-band", type=int, default=44, help="rows per band of the reference arm (40 + valid rows)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c3-scheme", default="rows", choices=["pairs", "slices", "rows"],
                    help="c3 only: 'pairs' = north_star's pair sharding + NCCL reduce of the AD volume; 'slices' = disparity-slice / direction / row sharding; "
                         "'rows' = row blocks end to end, row-sweeping paths as a pipeline between neighbours (DESIGN.md §7)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "sva" else args.warmup
    with StdoutToStderr():
        if args.impl == "reference":
            run_reference(args)
        else:
            run_sva(args)


if __name__ == "__main__":
    main()
