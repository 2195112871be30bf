"""The five BASELINE.json configurations as parameter sets + synthetic-scene recipes."""
from . import abi, synth


def grid_offsets(rows, cols, ref_index):
    """grid offsets (gx, gy) of every other camera of a rows x cols array, in index order (the TO_CENTER analogue)"""
    rr, rc = divmod(ref_index, cols)
    return [(i % cols - rc, i // cols - rr) for i in range(rows * cols) if i != ref_index]


CONFIGS = {
    # name: width, height, D, grid (rows, cols, ref), paths, face-ROI, frames
    "c0": dict(index=0, desc="2-camera rectified pair 640x480, D=64, 4-path SGM", width=640, height=480, num_disp=64, grid=(1, 2, 1), n_paths=4, face=False, frames=1),
    "c1": dict(index=1, desc="3x3 camera array 1280x960, D=128, 8-path SGM, summed cost volume", width=1280, height=960, num_disp=128, grid=(3, 3, 4), n_paths=8, face=False, frames=1),
    "c2": dict(index=2, desc="face-ROI crop 1024x1024 from a 9-camera array, D=192, sub-pixel", width=1024, height=1024, num_disp=192, grid=(3, 3, 4), n_paths=8, face=True, frames=1),
    "c3": dict(index=3, desc="16-camera array 3840x2160, D=256, pairs sharded + cost-volume reduce", width=3840, height=2160, num_disp=256, grid=(4, 4, 5), n_paths=8, face=False, frames=1),
    "c4": dict(index=4, desc="capture stream 64 frames x 1920x1080, 9 cameras, D=192, frames partitioned", width=1920, height=1080, num_disp=192, grid=(3, 3, 4), n_paths=8, face=False, frames=64),
}


def offsets(name):
    r, c, ref = CONFIGS[name]["grid"]
    return grid_offsets(r, c, ref)


def params(name, win_half=20, **kw):
    c = CONFIGS[name]
    return abi.make_params(c["width"], c["height"], c["num_disp"], offsets(name), win_half=win_half, n_paths=c["n_paths"], lr_gx=-1, subpixel=1, **kw)


def scene(name, frame=0, height=None):
    """deterministic synthetic frame: seed = 1000 * config_index + frame_index; `height` crops the config to a row band"""
    c = CONFIGS[name]
    return synth.make_scene(height or c["height"], c["width"], c["num_disp"], offsets(name), 1000 * c["index"] + frame, face=c["face"])


def mde_per_frame(name):
    c = CONFIGS[name]
    return c["width"] * c["height"] * c["num_disp"] / 1e6
