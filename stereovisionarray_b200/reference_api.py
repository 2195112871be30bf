"""Python mirror of the reference's interface for the depth path — same names, argument meaning and error behaviour as
include/functions.h and include/Camera.h of Nahuel-M/StereoVisionArray — over the C ABI (include/sva_c_api.h).

Scalar helpers (Camera.project / inv_project, bresenham, getCameraPairs) are host-side shims inside libsva_b200.so;
getAbsDiff, shiftPerspectiveWithDisparity, improveWithDisparity and the batched replacement of the driver's loop nest
(matchLiteral) run on the GPU.  Nothing here falls back to the CPU: without the library and a B200 these raise."""
import ctypes as C

import numpy as np

from . import abi
from ._lib import SvaError, check, lib
from .abi import (CROSS, DIAGONAL, JUMP_CROSS, LINE_HORIZONTAL, LINE_VERTICAL, MID_LEFT, MID_TOP, ORTHOGONAL, TO_CENTER,  # noqa: F401
                  TO_CENTER_SMALL)


class CvException(RuntimeError):
    """what the reference surfaces as cv::Exception (e.g. a ROI outside the image, src/functions.cpp:30,34)"""


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_ctx = {}


def _context(device=0):
    if device not in _ctx:
        h = C.c_void_p()
        rc = lib().sva_create(device, C.byref(h))
        if rc != 0:
            raise RuntimeError("sva_create failed (%d): a B200 (sm_100) GPU is required; there is no CPU fallback" % rc)
        _ctx[device] = h
    return _ctx[device]


class Camera:
    """class Camera — include/Camera.h:6-21: Camera(focal_length, position, pixel_size); members pos3D, f, pixel_size."""

    def __init__(self, focal_length, position, pixel_size):
        self.f = float(focal_length)
        self.pos3D = tuple(float(v) for v in position)
        self.pixel_size = float(pixel_size)

    def _c(self):
        return abi.camera(self.pos3D, self.f, self.pixel_size)

    def project(self, Pos3D):
        """src/Camera.cpp:15-22 -> (x, y) integer pixel offset from the principal point"""
        out = (C.c_int32 * 2)()
        c = self._c()
        check(None, lib().sva_camera_project(C.byref(c), (C.c_double * 3)(*Pos3D), out))
        return int(out[0]), int(out[1])

    def inv_project(self, pixel):
        """src/Camera.cpp:25-33 -> unit ray (x, y, z)"""
        out = (C.c_double * 3)()
        c = self._c()
        check(None, lib().sva_camera_inv_project(C.byref(c), (C.c_int32 * 2)(int(pixel[0]), int(pixel[1])), out))
        return out[0], out[1], out[2]


def bresenham(point1, point2):
    """include/functions.h:45 — list of (x, y), ordered by increasing major-axis coordinate"""
    cap = 2 * (abs(point1[0] - point2[0]) + abs(point1[1] - point2[1])) + 8
    out = np.zeros((cap, 2), np.int32)
    n = check(None, lib().sva_bresenham(int(point1[0]), int(point1[1]), int(point2[0]), int(point2[1]), _p(out, C.c_int32), cap))
    return [tuple(int(v) for v in p) for p in out[:n]]


def getCameraPairs(cameras, pairs, cameraNum=None):
    """include/functions.h:34-36 — list of [ref, other] for the reference's 5x5 array"""
    out = np.zeros((64, 2), np.int32)
    n = check(None, lib().sva_get_camera_pairs(len(cameras), int(pairs), -1 if cameraNum is None else int(cameraNum), _p(out, C.c_int32), 64))
    return [[int(a), int(b)] for a, b in out[:n]]


def gridPairs(grid_rows, grid_cols, ref_index, pairs):
    """generalisation to any grid: ([ref, other] list, [(gx, gy)] list)"""
    pr = np.zeros((64, 2), np.int32); gx = np.zeros(64, np.int32); gy = np.zeros(64, np.int32)
    n = check(None, lib().sva_grid_pairs(grid_rows, grid_cols, ref_index, int(pairs), _p(pr, C.c_int32), _p(gx, C.c_int32), _p(gy, C.c_int32), 64))
    return [[int(a), int(b)] for a, b in pr[:n]], [(int(a), int(b)) for a, b in zip(gx[:n], gy[:n])]


def getAbsDiff(mat1, mat2, device=0):
    """include/functions.h:38 — exact sum |a-b| of two equal-size uint8 views, as float (like the reference's double)"""
    h = _context(device)
    a, ak = abi.image_u8(mat1)
    b, bk = abi.image_u8(mat2)
    out = C.c_double()
    check(h, lib().sva_abs_diff_u8(h, C.byref(a), C.byref(b), C.byref(out)))
    return out.value


def shiftPerspectiveWithDisparity(inputCam, outputCam, disparity, image, device=0):
    """include/functions.h:26 — gather remap of `image` into the reference view (zero where the reference leaves garbage)"""
    h = _context(device)
    d, dk = abi.image_u8(np.ascontiguousarray(disparity, np.uint8))
    im, ik = abi.image_u8(np.ascontiguousarray(image, np.uint8))
    out = np.zeros(ik.shape, np.uint8)
    ci, co = inputCam._c(), outputCam._c()
    check(h, lib().sva_shift_perspective_with_disparity(h, C.byref(ci), C.byref(co), C.byref(d), C.byref(im), _p(out, C.c_uint8)))
    return out


def improveWithDisparity(disparity, centerImage, images, cameras, windowSize, mask, device=0):
    """include/functions.h:22.  `mask` replaces the reference's internal getFaceMask(centerImage) call (src/functions.cpp:13): the
    detector is out of scope, its output is an input here.  Raises CvException where the reference throws."""
    h = _context(device)
    d, dk = abi.image_u8(np.ascontiguousarray(disparity, np.uint8))
    c, ck = abi.image_u8(np.ascontiguousarray(centerImage, np.uint8))
    imgs, keep = abi.image_array(images)
    cams = abi.camera_array([x._c() for pr in cameras for x in pr])
    m, mk = abi.image_u8(np.ascontiguousarray(mask, np.uint8))
    out = np.zeros(ck.shape, np.uint8)
    try:
        check(h, lib().sva_improve_with_disparity(h, C.byref(d), C.byref(c), imgs, cams, len(keep), C.byref(m), int(windowSize), _p(out, C.c_uint8)))
    except SvaError as e:
        if e.code == abi.SVA_ERR_ROI:
            raise CvException(str(e))
        raise
    return out


def matchLiteral(images, cameras, pairs, mask, kernelSize=20, ray_near=0.5, ray_far=1.0, device=0):
    """The driver's loop nest (src/CameraStereoVision.cpp:49-95) as one batched GPU call -> uint8 disparity."""
    h = _context(device)
    imgs, keep = abi.image_array(images)
    cams = abi.camera_array([x._c() for x in cameras])
    pr = np.ascontiguousarray(pairs, np.int32)
    m = None
    if mask is not None:
        m, mk = abi.image_u8(np.ascontiguousarray(mask, np.uint8))
    out = np.zeros(keep[0].shape, np.uint8)
    check(h, lib().sva_match_literal(h, imgs, cams, len(keep), _p(pr, C.c_int32), len(pr), C.byref(m) if m is not None else None,
                                     int(kernelSize), C.c_double(ray_near), C.c_double(ray_far), _p(out, C.c_uint8)))
    return out


def disparityToDepth(disparity, camDistance, f, pixelSize, device=0):
    """src/CameraStereoVision.cpp:47,98-100 — f64 depth, inf where disparity == 0"""
    h = _context(device)
    d, dk = abi.image_u8(np.ascontiguousarray(disparity, np.uint8))
    out = np.zeros(dk.shape, np.float64)
    check(h, lib().sva_disparity_to_depth(h, C.byref(d), C.c_double(camDistance), C.c_double(f), C.c_double(pixelSize), _p(out, C.c_double)))
    return out


# ---- consumers of the depth output ----
def shiftPerspective2(inputCam, outputCam, depthMap, device=0):
    """include/functions.h:24, src/functions.cpp:79-104 — f64 depth map forward-warped into outputCam's view (0 where nothing lands)"""
    h = _context(device)
    d = np.ascontiguousarray(depthMap, np.float64)
    out = np.zeros_like(d)
    ci, co = inputCam._c(), outputCam._c()
    check(h, lib().sva_shift_perspective2(h, C.byref(ci), C.byref(co), _p(d, C.c_double), d.shape[0], d.shape[1], _p(out, C.c_double)))
    return out


def Points3DToDepthMap(points, camera, resolution, device=0):
    """include/functions.h:30, src/functions.cpp:118-133 — resolution = (width, height) like cv::Size"""
    h = _context(device)
    pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    w, hh = int(resolution[0]), int(resolution[1])
    out = np.zeros((hh, w), np.float64)
    c = camera._c()
    check(h, lib().sva_points3d_to_depth_map(h, _p(pts, C.c_double), C.c_int64(len(pts)), C.byref(c), w, hh, _p(out, C.c_double)))
    return out


def DepthMapToPoints3D(depthMap, camera, resolution, device=0):
    """include/functions.h:32, src/functions.cpp:135-146 -> (n, 3) f64 points in the reference's push_back order"""
    h = _context(device)
    d = np.ascontiguousarray(depthMap, np.float64)
    out = np.zeros((d.size, 3), np.float64)
    n = C.c_int64()
    c = camera._c()
    check(h, lib().sva_depth_map_to_points3d(h, _p(d, C.c_double), d.shape[0], d.shape[1], C.byref(c), int(resolution[0]), int(resolution[1]),
                                             _p(out, C.c_double), C.c_int64(d.size), C.byref(n)))
    return out[:n.value].copy()


def calculateAverageError(image, mask, device=0):
    """include/functions.h:53, src/functions.cpp:348-354 — mean of an f64 error / depth map under the face mask (the mask is an input here)"""
    h = _context(device)
    a = np.ascontiguousarray(image, np.float64)
    m, mk = abi.image_u8(np.ascontiguousarray(mask, np.uint8))
    out = C.c_double()
    check(h, lib().sva_masked_mean_f64(h, _p(a, C.c_double), a.shape[0], a.shape[1], C.byref(m), C.byref(out)))
    return out.value


def getGroups(cameras, groupType):
    """include/functions.h:28, src/functions.cpp:107-116 -> list of groups, each a list of (ref, other) pairs"""
    pairs = np.zeros((256, 2), np.int32)
    sizes = np.zeros(64, np.int32)
    ng = check(None, lib().sva_get_groups(len(cameras), str(groupType).encode(), _p(pairs, C.c_int32), 256, _p(sizes, C.c_int32), 64))
    out, o = [], 0
    for g in range(ng):
        out.append([tuple(int(v) for v in pr) for pr in pairs[o:o + sizes[g]]])
        o += int(sizes[g])
    return out


# ---- ingest ----
def resizeHalf(image, device=0):
    """resize(img, img, Size(), 0.5, 0.5) of the driver's loader (src/CameraStereoVision.cpp:18) for even-sized u8 images, on the GPU"""
    h = _context(device)
    im, keep = abi.image_u8(np.ascontiguousarray(image, np.uint8))
    out = np.zeros((keep.shape[0] // 2, keep.shape[1] // 2), np.uint8)
    check(h, lib().sva_resize_half_u8(h, C.byref(im), _p(out, C.c_uint8)))
    return out


def _yaml_write(filename, name, image):
    a = np.ascontiguousarray(image)
    if a.ndim != 2 or a.dtype not in (np.uint8, np.float64):
        raise CvException("only single-channel u8 / f64 matrices are supported")
    rc = lib().sva_yaml_write_matrix(str(filename).encode(), name.encode(), a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1], 0 if a.dtype == np.uint8 else 1)
    if rc != 0:
        raise CvException("cannot write %s" % filename)


def _yaml_read(filename, name):
    r, c, t = C.c_int32(), C.c_int32(), C.c_int32()
    if lib().sva_yaml_read_matrix(str(filename).encode(), name.encode(), None, C.c_int64(0), C.byref(r), C.byref(c), C.byref(t)) != 0:
        return np.zeros((0, 0), np.uint8)  # the reference returns an empty Mat for a missing file / key
    out = np.zeros((r.value, c.value), np.uint8 if t.value == 0 else np.float64)
    if lib().sva_yaml_read_matrix(str(filename).encode(), name.encode(), out.ctypes.data_as(C.c_void_p), C.c_int64(out.nbytes), C.byref(r), C.byref(c), C.byref(t)) != 0:
        raise CvException("malformed matrix in %s" % filename)
    return out


def saveImage(filename, image):
    """include/functions.h:49, src/functions.cpp:332-338 — cv::FileStorage YAML under the key `image`"""
    _yaml_write(filename, "image", image)


def loadImage(filename):
    """include/functions.h:51, src/functions.cpp:340-346"""
    return _yaml_read(filename, "image")


def getIdealRef(filename="idealRef.yml"):
    """include/functions.h:47, src/functions.cpp:323-330 — the ground-truth depth map under the key `R`"""
    return _yaml_read(filename, "R")


def getImagesPathsFromFolder(folderPath):
    """include/functions.h:43, src/functions.cpp:241-251 — directory order, like std::filesystem::directory_iterator"""
    import os
    with os.scandir(folderPath) as it:
        return [e.path for e in it]
