"""TEST INFRASTRUCTURE ONLY — numpy-facing ctypes wrappers over

  * oracle/_build/libsva_oracle.so  (our CPU restatement, oracle/sva_oracle.c)            -> class Oracle
  * oracle/_ref/libsva_ref.so       (the reference's own sources, compiled against cvshim)  -> class Reference

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes as C
import os
import tempfile

import numpy as np

from stereovisionarray_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libsva_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsva_ref.so")

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _cam5(c):
    return np.array([c.pos[0], c.pos[1], c.pos[2], c.f, c.pixel_size], dtype=np.float64)


class Oracle:
    def __init__(self, build=True):
        if build and (not os.path.exists(ORACLE_SO) or os.path.exists(os.path.join(HERE, "sva_oracle.c"))):
            from oracle import build_oracle
            build_oracle.build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_abs_diff.restype = C.c_double
        L.orc_abs_diff.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.orc_raw_cost_cell.restype = C.c_uint32

    # ---- literal ----
    def camera_project(self, cam, p):
        out = (C.c_int32 * 2)()
        self.lib.orc_camera_project(C.byref(cam), (C.c_double * 3)(*p), out)
        return int(out[0]), int(out[1])

    def camera_inv_project(self, cam, px):
        out = (C.c_double * 3)()
        self.lib.orc_camera_inv_project(C.byref(cam), (C.c_int32 * 2)(*px), out)
        return np.array(out[:], dtype=np.float64)

    def bresenham(self, a, b):
        cap = 2 * (abs(a[0] - b[0]) + abs(a[1] - b[1])) + 8
        out = np.zeros((cap, 2), dtype=np.int32)
        n = self.lib.orc_bresenham(a[0], a[1], b[0], b[1], _ptr(out, C.c_int32), cap)
        return out[:n].copy()

    def get_camera_pairs(self, n_cameras, pair_type, camera_num=-1):
        out = np.zeros((64, 2), dtype=np.int32)
        n = self.lib.orc_get_camera_pairs(n_cameras, pair_type, camera_num, _ptr(out, C.c_int32), 64)
        return out[:n].copy()

    def abs_diff(self, a, b):
        assert a.shape == b.shape and a.dtype == np.uint8 and b.dtype == np.uint8
        return float(self.lib.orc_abs_diff(a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0], a.shape[1], a.shape[0]))

    def match_literal(self, images, cams, pairs, mask, k=20, ray_near=0.5, ray_far=1.0, n_threads=1):
        imgs, keep = abi.image_array(images)
        cam_arr = abi.camera_array(cams)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        H, W = keep[0].shape
        out = np.zeros((H, W), dtype=np.uint8)
        m = None
        if mask is not None:
            m, mk = abi.image_u8(np.ascontiguousarray(mask, dtype=np.uint8))
        rc = self.lib.orc_match_literal(imgs, cam_arr, len(keep), _ptr(pairs, C.c_int32), len(pairs), C.byref(m) if m is not None else None,
                                        k, C.c_double(ray_near), C.c_double(ray_far), _ptr(out, C.c_uint8), n_threads)
        if rc != 0:
            raise RuntimeError("orc_match_literal rc=%d" % rc)
        return out

    def shift_perspective_with_disparity(self, in_cam, out_cam, disparity, image):
        d, dk = abi.image_u8(np.ascontiguousarray(disparity, dtype=np.uint8))
        i, ik = abi.image_u8(np.ascontiguousarray(image, dtype=np.uint8))
        out = np.zeros(ik.shape, dtype=np.uint8)
        self.lib.orc_shift_perspective_with_disparity(C.byref(in_cam), C.byref(out_cam), C.byref(d), C.byref(i), _ptr(out, C.c_uint8))
        return out

    def improve_with_disparity(self, disparity, center, images, cam_pairs, mask, window_size=21):
        d, dk = abi.image_u8(np.ascontiguousarray(disparity, dtype=np.uint8))
        c, ck = abi.image_u8(np.ascontiguousarray(center, dtype=np.uint8))
        imgs, keep = abi.image_array(images)
        cams = abi.camera_array([x for pr in cam_pairs for x in pr])
        m, mk = abi.image_u8(np.ascontiguousarray(mask, dtype=np.uint8))
        out = np.zeros(ck.shape, dtype=np.uint8)
        rc = self.lib.orc_improve_with_disparity(C.byref(d), C.byref(c), imgs, cams, len(keep), C.byref(m), window_size, _ptr(out, C.c_uint8))
        return rc, out

    def disparity_to_depth(self, disparity, baseline, f, pixel_size):
        d, dk = abi.image_u8(np.ascontiguousarray(disparity, dtype=np.uint8))
        out = np.zeros(dk.shape, dtype=np.float64)
        self.lib.orc_disparity_to_depth(C.byref(d), C.c_double(baseline), C.c_double(f), C.c_double(pixel_size), _ptr(out, C.c_double))
        return out

    # ---- consumers of the depth output ----
    def shift_perspective2(self, in_cam, out_cam, depth):
        d = np.ascontiguousarray(depth, dtype=np.float64)
        out = np.zeros_like(d)
        self.lib.orc_shift_perspective2(C.byref(in_cam), C.byref(out_cam), _ptr(d, C.c_double), d.shape[0], d.shape[1], _ptr(out, C.c_double))
        return out

    def points3d_to_depth_map(self, points, cam, width, height):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((height, width), dtype=np.float64)
        self.lib.orc_points3d_to_depth_map(_ptr(pts, C.c_double), C.c_longlong(len(pts)), C.byref(cam), width, height, _ptr(out, C.c_double))
        return out

    def depth_map_to_points3d(self, depth, cam, width, height):
        d = np.ascontiguousarray(depth, dtype=np.float64)
        out = np.zeros((d.size, 3), dtype=np.float64)
        self.lib.orc_depth_map_to_points3d.restype = C.c_longlong
        n = self.lib.orc_depth_map_to_points3d(_ptr(d, C.c_double), d.shape[0], d.shape[1], C.byref(cam), width, height, _ptr(out, C.c_double), C.c_longlong(d.size))
        return out[:n].copy()

    def get_groups(self, n_cameras, group_type="CHESS"):
        pairs = np.zeros((256, 2), dtype=np.int32)
        sizes = np.zeros(64, dtype=np.int32)
        ng = self.lib.orc_get_groups(n_cameras, group_type.encode(), _ptr(pairs, C.c_int32), 256, _ptr(sizes, C.c_int32), 64)
        out, o = [], 0
        for g in range(ng):
            out.append(pairs[o:o + sizes[g]].copy())
            o += sizes[g]
        return out

    # ---- volume ----
    def _shape(self, p):
        return (p.height, p.width, p.num_disp)

    def census_transform(self, img):
        i, ik = abi.image_u8(np.ascontiguousarray(img, dtype=np.uint8))
        out = np.zeros(ik.shape, dtype=np.uint64)
        self.lib.orc_census_transform(C.byref(i), _ptr(out, C.c_uint64))
        return out

    def ad_volume(self, p, ref, others, pair_begin=0, pair_end=None):
        r, rk = abi.image_u8(np.ascontiguousarray(ref, dtype=np.uint8))
        o, ok = abi.image_array(others)
        A = np.zeros(self._shape(p), dtype=np.uint16)
        rc = self.lib.orc_ad_volume(C.byref(p), C.byref(r), o, pair_begin, p.n_pairs if pair_end is None else pair_end, _ptr(A, C.c_uint16))
        assert rc == 0, rc
        return A

    def box_cost(self, p, A, raw=False):
        A = np.ascontiguousarray(A, dtype=np.uint16)
        C16 = np.zeros(self._shape(p), dtype=np.uint16)
        C32 = np.zeros(self._shape(p), dtype=np.uint32) if raw else None
        rc = self.lib.orc_box_cost(C.byref(p), _ptr(A, C.c_uint16), _ptr(C16, C.c_uint16), _ptr(C32, C.c_uint32) if raw else None)
        assert rc == 0, rc
        return (C16, C32) if raw else C16

    def raw_cost_cell(self, p, ref, others, y, x, d):
        r, rk = abi.image_u8(np.ascontiguousarray(ref, dtype=np.uint8))
        o, ok = abi.image_array(others)
        return int(self.lib.orc_raw_cost_cell(C.byref(p), C.byref(r), o, y, x, d))

    def cell_valid(self, p, y, x, d):
        return bool(self.lib.orc_cell_valid(C.byref(p), y, x, d))

    def sgm_single_path(self, p, Cv, dir_index):
        Cv = np.ascontiguousarray(Cv, dtype=np.uint16)
        L = np.zeros(self._shape(p), dtype=np.uint16)
        rc = self.lib.orc_sgm_single_path(C.byref(p), _ptr(Cv, C.c_uint16), dir_index, _ptr(L, C.c_uint16))
        assert rc == 0, rc
        return L

    def sgm_rows(self, p, Cv, dir_index, y0, rows, prev_in, S, prev_out):
        """one row-sweeping direction on image rows [y0, y0 + rows), added into S in place; prev_in / prev_out: [W][D] u16 or None"""
        Cv = np.ascontiguousarray(Cv, dtype=np.uint16)
        assert S.dtype == np.uint16 and S.flags.c_contiguous
        rc = self.lib.orc_sgm_rows(C.byref(p), _ptr(Cv, C.c_uint16), dir_index, y0, rows, _ptr(prev_in, C.c_uint16) if prev_in is not None else None,
                                   _ptr(S, C.c_uint16), _ptr(prev_out, C.c_uint16) if prev_out is not None else None)
        assert rc == 0, rc

    def sgm_aggregate(self, p, Cv, n_use=None):
        Cv = np.ascontiguousarray(Cv, dtype=np.uint16)
        S = np.zeros(self._shape(p), dtype=np.uint16)
        rc = self.lib.orc_sgm_aggregate(C.byref(p), _ptr(Cv, C.c_uint16), p.n_paths if n_use is None else n_use, _ptr(S, C.c_uint16))
        assert rc == 0, rc
        return S

    def wta(self, p, S, mask=None):
        S = np.ascontiguousarray(S, dtype=np.uint16)
        disp = np.zeros((p.height, p.width), dtype=np.uint16)
        sub = np.zeros((p.height, p.width), dtype=np.float32)
        m = None
        if mask is not None:
            m, mk = abi.image_u8(np.ascontiguousarray(mask, dtype=np.uint8))
        rc = self.lib.orc_wta(C.byref(p), _ptr(S, C.c_uint16), C.byref(m) if m is not None else None, _ptr(disp, C.c_uint16), _ptr(sub, C.c_float))
        assert rc == 0, rc
        return disp, sub

    def depth_from_array(self, p, ref, others, mask=None, want_volumes=False):
        r, rk = abi.image_u8(np.ascontiguousarray(ref, dtype=np.uint8))
        o, ok = abi.image_array(others)
        disp = np.zeros((p.height, p.width), dtype=np.uint16)
        sub = np.zeros((p.height, p.width), dtype=np.float32)
        m = None
        if mask is not None:
            m, mk = abi.image_u8(np.ascontiguousarray(mask, dtype=np.uint8))
        vols = [np.zeros(self._shape(p), dtype=np.uint16) for _ in range(3)] if want_volumes else [None] * 3
        rc = self.lib.orc_depth_from_array(C.byref(p), C.byref(r), o, C.byref(m) if m is not None else None,
                                           *[_ptr(v, C.c_uint16) if v is not None else None for v in vols],
                                           _ptr(disp, C.c_uint16), _ptr(sub, C.c_float))
        assert rc == 0, rc
        return (disp, sub, vols) if want_volumes else (disp, sub)

    def num_threads(self):
        return int(self.lib.orc_num_threads())


class Reference:
    """The reference's own code (compiled from /root/reference in the build container; prebuilt .so elsewhere)."""

    @staticmethod
    def available():
        if os.path.exists("/root/reference/src/functions.cpp"):
            from oracle import build_oracle
            build_oracle.build()
        return os.path.exists(REF_SO)

    def __init__(self):
        if not Reference.available():
            raise RuntimeError("oracle/_ref/libsva_ref.so is not built (needs /root/reference)")
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_get_abs_diff.restype = C.c_double
        self.lib.ref_last_error.restype = C.c_char_p

    def camera_project(self, cam, p):
        px, py = C.c_int(), C.c_int()
        self.lib.ref_camera_project(_ptr(_cam5(cam), C.c_double), C.c_double(p[0]), C.c_double(p[1]), C.c_double(p[2]), C.byref(px), C.byref(py))
        return px.value, py.value

    def camera_inv_project(self, cam, px):
        out = np.zeros(3, dtype=np.float64)
        self.lib.ref_camera_inv_project(_ptr(_cam5(cam), C.c_double), int(px[0]), int(px[1]), _ptr(out, C.c_double))
        return out

    def bresenham(self, a, b):
        cap = 2 * (abs(a[0] - b[0]) + abs(a[1] - b[1])) + 8
        out = np.zeros((cap, 2), dtype=np.int32)
        n = self.lib.ref_bresenham(a[0], a[1], b[0], b[1], _ptr(out, C.c_int32), cap)
        return out[:n].copy()

    def shift_perspective2(self, in_cam, out_cam, depth):
        d = np.ascontiguousarray(depth, dtype=np.float64)
        out = np.zeros_like(d)
        rc = self.lib.ref_shift_perspective2(_ptr(_cam5(in_cam), C.c_double), _ptr(_cam5(out_cam), C.c_double), _ptr(d, C.c_double), d.shape[1], d.shape[0], _ptr(out, C.c_double))
        assert rc == 0, self.lib.ref_last_error()
        return out

    def points3d_to_depth_map(self, points, cam, width, height):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((height, width), dtype=np.float64)
        rc = self.lib.ref_points3d_to_depth_map(_ptr(pts, C.c_double), C.c_longlong(len(pts)), _ptr(_cam5(cam), C.c_double), width, height, _ptr(out, C.c_double))
        assert rc == 0, self.lib.ref_last_error()
        return out

    def depth_map_to_points3d(self, depth, cam, width, height):
        d = np.ascontiguousarray(depth, dtype=np.float64)
        out = np.zeros((d.size, 3), dtype=np.float64)
        self.lib.ref_depth_map_to_points3d.restype = C.c_longlong
        n = self.lib.ref_depth_map_to_points3d(_ptr(d, C.c_double), d.shape[1], d.shape[0], _ptr(_cam5(cam), C.c_double), width, height, _ptr(out, C.c_double), C.c_longlong(d.size))
        assert n >= 0, self.lib.ref_last_error()
        return out[:n].copy()

    def get_groups(self, n_cameras, group_type="CHESS"):
        pairs = np.zeros((256, 2), dtype=np.int32)
        sizes = np.zeros(64, dtype=np.int32)
        ng = self.lib.ref_get_groups(n_cameras, group_type.encode(), _ptr(pairs, C.c_int32), 256, _ptr(sizes, C.c_int32), 64)
        out, o = [], 0
        for g in range(ng):
            out.append(pairs[o:o + sizes[g]].copy())
            o += sizes[g]
        return out

    def get_camera_pairs(self, n_cameras, pair_type, camera_num=-1):
        out = np.zeros((64, 2), dtype=np.int32)
        n = self.lib.ref_get_camera_pairs(n_cameras, pair_type, camera_num, _ptr(out, C.c_int32), 64)
        return out[:n].copy()

    def abs_diff_roi(self, a, ax, ay, b, bx, by, w, h):
        a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
        return float(self.lib.ref_get_abs_diff(_ptr(a, C.c_uint8), a.shape[1], a.shape[0], ax, ay, _ptr(b, C.c_uint8), b.shape[1], b.shape[0], bx, by, w, h))

    def shift_perspective_with_disparity(self, in_cam, out_cam, disparity, image):
        d = np.ascontiguousarray(disparity, dtype=np.uint8); im = np.ascontiguousarray(image, dtype=np.uint8)
        out = np.zeros(im.shape, dtype=np.uint8)
        rc = self.lib.ref_shift_perspective_with_disparity(_ptr(_cam5(in_cam), C.c_double), _ptr(_cam5(out_cam), C.c_double),
                                                           _ptr(d, C.c_uint8), _ptr(im, C.c_uint8), im.shape[1], im.shape[0], _ptr(out, C.c_uint8))
        assert rc == 0, self.lib.ref_last_error()
        return out

    def improve_with_disparity(self, disparity, center, images, cam_pairs, mask, window_size=21):
        d = np.ascontiguousarray(disparity, dtype=np.uint8); c = np.ascontiguousarray(center, dtype=np.uint8)
        imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in images]
        ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
        cams = np.concatenate([np.concatenate([_cam5(a), _cam5(b)]) for a, b in cam_pairs]).astype(np.float64)
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        out = np.zeros(c.shape, dtype=np.uint8)
        rc = self.lib.ref_improve_with_disparity(_ptr(d, C.c_uint8), _ptr(c, C.c_uint8), ptrs, _ptr(cams, C.c_double), len(imgs),
                                                 _ptr(m, C.c_uint8), c.shape[1], c.shape[0], window_size, _ptr(out, C.c_uint8))
        return rc, out

    def main_run(self, images25_2x, mask, ideal_ref=None):
        """Runs the reference driver (main) on 25 images of size (2H, 2W); returns (disparity u8, depth f64, improved u8), each (H, W)."""
        imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in images25_2x]
        assert len(imgs) == 25
        h2, w2 = imgs[0].shape
        h, w = h2 // 2, w2 // 2
        ptrs = (C.c_void_p * 25)(*[i.ctypes.data for i in imgs])
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        ideal = np.ascontiguousarray(ideal_ref if ideal_ref is not None else np.ones((h, w)), dtype=np.float64)
        disp = np.zeros((h, w), dtype=np.uint8); imp = np.zeros((h, w), dtype=np.uint8); dep = np.zeros((h, w), dtype=np.float64)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "Renders2"))
            for i in range(25):
                open(os.path.join(td, "Renders2", "cam%02d.png" % i), "wb").close()
            try:
                rc = self.lib.ref_main_run(td.encode(), ptrs, w2, h2, _ptr(m, C.c_uint8), _ptr(ideal, C.c_double),
                                           _ptr(disp, C.c_uint8), _ptr(dep, C.c_double), _ptr(imp, C.c_uint8))
            finally:
                os.chdir(cwd)
        if rc != 0:
            raise RuntimeError("reference main failed: %s" % self.lib.ref_last_error().decode())
        return disp, dep, imp
