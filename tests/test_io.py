"""Ingest side: the YAML matrices of saveImage / loadImage / getIdealRef round-trip with OpenCV's own cv::FileStorage in
both directions (host code of libsva_b200.so: runs without a GPU), and — on a B200 — the x0.5 resize pre-pass equals cv2.resize."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def test_yaml_matrices_round_trip_with_opencv(tmp_path):
    from stereovisionarray_b200 import reference_api as api
    rng = np.random.default_rng(4)
    u8 = rng.integers(0, 256, size=(37, 53), dtype=np.uint8)
    f64 = rng.random((21, 17)) * 3.0 - 1.0
    f64[0, 0], f64[0, 1], f64[0, 2], f64[1, 0], f64[1, 1] = 0.0, 1.0, -2.0, np.inf, 1e-300
    for name, a in (("u8", u8), ("f64", f64)):
        ours, theirs = str(tmp_path / (name + "_ours.yml")), str(tmp_path / (name + "_cv.yml"))
        api.saveImage(ours, a)                                   # written here ...
        fs = cv2.FileStorage(ours, cv2.FILE_STORAGE_READ)
        got = fs.getNode("image").mat()
        fs.release()
        assert got.dtype == a.dtype and np.array_equal(got, a)   # ... read by OpenCV
        fs = cv2.FileStorage(theirs, cv2.FILE_STORAGE_WRITE)
        fs.write("image", a)
        fs.write("R", a.T.copy())
        fs.release()
        assert np.array_equal(api.loadImage(theirs), a)          # written by OpenCV, read here
        assert np.array_equal(api.getIdealRef(theirs), a.T)
        assert np.array_equal(api.loadImage(ours), a)
    assert api.loadImage(str(tmp_path / "missing.yml")).size == 0
    assert api.getIdealRef(str(tmp_path / "u8_ours.yml")).size == 0  # no "R" key in that file
    assert sorted(os.path.basename(p) for p in api.getImagesPathsFromFolder(str(tmp_path))) == ["f64_cv.yml", "f64_ours.yml", "u8_cv.yml", "u8_ours.yml"]


@pytest.mark.gpu
def test_resize_half_equals_opencv():
    from stereovisionarray_b200 import reference_api as api
    from stereovisionarray_b200._lib import SvaError
    rng = np.random.default_rng(5)
    for h, w in [(8, 10), (480, 640), (1920, 2560)]:
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        assert np.array_equal(api.resizeHalf(img), cv2.resize(img, None, fx=0.5, fy=0.5))
    view = rng.integers(0, 256, size=(64, 100), dtype=np.uint8)[:, 10:74]  # a non-contiguous ROI goes through its step
    assert np.array_equal(api.resizeHalf(view), cv2.resize(np.ascontiguousarray(view), None, fx=0.5, fy=0.5))
    with pytest.raises(SvaError):
        api.resizeHalf(np.zeros((9, 10), np.uint8))
