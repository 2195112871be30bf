// TEST INFRASTRUCTURE ONLY — see opencv2/core.hpp in this directory.
// resize(): only used by the reference's driver for the x0.5 preprocess and the
// (GUI-only) error display; it is OFF the parity path.  The shim implements the
// exact 2x2 block mean that INTER_LINEAR degenerates to at scale 0.5, and
// nearest sampling for any other target size.
#pragma once
#include "core.hpp"
namespace cv {
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
inline void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int = INTER_LINEAR) {
    int w = dsize.width, h = dsize.height;
    if (w == 0 || h == 0) { w = cvRound(src.cols * fx); h = cvRound(src.rows * fy); }
    if (w <= 0 || h <= 0) throw Exception("cvshim: resize to an empty size");
    Mat out(h, w, src.type());
    if (2 * w == src.cols && 2 * h == src.rows) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                out.put(y, x, (src.get(2 * y, 2 * x) + src.get(2 * y, 2 * x + 1) + src.get(2 * y + 1, 2 * x) + src.get(2 * y + 1, 2 * x + 1)) / 4.0);
    } else {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                out.put(y, x, src.get(std::min(src.rows - 1, (int)((y + 0.5) * src.rows / h)), std::min(src.cols - 1, (int)((x + 0.5) * src.cols / w))));
    }
    dst = out;
}
}  // namespace cv
