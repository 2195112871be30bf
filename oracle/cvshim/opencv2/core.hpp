// TEST INFRASTRUCTURE ONLY (oracle/).  Minimal stand-in for the handful of
// OpenCV 4.2 types and functions that the reference's depth path touches, so
// that the UNMODIFIED reference sources (/root/reference/src/Camera.cpp,
// functions.cpp, CameraStereoVision.cpp) compile and run in a container that
// has no OpenCV C++ SDK.  Nothing here is shipped or linked into the product
// library; the build recipe is oracle/build_oracle.py.
//
// Semantics restated from the OpenCV 4.x public documentation / types.hpp:
//  * Rect(Point,Point) is half-open; Mat::operator()(Rect) is a view and
//    asserts the ROI lies inside the parent.
//  * abs(A - B) on Mat folds to absdiff (true |a-b|, not a saturating sub).
//  * sum() returns an exact per-channel total in double.
//  * norm(Point) = sqrt(x*x + y*y [+ z*z]) evaluated left to right in double.
//  * Point3d / double and * double act per component (real division).
//  * Point_<int>(Point_<double>) rounds with cvRound (round-half-even).
//  * Mat::at<uchar>() = int is a plain C++ narrowing.
//  * new Mat buffers are zero-filled here (OpenCV leaves them uninitialised);
//    parity is only ever asserted on pixels the reference writes.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cv {

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_16UC1 2
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_64FC1 6

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

inline int cvRound(double v) { return (int)std::nearbyint(v); }  // round-half-even in default FP mode
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <typename T> inline T saturate_cast(int v) { return (T)v; }

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
template <typename T> inline Size_<T> operator/(const Size_<T>& a, T b) { return Size_<T>(a.width / b, a.height / b); }
template <typename T> inline std::ostream& operator<<(std::ostream& os, const Size_<T>& s) {
    return os << "[" << s.width << " x " << s.height << "]";
}
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;
typedef Size2i Size;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_(const Size_<T>& s) : x(s.width), y(s.height) {}
    template <typename U> operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, int b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> inline std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
template <typename T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> inline Point3_<T> operator*(const Point3_<T>& a, double b) { return Point3_<T>((T)(a.x * b), (T)(a.y * b), (T)(a.z * b)); }
template <typename T> inline Point3_<T> operator/(const Point3_<T>& a, double b) { return Point3_<T>((T)(a.x / b), (T)(a.y / b), (T)(a.z / b)); }
template <typename T> inline std::ostream& operator<<(std::ostream& os, const Point3_<T>& p) { return os << "[" << p.x << ", " << p.y << ", " << p.z << "]"; }
typedef Point3_<double> Point3d;
typedef Point3_<int> Point3i;

template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
template <typename T> inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }
inline double norm(double v) { return std::fabs(v); }  // norm(InputArray(double)) == L2 norm of a 1x1 matrix

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T>& a, const Point_<T>& b) {
        x = std::min(a.x, b.x); y = std::min(a.y, b.y);
        width = std::max(a.x, b.x) - x; height = std::max(a.y, b.y) - y;
    }
};
typedef Rect_<int> Rect;

struct Scalar {
    double val[4];
    Scalar() : val{0, 0, 0, 0} {}
    Scalar(double a) : val{a, 0, 0, 0} {}
    double operator[](int i) const { return val[i]; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;

    Mat() {}
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type, const Scalar& v) { create(s.height, s.width, type); setTo(v.val[0]); }
    template <typename T> Mat(const std::initializer_list<T> l) {
        create(1, (int)l.size(), CV_32S);
        int i = 0;
        for (auto v : l) ((int*)data)[i++] = (int)v;
    }
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[std::max<size_t>(1, step * (size_t)r)](), std::default_delete<uchar[]>());
        data = buf_.get();
    }
    size_t elemSize() const {
        static const size_t sz[7] = {1, 1, 2, 2, 4, 4, 8};
        return sz[type_];
    }
    int type() const { return type_; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void setTo(double v) {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) put(y, x, v);
    }
    double get(int y, int x) const {
        const uchar* p = data + (size_t)y * step + (size_t)x * elemSize();
        switch (type_) {
            case CV_8U: return *p;
            case CV_8S: return *(const signed char*)p;
            case CV_16U: return *(const uint16_t*)p;
            case CV_16S: return *(const int16_t*)p;
            case CV_32S: return *(const int32_t*)p;
            case CV_32F: return *(const float*)p;
            default: return *(const double*)p;
        }
    }
    void put(int y, int x, double v) {
        uchar* p = data + (size_t)y * step + (size_t)x * elemSize();
        switch (type_) {
            case CV_8U: *p = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(v))); break;
            case CV_16U: *(uint16_t*)p = (uint16_t)std::min(65535.0, std::max(0.0, std::nearbyint(v))); break;
            case CV_32S: *(int32_t*)p = (int32_t)std::nearbyint(v); break;
            case CV_32F: *(float*)p = (float)v; break;
            case CV_64F: *(double*)p = v; break;
            default: throw Exception("cvshim: unsupported type in put()");
        }
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    Mat operator()(const Rect& r) const {
        if (!(0 <= r.x && 0 <= r.width && r.x + r.width <= cols && 0 <= r.y && 0 <= r.height && r.y + r.height <= rows))
            throw Exception("cvshim: ROI outside of the matrix (Mat::operator()(Rect) assertion)");
        Mat m;
        m.type_ = type_; m.rows = r.height; m.cols = r.width; m.step = step; m.buf_ = buf_;
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }

private:
    int type_ = CV_8U;
    std::shared_ptr<uchar> buf_;
};

// --- the few matrix expressions the path uses -------------------------------------------
struct MatSubExpr { Mat a, b; };
inline MatSubExpr operator-(const Mat& a, const Mat& b) {
    if (!(a.size() == b.size()) || a.type() != b.type()) throw Exception("cvshim: size/type mismatch in A - B");
    return MatSubExpr{a, b};
}
inline Mat abs(const MatSubExpr& e) {  // folds to absdiff()
    Mat r(e.a.rows, e.a.cols, e.a.type());
    if (e.a.type() == CV_8U) {  // the hot call (getAbsDiff): a tight row loop the compiler vectorises, like OpenCV's SIMD absdiff
        for (int y = 0; y < r.rows; y++) {
            const uchar* pa = e.a.data + (size_t)y * e.a.step; const uchar* pb = e.b.data + (size_t)y * e.b.step;
            uchar* pr = r.data + (size_t)y * r.step;
            for (int x = 0; x < r.cols; x++) pr[x] = pa[x] > pb[x] ? (uchar)(pa[x] - pb[x]) : (uchar)(pb[x] - pa[x]);
        }
        return r;
    }
    for (int y = 0; y < r.rows; y++)
        for (int x = 0; x < r.cols; x++) r.put(y, x, std::fabs(e.a.get(y, x) - e.b.get(y, x)));
    return r;
}
inline Mat operator*(const MatSubExpr& e, double s) {
    Mat r(e.a.rows, e.a.cols, e.a.type());
    for (int y = 0; y < r.rows; y++)
        for (int x = 0; x < r.cols; x++) r.put(y, x, (e.a.get(y, x) - e.b.get(y, x)) * s);
    return r;
}
inline Mat operator/(double s, const Mat& m) {  // IEEE division for floating types (OpenCV >= 4)
    Mat r(m.rows, m.cols, m.type());
    for (int y = 0; y < r.rows; y++)
        for (int x = 0; x < r.cols; x++) r.put(y, x, s / m.get(y, x));
    return r;
}
inline Scalar sum(const Mat& m) {
    if (m.type() == CV_8U) {  // exact integer total (OpenCV accumulates u8 in integers, then converts)
        unsigned long long t = 0;
        for (int y = 0; y < m.rows; y++) {
            const uchar* p = m.data + (size_t)y * m.step;
            unsigned int rs = 0;
            for (int x = 0; x < m.cols; x++) rs += p[x];
            t += rs;
        }
        return Scalar((double)t);
    }
    double s = 0;
    for (int y = 0; y < m.rows; y++)
        for (int x = 0; x < m.cols; x++) s += m.get(y, x);
    return Scalar(s);
}
inline Scalar mean(const Mat& m, const Mat& mask) {
    double s = 0; long n = 0;
    for (int y = 0; y < m.rows; y++)
        for (int x = 0; x < m.cols; x++)
            if (mask.get(y, x) != 0) { s += m.get(y, x); n++; }
    return Scalar(n ? s / n : 0.0);
}
inline void multiply(const Mat& a, double b, Mat& dst, double scale = 1, int dtype = -1) {
    Mat r(a.rows, a.cols, dtype < 0 ? a.type() : dtype);
    for (int y = 0; y < r.rows; y++)
        for (int x = 0; x < r.cols; x++) r.put(y, x, a.get(y, x) * b * scale);
    dst = r;
}

// --- persistence: an in-memory registry keyed "<file>:<key>" (the real YAML IO is host-side, out of scope)
inline std::map<std::string, Mat>& shim_registry() { static std::map<std::string, Mat> r; return r; }
class FileNode {
public:
    std::string key;
    void operator>>(Mat& m) const {
        auto it = shim_registry().find(key);
        m = (it == shim_registry().end()) ? Mat() : it->second.clone();
    }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string& f, int) : file_(f) {}
    bool open(const std::string& f, int) { file_ = f; return true; }
    FileNode operator[](const char* k) const { return FileNode{file_ + ":" + k}; }
    std::string file_, pending_;
};
inline FileStorage& operator<<(FileStorage& fs, const char* k) { fs.pending_ = k; return fs; }
inline FileStorage& operator<<(FileStorage& fs, const Mat& m) { shim_registry()[fs.file_ + ":" + fs.pending_] = m.clone(); return fs; }

}  // namespace cv
