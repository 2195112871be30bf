// sva_io.cu — ingest side of the path: the x0.5 resize the driver applies to every capture
// (src/CameraStereoVision.cpp:18: resize(img, img, Size(), 0.5, 0.5), default INTER_LINEAR) as a GPU pre-pass, and the
// cv::FileStorage YAML matrices the reference saves / loads (saveImage / loadImage / getIdealRef — src/functions.cpp:323-346).
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sva_common.cuh"

// For an exact 2x decimation OpenCV's INTER_LINEAR takes its "area fast" path: the rounded mean of each 2x2 block, (a+b+c+d+2)>>2.
// (Checked against cv2.resize in tests/test_io.py.  Odd sizes go through OpenCV's general fixed-point bilinear code, whose rounding
// differs between OpenCV versions, so they are rejected rather than approximated.)
__global__ void k_resize_half(const uint8_t* __restrict__ src, size_t pitch, int dw, int dh, uint8_t* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const uint8_t* r0 = src + (size_t)(2 * y) * pitch + 2 * x;
    const uint8_t* r1 = r0 + pitch;
    dst[(size_t)y * dw + x] = (uint8_t)(((int)r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2);
}

namespace {

const char* skip_ws(const char* p) {
    while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') p++;
    return p;
}

// value of "   key: <int>" inside the matrix node that starts at `node`
bool field_int(const char* node, const char* end, const char* key, long& out) {
    const std::string k = std::string(key) + ":";
    const char* p = std::strstr(node, k.c_str());
    if (!p || p >= end) return false;
    out = std::strtol(p + k.size(), nullptr, 10);
    return true;
}

}  // namespace

extern "C" {

int sva_resize_half_u8(sva_ctx* c, const sva_image_u8* img, uint8_t* out) {
    if (!c || !img || !img->data || !out) return c ? c->fail(SVA_ERR_BAD_ARG, "resize_half: null argument") : SVA_ERR_BAD_ARG;
    if (img->rows < 2 || img->cols < 2 || (img->rows & 1) || (img->cols & 1) || img->step < (size_t)img->cols)
        return c->fail(SVA_ERR_BAD_ARG, "resize_half: rows and cols must be even (exact 2x decimation)");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const int W = img->cols, H = img->rows, dw = W / 2, dh = H / 2;
    SVA_TRY(c->reserve(c->scratch, (size_t)W * H + (size_t)dw * dh));
    uint8_t* d_src = c->scratch.as<uint8_t>();
    uint8_t* d_dst = d_src + (size_t)W * H;
    SVA_CUDA_OK(c, cudaMemcpy2DAsync(d_src, W, img->data, img->step, W, H, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, "k_resize_half");
        k_resize_half<<<dim3(div_up(dw, 128), dh), 128, 0, c->stream>>>(d_src, (size_t)W, dw, dh, d_dst);
    }
    SVA_CUDA_OK(c, cudaGetLastError());
    SVA_CUDA_OK(c, cudaMemcpyAsync(out, d_dst, (size_t)dw * dh, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}

/* cv::FileStorage YAML, one single-channel matrix per file under `name` ("image" for saveImage / loadImage, "R" for getIdealRef).
 * dtype: 0 = u8 ("dt: u"), 1 = f64 ("dt: d").  Files written here load in OpenCV and vice versa. */
int sva_yaml_write_matrix(const char* path, const char* name, const void* data, int32_t rows, int32_t cols, int32_t dtype) {
    if (!path || !name || !data || rows < 0 || cols < 0 || (dtype != 0 && dtype != 1)) return SVA_ERR_BAD_ARG;
    FILE* f = std::fopen(path, "w");
    if (!f) return SVA_ERR_BAD_ARG;
    std::fprintf(f, "%%YAML:1.0\n---\n%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: %c\n   data: [ ", name, rows, cols, dtype == 0 ? 'u' : 'd');
    const size_t n = (size_t)rows * cols;
    int col = 11;
    for (size_t i = 0; i < n; i++) {
        char buf[40];
        int len;
        if (dtype == 0) len = std::snprintf(buf, sizeof buf, "%u", (unsigned)((const uint8_t*)data)[i]);
        else {
            const double v = ((const double*)data)[i];
            if (std::isnan(v)) len = std::snprintf(buf, sizeof buf, ".Nan");
            else if (std::isinf(v)) len = std::snprintf(buf, sizeof buf, v < 0 ? "-.Inf" : ".Inf");
            else {
                len = std::snprintf(buf, sizeof buf, "%.17g", v);
                if (!std::strpbrk(buf, ".eE")) { buf[len++] = '.'; buf[len] = 0; }  // OpenCV marks reals with a '.'
            }
        }
        if (col + len + 2 > 76 && i > 0) { std::fputs("\n       ", f); col = 7; }
        std::fputs(buf, f);
        if (i + 1 < n) std::fputs(", ", f);
        col += len + 2;
    }
    std::fputs(" ]\n", f);
    return std::fclose(f) == 0 ? SVA_OK : SVA_ERR_BAD_ARG;
}

/* out may be NULL to query rows / cols / dtype; otherwise at most cap_bytes are written (SVA_ERR_BAD_ARG if the matrix is larger). */
int sva_yaml_read_matrix(const char* path, const char* name, void* out, int64_t cap_bytes, int32_t* rows, int32_t* cols, int32_t* dtype) {
    if (!path || !name || !rows || !cols || !dtype) return SVA_ERR_BAD_ARG;
    FILE* f = std::fopen(path, "rb");
    if (!f) return SVA_ERR_BAD_ARG;
    std::string text;
    char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
    std::fclose(f);
    const std::string key = std::string(name) + ":";
    size_t pos = 0;
    const char* node = nullptr;
    while ((pos = text.find(key, pos)) != std::string::npos) {  // the key must start a line
        if (pos == 0 || text[pos - 1] == '\n') { node = text.c_str() + pos; break; }
        pos += key.size();
    }
    if (!node || !std::strstr(node, "!!opencv-matrix")) return SVA_ERR_BAD_ARG;
    const char* data = std::strstr(node, "data:");
    if (!data) return SVA_ERR_BAD_ARG;
    long r = 0, cc = 0;
    if (!field_int(node, data, "rows", r) || !field_int(node, data, "cols", cc) || r < 0 || cc < 0) return SVA_ERR_BAD_ARG;
    if (r > INT32_MAX || cc > INT32_MAX) return SVA_ERR_BAD_ARG;  // untrusted input: the shape must survive the int32 outputs ...
    const char* dt = std::strstr(node, "dt:");
    if (!dt || dt >= data) return SVA_ERR_BAD_ARG;
    dt = skip_ws(dt + 3);
    if (*dt == '"') dt++;
    int type;
    if (*dt == 'u') type = 0; else if (*dt == 'd') type = 1; else return SVA_ERR_BAD_ARG;  // single-channel u8 / f64 only
    *rows = (int32_t)r; *cols = (int32_t)cc; *dtype = type;
    if (!out) return SVA_OK;
    const size_t n = (size_t)r * cc, esz = type == 0 ? 1 : 8;  // ... so n < 2^62 and the product below cannot wrap past the check
    if (cap_bytes < 0 || n > (size_t)INT64_MAX / 8 || (int64_t)(n * esz) > cap_bytes) return SVA_ERR_BAD_ARG;
    const char* p = std::strchr(data, '[');
    if (!p) return SVA_ERR_BAD_ARG;
    p++;
    for (size_t i = 0; i < n; i++) {
        p = skip_ws(p);
        if (*p == ',') p = skip_ws(p + 1);
        if (*p == ']' || *p == 0) return SVA_ERR_BAD_ARG;  // fewer values than rows * cols
        double v;
        if (!std::strncmp(p, ".Inf", 4) || !std::strncmp(p, ".inf", 4)) { v = INFINITY; p += 4; }
        else if (!std::strncmp(p, "-.Inf", 5) || !std::strncmp(p, "-.inf", 5)) { v = -INFINITY; p += 5; }
        else if (!std::strncmp(p, ".Nan", 4) || !std::strncmp(p, ".NaN", 4) || !std::strncmp(p, ".nan", 4)) { v = NAN; p += 4; }
        else {
            char* e = nullptr;
            v = std::strtod(p, &e);
            if (e == p) return SVA_ERR_BAD_ARG;
            p = e;
        }
        if (type == 0) ((uint8_t*)out)[i] = (uint8_t)(v >= 0.0 && v <= 255.0 ? v : (v > 255.0 ? 255.0 : 0.0));  // NaN / out of range: defined, clamped
        else ((double*)out)[i] = v;
    }
    return SVA_OK;
}

}  // extern "C"
