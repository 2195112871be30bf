// k_cloud.cu — consumers of the depth output: shiftPerspective2, Points3DToDepthMap, DepthMapToPoints3D
// (reference include/functions.h:24,30,32; src/functions.cpp:79-104, 118-146), f64, bit-exact with the reference's serial loops.
//
// The two scatters are "last writer wins" in the reference's iteration order (x outer / y inner, resp. point index).  On the GPU the
// order is made explicit: pass 1 takes, per target pixel, the maximum order key of all sources that land on it (atomicMax), pass 2 lets
// exactly the source holding that key write.  DepthMapToPoints3D is a stream compaction in (u outer, v inner) order: flags, an
// exclusive prefix sum (cub), ordered scatter.  All arithmetic uses explicit round-to-nearest f64 intrinsics (no FMA contraction).
#include <cub/device/device_scan.cuh>

#include "sva_cam.cuh"
#include "sva_common.cuh"

// ---- shiftPerspective2 — src/functions.cpp:79-104 ----
__device__ __forceinline__ bool sp2_target(const double* __restrict__ depth, int rows, int cols, double pmx, double pmy, int x, int y, int& sx, int& sy, double& d) {
    d = depth[(size_t)y * cols + x];
    if (d < 0.5) return false;                                         // :89
    sx = (int)__ddiv_rn(pmx, d) + x;                                   // :91
    sy = (int)__ddiv_rn(pmy, d) + y;                                   // :92
    return !(sy >= rows || sy < 0 || sx >= cols || sx < 0);            // :93
}
__global__ void k_sp2_order(const double* __restrict__ depth, int rows, int cols, double pmx, double pmy, unsigned int* __restrict__ order) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    int sx, sy; double d;
    if (sp2_target(depth, rows, cols, pmx, pmy, x, y, sx, sy, d)) atomicMax(&order[(size_t)sy * cols + sx], (unsigned int)(x * rows + y) + 1u);  // :86-87 order
}
__global__ void k_sp2_write(const double* __restrict__ depth, int rows, int cols, double pmx, double pmy, const unsigned int* __restrict__ order,
                            double* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    int sx, sy; double d;
    if (sp2_target(depth, rows, cols, pmx, pmy, x, y, sx, sy, d) && order[(size_t)sy * cols + sx] == (unsigned int)(x * rows + y) + 1u)
        out[(size_t)sy * cols + sx] = d;                               // :95
}

// ---- Points3DToDepthMap — src/functions.cpp:118-133 ----
__device__ __forceinline__ bool p2d_target(const double* __restrict__ xyz, long long i, const DevCam& c, int W, int H, int& x, int& y) {
    int u, v;
    dev_project(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], u, v);  // :125
    x = u + W / 2; y = v + H / 2;                                      // halfRes = resolution / 2 (:122)
    return x >= 0 && x < W && y >= 0 && y < H;                         // :126
}
__global__ void k_p2d_order(const double* __restrict__ xyz, long long n, DevCam c, int W, int H, unsigned int* __restrict__ order) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y;
    if (p2d_target(xyz, i, c, W, H, x, y)) atomicMax(&order[(size_t)y * W + x], (unsigned int)i + 1u);
}
__global__ void k_p2d_write(const double* __restrict__ xyz, long long n, DevCam c, int W, int H, const unsigned int* __restrict__ order, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y;
    if (p2d_target(xyz, i, c, W, H, x, y) && order[(size_t)y * W + x] == (unsigned int)i + 1u)
        out[(size_t)y * W + x] = __dsub_rn(xyz[3 * i + 2], c.pz);     // :128
}

// ---- DepthMapToPoints3D — src/functions.cpp:135-146 ----
__global__ void k_d2p_flags(const double* __restrict__ depth, int rows, int cols, unsigned int* __restrict__ flags) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, u = blockIdx.y;  // element index u * rows + v: the reference's push_back order (:139-140)
    if (v >= rows) return;
    flags[(size_t)u * rows + v] = depth[(size_t)v * cols + u] > 0.1 ? 1u : 0u;  // :142
}
__global__ void k_d2p_scatter(const double* __restrict__ depth, int rows, int cols, DevCam c, int hx, int hy, const unsigned int* __restrict__ flags,
                              const unsigned int* __restrict__ pos, long long cap, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, u = blockIdx.y;
    if (v >= rows) return;
    const size_t e = (size_t)u * rows + v;
    if (!flags[e] || (long long)pos[e] >= cap) return;
    const double d = depth[(size_t)v * cols + u];
    double rx, ry, rz;
    dev_inv_project(c, u - hx, v - hy, rx, ry, rz);                    // :143
    double* o = out + 3 * (size_t)pos[e];
    o[0] = __dadd_rn(c.px, __dmul_rn(rx, d)); o[1] = __dadd_rn(c.py, __dmul_rn(ry, d)); o[2] = __dadd_rn(c.pz, __dmul_rn(rz, d));
}

// ---- calculateAverageError — src/functions.cpp:348-354: cv::mean(image, mask)[0] ----
// per-block partial sums in a fixed order (thread-strided, then a shared-memory tree), block partials added on the host in block order:
// deterministic, and within a few ulp of any other summation order (cv::mean's own blocking is not specified)
__global__ void k_masked_sum(const double* __restrict__ img, const uint8_t* __restrict__ mask, size_t n, double* __restrict__ part_sum,
                             unsigned long long* __restrict__ part_cnt) {
    __shared__ double s_sum[256];
    __shared__ unsigned long long s_cnt[256];
    double acc = 0.0;
    unsigned long long cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (mask[i]) { acc = __dadd_rn(acc, img[i]); cnt++; }
    s_sum[threadIdx.x] = acc; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s_sum[threadIdx.x] = __dadd_rn(s_sum[threadIdx.x], s_sum[threadIdx.x + o]); s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_sum[blockIdx.x] = s_sum[0]; part_cnt[blockIdx.x] = s_cnt[0]; }
}

static DevCam dev_cam(const sva_camera* c) { return DevCam{c->pos[0], c->pos[1], c->pos[2], c->f, c->pixel_size}; }

extern "C" {

int sva_shift_perspective2(sva_ctx* c, const sva_camera* in_cam, const sva_camera* out_cam, const double* depth, int32_t rows, int32_t cols, double* out) {
    if (!c || !in_cam || !out_cam || !depth || !out || rows < 1 || cols < 1) return c ? c->fail(SVA_ERR_BAD_ARG, "shift_perspective2: bad argument") : SVA_ERR_BAD_ARG;
    if ((long long)rows * cols >= 0xFFFFFFFFll) return c->fail(SVA_ERR_BAD_ARG, "shift_perspective2: image too large");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const size_t n = (size_t)rows * cols;
    SVA_TRY(c->reserve(c->scratch, n * 16));        // [depth in][depth out]
    SVA_TRY(c->reserve(c->scratch2, n * 4));        // order keys
    double* d_in = c->scratch.as<double>();
    double* d_out = d_in + n;
    unsigned int* d_ord = c->scratch2.as<unsigned int>();
    SVA_CUDA_OK(c, cudaMemcpyAsync(d_in, depth, n * 8, cudaMemcpyHostToDevice, c->stream));
    SVA_CUDA_OK(c, cudaMemsetAsync(d_out, 0, n * 8, c->stream));
    SVA_CUDA_OK(c, cudaMemsetAsync(d_ord, 0, n * 4, c->stream));
    const double pmx = (in_cam->pos[0] - out_cam->pos[0]) * in_cam->f / in_cam->pixel_size;  // :82 (host f64, -ffp-contract=off)
    const double pmy = (in_cam->pos[1] - out_cam->pos[1]) * in_cam->f / in_cam->pixel_size;  // :83
    const dim3 grid(div_up(cols, 128), rows);
    { LaunchScope ls(c, "k_sp2_order"); k_sp2_order<<<grid, 128, 0, c->stream>>>(d_in, rows, cols, pmx, pmy, d_ord); }
    { LaunchScope ls(c, "k_sp2_write"); k_sp2_write<<<grid, 128, 0, c->stream>>>(d_in, rows, cols, pmx, pmy, d_ord, d_out); }
    SVA_CUDA_OK(c, cudaGetLastError());
    SVA_CUDA_OK(c, cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}

int sva_points3d_to_depth_map(sva_ctx* c, const double* points_xyz, int64_t n_points, const sva_camera* cam, int32_t width, int32_t height, double* out) {
    if (!c || !cam || !out || width < 1 || height < 1 || n_points < 0 || (n_points > 0 && !points_xyz)) return c ? c->fail(SVA_ERR_BAD_ARG, "points3d_to_depth_map: bad argument") : SVA_ERR_BAD_ARG;
    if (n_points >= 0xFFFFFFFFll) return c->fail(SVA_ERR_BAD_ARG, "points3d_to_depth_map: too many points");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const size_t px = (size_t)width * height;
    SVA_TRY(c->reserve(c->scratch, (size_t)n_points * 24 + px * 8 + 64));
    SVA_TRY(c->reserve(c->scratch2, px * 4));
    double* d_pts = c->scratch.as<double>();
    double* d_out = d_pts + 3 * (size_t)n_points;
    unsigned int* d_ord = c->scratch2.as<unsigned int>();
    if (n_points) SVA_CUDA_OK(c, cudaMemcpyAsync(d_pts, points_xyz, (size_t)n_points * 24, cudaMemcpyHostToDevice, c->stream));
    SVA_CUDA_OK(c, cudaMemsetAsync(d_out, 0, px * 8, c->stream));
    SVA_CUDA_OK(c, cudaMemsetAsync(d_ord, 0, px * 4, c->stream));
    if (n_points) {
        const unsigned int blocks = (unsigned int)((n_points + 255) / 256);
        { LaunchScope ls(c, "k_p2d_order"); k_p2d_order<<<blocks, 256, 0, c->stream>>>(d_pts, n_points, dev_cam(cam), width, height, d_ord); }
        { LaunchScope ls(c, "k_p2d_write"); k_p2d_write<<<blocks, 256, 0, c->stream>>>(d_pts, n_points, dev_cam(cam), width, height, d_ord, d_out); }
        SVA_CUDA_OK(c, cudaGetLastError());
    }
    SVA_CUDA_OK(c, cudaMemcpyAsync(out, d_out, px * 8, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}

int sva_depth_map_to_points3d(sva_ctx* c, const double* depth, int32_t rows, int32_t cols, const sva_camera* cam, int32_t width, int32_t height,
                              double* out_xyz, int64_t cap, int64_t* out_count) {
    if (!c || !depth || !cam || !out_count || rows < 1 || cols < 1 || cap < 0 || (cap > 0 && !out_xyz)) return c ? c->fail(SVA_ERR_BAD_ARG, "depth_map_to_points3d: bad argument") : SVA_ERR_BAD_ARG;
    if ((long long)rows * cols >= 0x7FFFFFFFll) return c->fail(SVA_ERR_BAD_ARG, "depth_map_to_points3d: image too large");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const size_t n = (size_t)rows * cols;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)n, c->stream);
    SVA_TRY(c->reserve(c->scratch, n * 8 + n * 24 + 64));                       // [depth][points]
    SVA_TRY(c->reserve(c->scratch2, n * 8 + tmp_bytes + 256));                  // [flags][positions][cub temp]
    double* d_depth = c->scratch.as<double>();
    double* d_pts = d_depth + n;
    unsigned int* d_flags = c->scratch2.as<unsigned int>();
    unsigned int* d_pos = d_flags + n;
    void* d_tmp = (void*)(((uintptr_t)(d_pos + n) + 255) & ~(uintptr_t)255);
    SVA_CUDA_OK(c, cudaMemcpyAsync(d_depth, depth, n * 8, cudaMemcpyHostToDevice, c->stream));
    const dim3 grid(div_up(rows, 128), cols);
    { LaunchScope ls(c, "k_d2p_flags"); k_d2p_flags<<<grid, 128, 0, c->stream>>>(d_depth, rows, cols, d_flags); }
    SVA_CUDA_OK(c, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flags, d_pos, (int)n, c->stream));
    c->launches++;
    unsigned int last_pos = 0, last_flag = 0;
    SVA_CUDA_OK(c, cudaMemcpyAsync(&last_pos, d_pos + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaMemcpyAsync(&last_flag, d_flags + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    { LaunchScope ls(c, "k_d2p_scatter"); k_d2p_scatter<<<grid, 128, 0, c->stream>>>(d_depth, rows, cols, dev_cam(cam), width / 2, height / 2, d_flags, d_pos, cap, d_pts); }
    SVA_CUDA_OK(c, cudaGetLastError());
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    const int64_t count = (int64_t)last_pos + last_flag;
    *out_count = count;
    const int64_t wr = count < cap ? count : cap;
    if (wr > 0) {
        SVA_CUDA_OK(c, cudaMemcpyAsync(out_xyz, d_pts, (size_t)wr * 24, cudaMemcpyDeviceToHost, c->stream));
        SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    }
    return SVA_OK;
}

int sva_masked_mean_f64(sva_ctx* c, const double* image, int32_t rows, int32_t cols, const sva_image_u8* mask, double* out_mean) {
    if (!c || !image || !mask || !mask->data || !out_mean || rows < 1 || cols < 1 || mask->rows != rows || mask->cols != cols || mask->step < (size_t)cols)
        return c ? c->fail(SVA_ERR_BAD_ARG, "masked_mean: bad argument") : SVA_ERR_BAD_ARG;
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const size_t n = (size_t)rows * cols;
    const int blocks = 296;
    SVA_TRY(c->reserve(c->scratch, n * 8 + n + 64));
    SVA_TRY(c->reserve(c->scratch2, (size_t)blocks * 16));
    double* d_img = c->scratch.as<double>();
    uint8_t* d_mask = reinterpret_cast<uint8_t*>(d_img + n);
    double* d_sum = c->scratch2.as<double>();
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_sum + blocks);
    SVA_CUDA_OK(c, cudaMemcpyAsync(d_img, image, n * 8, cudaMemcpyHostToDevice, c->stream));
    SVA_CUDA_OK(c, cudaMemcpy2DAsync(d_mask, cols, mask->data, mask->step, cols, rows, cudaMemcpyHostToDevice, c->stream));
    { LaunchScope ls(c, "k_masked_sum"); k_masked_sum<<<blocks, 256, 0, c->stream>>>(d_img, d_mask, n, d_sum, d_cnt); }
    SVA_CUDA_OK(c, cudaGetLastError());
    std::vector<double> hs(blocks);
    std::vector<unsigned long long> hc(blocks);
    SVA_CUDA_OK(c, cudaMemcpyAsync(hs.data(), d_sum, blocks * 8, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaMemcpyAsync(hc.data(), d_cnt, blocks * 8, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    double sum = 0.0;
    unsigned long long cnt = 0;
    for (int i = 0; i < blocks; i++) { sum += hs[i]; cnt += hc[i]; }
    *out_mean = cnt ? sum / (double)cnt : 0.0;  // cv::mean of an empty mask is 0
    return SVA_OK;
}

}  // extern "C"
