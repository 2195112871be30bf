// k_misc.cu — the small stages either side of the matcher: getAbsDiff, the disparity-driven gather remap (K0),
// improveWithDisparity's +-5 px residual search (K4) and disparity -> depth.
#include "sva_common.cuh"

int sva_launch_box(sva_ctx* ctx, const uint16_t* A, void* out, int W, int H, int D, int k, const sva_params* prm, bool raw, bool apply_validity);

// ---- getAbsDiff — reference src/functions.cpp:215-218 ---------------------------------------------------------------
__global__ void k_abs_diff(const uint8_t* __restrict__ a, size_t apitch, const uint8_t* __restrict__ b, size_t bpitch, int w, int h,
                           unsigned long long* __restrict__ out) {
    unsigned int s = 0;
    for (int y = blockIdx.x; y < h; y += gridDim.x)
        for (int x = threadIdx.x; x < w; x += blockDim.x) s = __sad((int)a[y * apitch + x], (int)b[y * bpitch + x], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, (unsigned long long)s);
}

// ---- K0: shiftPerspectiveWithDisparity — reference src/functions.cpp:55-77 -------------------------------------------
// out(y,x) = img(int(d*uy + y), int(d*ux + x)), skipping d == 0 and out-of-bounds sources (zero there).  A TEXTURE GATHER: the source view
// is bound as a 2-D texture object (point sampling, unnormalised coordinates, border addressing), so the data-dependent reads go through
// the texture path and a source outside the image returns the border value 0 — the reference's bounds test `:66-69` in hardware.  The
// source coordinate itself is the reference's f64 expression, with explicit round-to-nearest intrinsics so that no FMA contraction can
// change the truncation; integer coordinates + 0.5 are exact in f32 for any image the texture unit can hold.
__global__ void k_shift_perspective(const uint8_t* __restrict__ disp, cudaTextureObject_t img, int W, int H, double ux, double uy, uint8_t* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    uint8_t d8 = disp[(size_t)y * W + x];
    uint8_t v = 0;
    if (d8 != 0) {
        double d = (double)d8;
        int sx = (int)__dadd_rn(__dmul_rn(d, ux), (double)x);
        int sy = (int)__dadd_rn(__dmul_rn(d, uy), (double)y);
        v = tex2D<unsigned char>(img, (float)sx + 0.5f, (float)sy + 0.5f);
    }
    out[(size_t)y * W + x] = v;
}

// ---- K4: improveWithDisparity — reference src/functions.cpp:11-52 ----------------------------------------------------
// plane p in [0, 11): |center(y,x) - shifted(y + diry*(p-5), x + dirx*(p-5))|   (u16 [H][W][16], planes 11..15 = zero padding)
#define REFINE_PLANES 16   // 11 used; a multiple of 8 for the box filter's 16-byte copies
__global__ void k_refine_planes(const uint8_t* __restrict__ center, const uint8_t* __restrict__ shifted, int W, int H, int dirx, int diry,
                                uint16_t* __restrict__ A) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int c = center[(size_t)y * W + x];
    uint16_t* o = A + ((size_t)y * W + x) * REFINE_PLANES;
#pragma unroll
    for (int p = 0; p < REFINE_PLANES; p++) {
        int sx = x + dirx * (p - 5), sy = y + diry * (p - 5), v = 0;
        if (p < 11 && sx >= 0 && sx < W && sy >= 0 && sy < H) v = shifted[(size_t)sy * W + sx];
        o[p] = p < 11 ? (uint16_t)abs(c - v) : (uint16_t)0;
    }
}
// first minimum over the 11 window sums, new = disp + (idx-5)*(dirx+diry) narrowed to u8 (:37-38); zero where mask == 0
__global__ void k_refine_select(const uint32_t* __restrict__ Craw, const uint8_t* __restrict__ disp, const uint8_t* __restrict__ mask, int W, int H,
                                int dsum, uint8_t* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    if (mask[i] == 0) return;  // out keeps what earlier cameras wrote (0 initially), like the reference's `continue`
    const uint32_t* c = Craw + i * REFINE_PLANES;
    uint32_t best = c[0];
    int bi = 0;
#pragma unroll
    for (int p = 1; p < 11; p++)
        if (c[p] < best) { best = c[p]; bi = p; }
    out[i] = (uint8_t)((int)disp[i] + (bi - 5) * dsum);
}

// ---- disparity -> depth — reference src/CameraStereoVision.cpp:47,98-100 ---------------------------------------------
__global__ void k_disparity_to_depth(const uint8_t* __restrict__ disp, size_t n, double num, double ps, double* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ddiv_rn(num, __dmul_rn((double)disp[i], ps));  // IEEE: inf where disp == 0
}

static int upload_u8(sva_ctx* c, DevBuf& b, size_t offset, const sva_image_u8* im) {
    SVA_CUDA_OK(c, cudaMemcpy2DAsync((uint8_t*)b.p + offset, im->cols, im->data, im->step, im->cols, im->rows, cudaMemcpyHostToDevice, c->stream));
    return SVA_OK;
}
static bool img_ok(const sva_image_u8* im) { return im && im->data && im->rows > 0 && im->cols > 0 && im->step >= (size_t)im->cols; }

extern "C" {

int sva_abs_diff_u8(sva_ctx* c, const sva_image_u8* a, const sva_image_u8* b, double* out_sum) {
    if (!c || !out_sum) return SVA_ERR_BAD_ARG;
    if (!img_ok(a) || !img_ok(b) || a->rows != b->rows || a->cols != b->cols) return c->fail(SVA_ERR_BAD_ARG, "getAbsDiff: ROIs must be non-empty and of equal size");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    size_t n = (size_t)a->rows * a->cols;
    SVA_TRY(c->reserve(c->scratch, 2 * n + 64));
    SVA_TRY(c->reserve(c->scratch2, 64));
    SVA_TRY(upload_u8(c, c->scratch, 0, a));
    SVA_TRY(upload_u8(c, c->scratch, n, b));
    SVA_CUDA_OK(c, cudaMemsetAsync(c->scratch2.p, 0, 8, c->stream));
    {
        LaunchScope ls(c, "k_abs_diff");
        k_abs_diff<<<a->rows < 1024 ? a->rows : 1024, 128, 0, c->stream>>>(c->scratch.as<uint8_t>(), a->cols, c->scratch.as<uint8_t>() + n, a->cols, a->cols, a->rows,
                                                                      c->scratch2.as<unsigned long long>());
    }
    unsigned long long s = 0;
    SVA_CUDA_OK(c, cudaMemcpyAsync(&s, c->scratch2.p, 8, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    *out_sum = (double)s;
    return SVA_OK;
}

// The view to be warped, as a texture: uploaded into a pitched buffer of the context (the pitch a 2-D texture needs) and bound once per
// buffer / geometry; later uploads of the same geometry reuse the object (stream order keeps kernel and copy apart).
static int texture_upload(sva_ctx* c, const sva_image_u8* im, cudaTextureObject_t* out) {
    const int W = im->cols, H = im->rows;
    const size_t pitch = ((size_t)W + 511) & ~(size_t)511;
    SVA_TRY(c->reserve(c->tex_img, pitch * H));
    const uint64_t key = ((uint64_t)(uintptr_t)c->tex_img.p * 1000003u + (uint64_t)W) * 1000003u + (uint64_t)H;
    if (key != c->tex_key || !c->tex) {
        if (c->tex) { SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream)); cudaDestroyTextureObject(c->tex); c->tex = 0; }
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = c->tex_img.p;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<unsigned char>();
        rd.res.pitch2D.width = W; rd.res.pitch2D.height = H; rd.res.pitch2D.pitchInBytes = pitch;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;  // outside the image: 0
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaTextureObject_t t = 0;
        SVA_CUDA_OK(c, cudaCreateTextureObject(&t, &rd, &td, nullptr));
        c->tex = (unsigned long long)t; c->tex_key = key;
    }
    SVA_CUDA_OK(c, cudaMemcpy2DAsync(c->tex_img.p, pitch, im->data, im->step, W, H, cudaMemcpyHostToDevice, c->stream));
    *out = (cudaTextureObject_t)c->tex;
    return SVA_OK;
}

// device-side core shared by the public call and improveWithDisparity: d_disp, d_out are W*H u8 in HBM, d_img the view as a texture
static int shift_perspective_dev(sva_ctx* c, const sva_camera* in_cam, const sva_camera* out_cam, const uint8_t* d_disp, cudaTextureObject_t d_img, int W, int H,
                                 uint8_t* d_out) {
    double dx = in_cam->pos[0] - out_cam->pos[0], dy = in_cam->pos[1] - out_cam->pos[1], dz = in_cam->pos[2] - out_cam->pos[2];
    double dist = sqrt(dx * dx + dy * dy + dz * dz);  // host TU is built with -ffp-contract=off (:58,61-62)
    double ux = dx / dist, uy = dy / dist;
    LaunchScope ls(c, "k_shift_perspective");
    k_shift_perspective<<<dim3(div_up(W, 256), H), 256, 0, c->stream>>>(d_disp, d_img, W, H, ux, uy, d_out);
    SVA_CUDA_OK(c, cudaGetLastError());
    return SVA_OK;
}

int sva_shift_perspective_with_disparity(sva_ctx* c, const sva_camera* input_cam, const sva_camera* output_cam, const sva_image_u8* disparity,
                                         const sva_image_u8* image, uint8_t* out) {
    if (!c || !input_cam || !output_cam || !out) return SVA_ERR_BAD_ARG;
    if (!img_ok(disparity) || !img_ok(image) || disparity->rows != image->rows || disparity->cols != image->cols)
        return c->fail(SVA_ERR_BAD_ARG, "shiftPerspectiveWithDisparity: disparity and image must have the same size");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const int W = image->cols, H = image->rows;
    size_t n = (size_t)W * H;
    SVA_TRY(c->reserve(c->scratch, 3 * n));
    SVA_TRY(upload_u8(c, c->scratch, 0, disparity));
    cudaTextureObject_t tex = 0;
    SVA_TRY(texture_upload(c, image, &tex));
    uint8_t* base = c->scratch.as<uint8_t>();
    SVA_TRY(shift_perspective_dev(c, input_cam, output_cam, base, tex, W, H, base + 2 * n));
    SVA_CUDA_OK(c, cudaMemcpyAsync(out, base + 2 * n, n, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}

int sva_improve_with_disparity(sva_ctx* c, const sva_image_u8* disparity, const sva_image_u8* center, const sva_image_u8* images, const sva_camera* cams,
                               int32_t n, const sva_image_u8* mask, int32_t window_size, uint8_t* out) {
    if (!c || !out || !cams || n < 0) return SVA_ERR_BAD_ARG;
    if (!img_ok(disparity) || !img_ok(center) || !img_ok(mask)) return c->fail(SVA_ERR_BAD_ARG, "improveWithDisparity: null image");
    const int W = center->cols, H = center->rows, k = (window_size - 1) / 2;  // :17
    if (k < 1 || k > 56) return c->fail(SVA_ERR_BAD_ARG, "improveWithDisparity: windowSize must be in 3..113");
    if (disparity->cols != W || disparity->rows != H || mask->cols != W || mask->rows != H) return c->fail(SVA_ERR_BAD_ARG, "improveWithDisparity: size mismatch");
    for (int i = 0; i < n; i++)
        if (!img_ok(&images[i]) || images[i].cols != W || images[i].rows != H) return c->fail(SVA_ERR_BAD_ARG, "improveWithDisparity: image size mismatch");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    // The reference has no bounds checks: a masked pixel whose windows leave the image throws from Mat::operator()(Rect) (:30,34).
    // Mirror that as SVA_ERR_ROI, decided on the host from the mask's bounding box and the per-camera search directions.
    int x0 = W, x1 = -1, y0 = H, y1 = -1;
    for (int y = 0; y < H; y++) {
        const uint8_t* r = mask->data + (size_t)y * mask->step;
        for (int x = 0; x < W; x++)
            if (r[x]) { x0 = x < x0 ? x : x0; x1 = x > x1 ? x : x1; y0 = y < y0 ? y : y0; y1 = y > y1 ? y : y1; }
    }
    const size_t npx = (size_t)W * H;
    SVA_TRY(c->reserve(c->scratch, 6 * npx));
    uint8_t* base = c->scratch.as<uint8_t>();
    uint8_t *d_disp = base, *d_center = base + npx, *d_mask = base + 2 * npx, *d_shift = base + 4 * npx, *d_out = base + 5 * npx;
    SVA_TRY(upload_u8(c, c->scratch, 0, disparity));
    SVA_TRY(upload_u8(c, c->scratch, npx, center));
    SVA_TRY(upload_u8(c, c->scratch, 2 * npx, mask));
    SVA_CUDA_OK(c, cudaMemsetAsync(d_out, 0, npx, c->stream));
    SVA_TRY(c->reserve(c->A, npx * REFINE_PLANES * sizeof(uint16_t)));
    SVA_TRY(c->reserve(c->Craw, npx * REFINE_PLANES * sizeof(uint32_t)));
    c->have_ad = c->have_cost = false;  // the volume buffers are reused as scratch
    for (int i = 0; i < n; i++) {
        const sva_camera *c0 = &cams[2 * i], *c1 = &cams[2 * i + 1];
        const int dirx = (c0->pos[0] - c1->pos[0]) > 0.001 ? 1 : 0;  // :23-25 (a negative baseline quantises to 0)
        const int diry = (c0->pos[1] - c1->pos[1]) > 0.001 ? 1 : 0;
        if (x1 >= 0) {
            if (x0 - k - 5 * dirx < 0 || y0 - k - 5 * diry < 0 || x1 + k + 5 * dirx > W || y1 + k + 5 * diry > H)
                return c->fail(SVA_ERR_ROI, "improveWithDisparity: a masked pixel's window leaves the image (the reference throws cv::Exception here)");
        }
        cudaTextureObject_t tex = 0;
        SVA_TRY(texture_upload(c, &images[i], &tex));
        SVA_TRY(shift_perspective_dev(c, c0, c1, d_disp, tex, W, H, d_shift));
        {
            LaunchScope ls(c, "k_refine_planes");
            k_refine_planes<<<dim3(div_up(W, 128), H), 128, 0, c->stream>>>(d_center, d_shift, W, H, dirx, diry, c->A.as<uint16_t>());
        }
        SVA_TRY(sva_launch_box(c, c->A.as<uint16_t>(), c->Craw.p, W, H, REFINE_PLANES, k, nullptr, true, false));
        {
            LaunchScope ls(c, "k_refine_select");
            k_refine_select<<<dim3(div_up(W, 128), H), 128, 0, c->stream>>>(c->Craw.as<uint32_t>(), d_disp, d_mask, W, H, dirx + diry, d_out);
        }
        SVA_CUDA_OK(c, cudaGetLastError());
    }
    SVA_CUDA_OK(c, cudaMemcpyAsync(out, d_out, npx, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}

int sva_disparity_to_depth(sva_ctx* c, const sva_image_u8* disparity, double baseline, double f, double pixel_size, double* out_depth) {
    if (!c || !out_depth) return SVA_ERR_BAD_ARG;
    if (!img_ok(disparity)) return c->fail(SVA_ERR_BAD_ARG, "disparity_to_depth: null image");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    size_t n = (size_t)disparity->rows * disparity->cols;
    SVA_TRY(c->reserve(c->scratch, n));
    SVA_TRY(c->reserve(c->scratch2, n * sizeof(double)));
    SVA_TRY(upload_u8(c, c->scratch, 0, disparity));
    {
        LaunchScope ls(c, "k_disparity_to_depth");
        k_disparity_to_depth<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->scratch.as<uint8_t>(), n, baseline * f, pixel_size, c->scratch2.as<double>());
    }
    SVA_CUDA_OK(c, cudaGetLastError());
    SVA_CUDA_OK(c, cudaMemcpyAsync(out_depth, c->scratch2.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}

}  // extern "C"
