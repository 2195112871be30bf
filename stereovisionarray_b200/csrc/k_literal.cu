// k_literal.cu — "literal mode": the reference driver's loop nest (src/CameraStereoVision.cpp:49-95) as one batched call,
// bit-exact with the reference (pinned through oracle/_ref, tests/golden/main_*.npz).
//
// Per masked pixel and pair the reference (a) builds the epipolar segment from the ray at lengths 0.5 and 1.0 in f64
// (Camera::inv_project / project, src/Camera.cpp:15-33), (b) rasterises it with bresenham (src/functions.cpp:253-321),
// (c) evaluates a 2k x 2k SAD per candidate (getAbsDiff, :215-218), (d) keeps the FIRST minimum and (e) stores
// (uchar)(int)||candidate - pixel||.  Work is Theta(H W D 4k^2) there.
//
// Here the SAD of candidate c for pixel p only depends on the integer offset o = c - p, so the cost of every (pixel, offset) is
// one cell of a box-filtered |R - I(.+o)| plane.  Pass 1 computes the f64 endpoints per pixel (explicit round-to-nearest
// intrinsics, no FMA contraction) and the bounding box of offsets; pass 2 marks which offsets any segment actually visits;
// the used offsets are then processed in chunks of LIT_CHUNK planes: |R - I(.+o)| planes -> K1b box filter (raw u32) -> per-pixel
// walk of its own Bresenham candidates keeping the lexicographic minimum of (cost, candidate index), which IS "first minimum"
// whatever order the chunks are processed in.  A last pass converts the winning candidate into the u8 disparity.
#include <algorithm>
#include <climits>

#include "sva_common.cuh"
#include "sva_cam.cuh"

int sva_launch_box(sva_ctx* ctx, const uint16_t* A, void* out, int W, int H, int D, int k, const sva_params* prm, bool raw, bool apply_validity);

#define LIT_CHUNK 64

// incremental form of plotLineLow / plotLineHigh — src/functions.cpp:253-321 (first argument = the definition's point2)
struct Bres {
    int m, m1, minor, inc, dmaj, dmin, e;
    bool x_major;
    __device__ __forceinline__ void init(int ax, int ay, int bx, int by) {
        x_major = abs(ay - by) < abs(ax - bx);
        bool start_a = x_major ? (bx > ax) : (by > ay);
        int sx = start_a ? ax : bx, sy = start_a ? ay : by, ex = start_a ? bx : ax, ey = start_a ? by : ay;
        m = x_major ? sx : sy; m1 = x_major ? ex : ey; minor = x_major ? sy : sx;
        dmaj = m1 - m; dmin = (x_major ? ey : ex) - minor; inc = 1;
        if (dmin < 0) { inc = -1; dmin = -dmin; }
        e = 2 * dmin - dmaj;
    }
    __device__ __forceinline__ bool done() const { return m > m1; }
    __device__ __forceinline__ int x() const { return x_major ? m : minor; }
    __device__ __forceinline__ int y() const { return x_major ? minor : m; }
    __device__ __forceinline__ void next() {
        if (e > 0) { minor += inc; e -= 2 * dmaj; }
        e += 2 * dmin;
        m++;
    }
};

// pass 1: segment endpoints per pixel (int4; x = INT_MIN when the pixel / pair is skipped) + bounding box of candidate offsets
__global__ void k_lit_endpoints(DevCam cr, DevCam co, const uint8_t* __restrict__ mask, int W, int H, int k, double rnear, double rfar, int4* __restrict__ ends,
                                int* __restrict__ bbox /* oxmin, oxmax, oymin, oymax */) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    int oxmin = INT_MAX, oxmax = INT_MIN, oymin = INT_MAX, oymax = INT_MIN;
    if (x < W) {
        int4 e = make_int4(INT_MIN, 0, 0, 0);
        bool in = x >= k && x < W - k && y >= k && y < H - k && (!mask || mask[(size_t)y * W + x] != 0);  // :49-53
        if (in) {
            const int hx = W / 2, hy = H / 2;                                                           // :28
            double rx, ry, rz;
            dev_inv_project(cr, x - hx, y - hy, rx, ry, rz);                                                // :60
            int ax, ay, bx, by;
            dev_project(co, __dadd_rn(cr.px, __dmul_rn(rx, rnear)), __dadd_rn(cr.py, __dmul_rn(ry, rnear)), __dadd_rn(cr.pz, __dmul_rn(rz, rnear)), ax, ay);  // :61,63
            dev_project(co, __dadd_rn(cr.px, __dmul_rn(rx, rfar)), __dadd_rn(cr.py, __dmul_rn(ry, rfar)), __dadd_rn(cr.pz, __dmul_rn(rz, rfar)), bx, by);    // :62,64
            ax += hx; ay += hy; bx += hx; by += hy;
            bool ok = !(ax < k || ay < k || ax > W - k || ay > H - k) && !(bx < k || by < k || bx > W - k || by > H - k);                                    // :66-71
            if (ok) {
                e = make_int4(ax, ay, bx, by);
                oxmin = min(ax, bx) - x; oxmax = max(ax, bx) - x; oymin = min(ay, by) - y; oymax = max(ay, by) - y;
            }
        }
        ends[(size_t)y * W + x] = e;
    }
    for (int o = 16; o > 0; o >>= 1) {
        oxmin = min(oxmin, __shfl_xor_sync(0xffffffffu, oxmin, o)); oxmax = max(oxmax, __shfl_xor_sync(0xffffffffu, oxmax, o));
        oymin = min(oymin, __shfl_xor_sync(0xffffffffu, oymin, o)); oymax = max(oymax, __shfl_xor_sync(0xffffffffu, oymax, o));
    }
    if ((threadIdx.x & 31) == 0 && oxmin != INT_MAX) {
        atomicMin(&bbox[0], oxmin); atomicMax(&bbox[1], oxmax); atomicMin(&bbox[2], oymin); atomicMax(&bbox[3], oymax);
    }
}

// pass 2: which offsets does any segment visit?  Neighbouring pixels visit nearly the same offsets, so every CTA first collects
// its visits in a shared-memory bitmap with a test before the atomic (after the first few pixels every bit is already set and no atomic
// is issued), then merges the non-empty words into the global bitmap — one global atomic per word per CTA instead of one per
// (pixel, candidate) on a handful of addresses (that form took 51 of the 55 ms of a 1280x960 frame).
__global__ void k_lit_mark(const int4* __restrict__ ends, int W, int H, int oxmin, int oymin, int nox, unsigned int* __restrict__ used, int words,
                           int use_smem) {
    extern __shared__ unsigned int s_used[];
    if (use_smem) {
        for (int i = threadIdx.x; i < words; i += blockDim.x) s_used[i] = 0u;
        __syncthreads();
    }
    volatile unsigned int* bm = use_smem ? s_used : used;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W) {
        int4 e = ends[(size_t)y * W + x];
        if (e.x != INT_MIN) {
            Bres b; b.init(e.x, e.y, e.z, e.w);
            for (; !b.done(); b.next()) {
                const int pi = (b.y() - y - oymin) * nox + (b.x() - x - oxmin);
                const unsigned int bit = 1u << (pi & 31);
                if (!(bm[pi >> 5] & bit)) atomicOr(const_cast<unsigned int*>(&bm[pi >> 5]), bit);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            const unsigned int v = s_used[i];
            if (v && (__ldcg(used + i) & v) != v) atomicOr(used + i, v);
        }
    }
}

// |R(y,x) - I(y+oy, x+ox)| for a chunk of offsets -> u16 [H][W][LIT_CHUNK]
__global__ void k_lit_planes(const uint8_t* __restrict__ R, const uint8_t* __restrict__ I, int W, int H, const int2* __restrict__ offs, int n_offs,
                             uint16_t* __restrict__ A) {
    int p = threadIdx.x;  // LIT_CHUNK threads per pixel-row slice
    int x = blockIdx.x * blockDim.y + threadIdx.y, y = blockIdx.y;
    if (x >= W) return;
    int v = 0;
    if (p < n_offs) {
        int2 o = offs[p];
        int sx = x + o.x, sy = y + o.y, s = 0;
        if (sx >= 0 && sx < W && sy >= 0 && sy < H) s = __ldg(I + (size_t)sy * W + sx);
        v = abs((int)R[(size_t)y * W + x] - s);
    }
    A[((size_t)y * W + x) * LIT_CHUNK + p] = (uint16_t)v;
}

// each pixel walks its own candidates; lexicographic min of (cost, candidate index) == first minimum (:85)
__global__ void k_lit_select(const int4* __restrict__ ends, const uint32_t* __restrict__ Craw, const int* __restrict__ slot_of, int W, int H, int oxmin, int oymin,
                             int nox, int slot_begin, int slot_end, unsigned long long* __restrict__ best) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    int4 e = ends[i];
    if (e.x == INT_MIN) return;
    unsigned long long bk = best[i];
    Bres b; b.init(e.x, e.y, e.z, e.w);
    const uint32_t* c = Craw + i * LIT_CHUNK;
    for (unsigned int idx = 0; !b.done(); b.next(), idx++) {
        int slot = slot_of[(b.y() - y - oymin) * nox + (b.x() - x - oxmin)];
        if (slot < slot_begin || slot >= slot_end) continue;
        unsigned long long key = ((unsigned long long)c[slot - slot_begin] << 32) | idx;
        bk = key < bk ? key : bk;
    }
    best[i] = bk;
}

// disparity = (uchar)(int)||candidate - pixel||  (:87-89; plain narrowing, f64 sqrt)
__global__ void k_lit_finalize(const int4* __restrict__ ends, const unsigned long long* __restrict__ best, int W, int H, uint8_t* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    int4 e = ends[i];
    if (e.x == INT_MIN) return;
    unsigned int idx = (unsigned int)(best[i] & 0xFFFFFFFFull);
    Bres b; b.init(e.x, e.y, e.z, e.w);
    for (unsigned int j = 0; j < idx; j++) b.next();
    int dx = b.x() - x, dy = b.y() - y;
    double n = __dsqrt_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
    out[i] = (uint8_t)(int)n;
}

extern "C" int sva_match_literal(sva_ctx* c, const sva_image_u8* images, const sva_camera* cams, int32_t n_images, const int32_t* pairs, int32_t n_pairs,
                                 const sva_image_u8* mask, int32_t k, double ray_near, double ray_far, uint8_t* out_disp) {
    if (!c || !out_disp) return SVA_ERR_BAD_ARG;
    if (!images || !cams || !pairs || n_images < 1 || n_pairs < 1) return c->fail(SVA_ERR_BAD_ARG, "match_literal: null / empty argument");
    if (k < 1 || k > 56) return c->fail(SVA_ERR_BAD_ARG, "match_literal: kernelSize must be in 1..56");
    const int W = images[0].cols, H = images[0].rows;
    for (int i = 0; i < n_images; i++)
        if (!images[i].data || images[i].cols != W || images[i].rows != H || images[i].step < (size_t)W) return c->fail(SVA_ERR_BAD_ARG, "match_literal: image size mismatch");
    if (mask && (!mask->data || mask->cols != W || mask->rows != H)) return c->fail(SVA_ERR_BAD_ARG, "match_literal: mask size mismatch");
    for (int i = 0; i < 2 * n_pairs; i++)
        if (pairs[i] < 0 || pairs[i] >= n_images) return c->fail(SVA_ERR_BAD_ARG, "match_literal: pair index out of range");
    SVA_CUDA_OK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)W * H;
    c->have_ad = c->have_cost = false;  // volume buffers double as scratch
    // scratch: [ref img][other img][mask][out]
    SVA_TRY(c->reserve(c->scratch, 4 * npx));
    uint8_t* base = c->scratch.as<uint8_t>();
    uint8_t *d_ref = base, *d_oth = base + npx, *d_mask = base + 2 * npx, *d_out = base + 3 * npx;
    if (mask) SVA_CUDA_OK(c, cudaMemcpy2DAsync(d_mask, W, mask->data, mask->step, W, H, cudaMemcpyHostToDevice, c->stream));
    SVA_CUDA_OK(c, cudaMemsetAsync(d_out, 0, npx, c->stream));
    // scratch2: [ends int4][best u64][bbox 4 ints]
    SVA_TRY(c->reserve(c->scratch2, npx * 16 + npx * 8 + 64));
    int4* d_ends = c->scratch2.as<int4>();
    unsigned long long* d_best = reinterpret_cast<unsigned long long*>(c->scratch2.as<uint8_t>() + npx * 16);
    int* d_bbox = reinterpret_cast<int*>(c->scratch2.as<uint8_t>() + npx * 24);
    SVA_TRY(c->reserve(c->A, npx * LIT_CHUNK * sizeof(uint16_t)));
    SVA_TRY(c->reserve(c->Craw, npx * LIT_CHUNK * sizeof(uint32_t)));
    const dim3 g2(div_up(W, 128), H);

    for (int pi = 0; pi < n_pairs; pi++) {
        const int r = pairs[2 * pi], o = pairs[2 * pi + 1];
        SVA_CUDA_OK(c, cudaMemcpy2DAsync(d_ref, W, images[r].data, images[r].step, W, H, cudaMemcpyHostToDevice, c->stream));
        SVA_CUDA_OK(c, cudaMemcpy2DAsync(d_oth, W, images[o].data, images[o].step, W, H, cudaMemcpyHostToDevice, c->stream));
        const int init_bbox[4] = {INT_MAX, INT_MIN, INT_MAX, INT_MIN};
        SVA_CUDA_OK(c, cudaMemcpyAsync(d_bbox, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice, c->stream));
        DevCam cr{cams[r].pos[0], cams[r].pos[1], cams[r].pos[2], cams[r].f, cams[r].pixel_size};
        DevCam co{cams[o].pos[0], cams[o].pos[1], cams[o].pos[2], cams[o].f, cams[o].pixel_size};
        {
            LaunchScope ls(c, "k_lit_endpoints");
            k_lit_endpoints<<<g2, 128, 0, c->stream>>>(cr, co, mask ? d_mask : nullptr, W, H, k, ray_near, ray_far, d_ends, d_bbox);
        }
        int bbox[4];
        SVA_CUDA_OK(c, cudaMemcpyAsync(bbox, d_bbox, sizeof(bbox), cudaMemcpyDeviceToHost, c->stream));
        SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
        if (bbox[0] == INT_MAX) continue;  // no pixel survives for this pair: it writes nothing (:66-71)
        const int oxmin = bbox[0], oymin = bbox[2], nox = bbox[1] - bbox[0] + 1, noy = bbox[3] - bbox[2] + 1;
        const long long nplanes = (long long)nox * noy;
        if (nplanes > (1 << 26)) return c->fail(SVA_ERR_BAD_ARG, "match_literal: candidate offset range too large");
        const size_t words = (size_t)((nplanes + 31) / 32);
        DevBuf& tab = c->other_d;  // [used bitmap][slot_of table][offset list]
        SVA_TRY(c->reserve(tab, words * 4 + (size_t)nplanes * 4 + (size_t)nplanes * 8 + 64));
        unsigned int* d_used = tab.as<unsigned int>();
        int* d_slot = reinterpret_cast<int*>(tab.as<uint8_t>() + words * 4);
        int2* d_offs = reinterpret_cast<int2*>(tab.as<uint8_t>() + ((words * 4 + (size_t)nplanes * 4 + 15) & ~(size_t)15));
        SVA_CUDA_OK(c, cudaMemsetAsync(d_used, 0, words * 4, c->stream));
        {
            LaunchScope ls(c, "k_lit_mark");
            const int use_smem = words <= 10240 ? 1 : 0;  // 40 KB of shared bitmap at most; larger offset ranges test-and-set the global one
            k_lit_mark<<<g2, 128, use_smem ? words * 4 : 0, c->stream>>>(d_ends, W, H, oxmin, oymin, nox, d_used, (int)words, use_smem);
        }
        std::vector<unsigned int> used(words);
        SVA_CUDA_OK(c, cudaMemcpyAsync(used.data(), d_used, words * 4, cudaMemcpyDeviceToHost, c->stream));
        SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
        std::vector<int> slot_of((size_t)nplanes, -1);
        std::vector<int2> offs;
        for (long long p = 0; p < nplanes; p++)
            if (used[p >> 5] & (1u << (p & 31))) {
                slot_of[p] = (int)offs.size();
                offs.push_back(make_int2((int)(p % nox) + oxmin, (int)(p / nox) + oymin));
            }
        SVA_CUDA_OK(c, cudaMemcpyAsync(d_slot, slot_of.data(), (size_t)nplanes * 4, cudaMemcpyHostToDevice, c->stream));
        SVA_CUDA_OK(c, cudaMemcpyAsync(d_offs, offs.data(), offs.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
        SVA_CUDA_OK(c, cudaMemsetAsync(d_best, 0xFF, npx * 8, c->stream));
        const int nslots = (int)offs.size();
        for (int s0 = 0; s0 < nslots; s0 += LIT_CHUNK) {
            const int ns = std::min(LIT_CHUNK, nslots - s0);
            {
                LaunchScope ls(c, "k_lit_planes");
                k_lit_planes<<<dim3(div_up(W, 4), H), dim3(LIT_CHUNK, 4), 0, c->stream>>>(d_ref, d_oth, W, H, d_offs + s0, ns, c->A.as<uint16_t>());
            }
            SVA_TRY(sva_launch_box(c, c->A.as<uint16_t>(), c->Craw.p, W, H, LIT_CHUNK, k, nullptr, true, false));
            {
                LaunchScope ls(c, "k_lit_select");
                k_lit_select<<<g2, 128, 0, c->stream>>>(d_ends, c->Craw.as<uint32_t>(), d_slot, W, H, oxmin, oymin, nox, s0, s0 + ns, d_best);
            }
        }
        {
            LaunchScope ls(c, "k_lit_finalize");
            k_lit_finalize<<<g2, 128, 0, c->stream>>>(d_ends, d_best, W, H, d_out);
        }
        SVA_CUDA_OK(c, cudaGetLastError());
        // host vectors (slot_of / offs) must outlive the async copies
        SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    }
    SVA_CUDA_OK(c, cudaMemcpyAsync(out_disp, d_out, npx, cudaMemcpyDeviceToHost, c->stream));
    SVA_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SVA_OK;
}
