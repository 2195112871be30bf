// k_ad.cu — K1a: per-pixel absolute-difference volume summed over camera pairs.
//
//   A(y,x,d) = sum_k |R(y,x) - I_k(y - gy_k*delta, x - gx_k*delta)|,  delta = min_disp + d      (u16, [H][W][D], d fastest)
//
// This is getAbsDiff's |a-b| term (reference src/functions.cpp:215-218) hoisted out of the per-candidate window loop of
// src/CameraStereoVision.cpp:76-83: the 2k x 2k window sum is linear, so sum_pairs(box(|.|)) == box(sum_pairs(|.|)) and the
// box filter runs once (k_box.cu) on this volume instead of once per pair.  For the pair-sharded multi-GPU layout this
// volume (<= 255*n_pairs, carry-free as packed u16) is what gets reduced across GPUs.
//
// Layout trick: each other view is re-laid out ONCE per frame as a "line image" in which the epipolar walk of a pixel
// (source = p - (gx,gy)*delta) is a walk of +g bytes along one row, g = gcd(|gx|,|gy|).  (X,Y) -> (c,t) with
// c = b*X - a*Y, t = u*X + v*Y, (a,b) = -(gx,gy)/g, u*a + v*b = 1 is unimodular, so it is a pure permutation of the
// pixels plus zero padding; out-of-image sources read the padding (= 0, the spec's OOB value).  With that, one thread
// produces 8 consecutive disparities of one pixel from 3 aligned 32-bit loads per pair (g = 1), VABSDIFF4.U8, and
// packed-u16 accumulation — no per-byte loads and no bounds checks in the hot loop.
#include "sva_common.cuh"

struct AdPairs {
    int32_t alpha[SVA_MAX_PAIRS], beta[SVA_MAX_PAIRS], base[SVA_MAX_PAIRS], g[SVA_MAX_PAIRS];
    unsigned long long line_off[SVA_MAX_PAIRS];
    int32_t n;
};

// scatter one other view into its line image (a permutation; the rest of the buffer was zeroed)
__global__ void k_build_line_image(const uint8_t* __restrict__ img, int W, int H, size_t img_pitch, uint8_t* __restrict__ line,
                                   int a, int b, int u, int v, int cmin, int tmin, int pad, int pitch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int c = b * x - a * y - cmin, t = u * x + v * y - tmin + pad;
    line[(size_t)c * pitch + t] = img[(size_t)y * img_pitch + x];
}

__device__ __forceinline__ void ad_accumulate(uint32_t diff, uint32_t& even, uint32_t& odd) {
    even += diff & 0x00FF00FFu;            // cells 0 and 2 as u16x2
    odd += __byte_perm(diff, 0, 0x4341);   // cells 1 and 3 as u16x2
}

// 8 consecutive disparities (bytes) of one pixel's epipolar walk in pair k's line image, starting at byte offset `off`
__device__ __forceinline__ void ad_gather8(const uint8_t* __restrict__ L, int off, int g, uint32_t& b0, uint32_t& b1) {
    const uint32_t* wp = (const uint32_t*)(L + (off & ~3));
    const int sh = (off & 3) * 8;
    if (g == 1) {
        uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
        b0 = __funnelshift_r(w0, w1, sh);
        b1 = __funnelshift_r(w1, w2, sh);
    } else if (g == 2) {
        uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3), w4 = __ldg(wp + 4);
        uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
        uint32_t a2 = __funnelshift_r(w2, w3, sh), a3 = __funnelshift_r(w3, w4, sh);
        b0 = __byte_perm(a0, a1, 0x6420);
        b1 = __byte_perm(a2, a3, 0x6420);
    } else {  // rare general stride: byte gathers
        const uint8_t* q = L + off;
        b0 = q[0] | (q[g] << 8) | (q[2 * g] << 16) | ((uint32_t)q[3 * g] << 24);
        b1 = q[4 * g] | (q[5 * g] << 8) | (q[6 * g] << 16) | ((uint32_t)q[7 * g] << 24);
    }
}

// one thread = CH chunks of 8 disparities of one pixel (CH = 4 when D % 32 == 0: the per-pair setup is amortised over 32 cells).
// The chunks of a thread are interleaved with those of the other threads of the pixel (chunk = gi + groups*c), so every
// 16-byte store instruction of a warp writes whole contiguous runs (full 32-byte sectors)
#define TW 2
#define TH 4
template <int CH>
__global__ void __launch_bounds__(256)
k_ad_volume(const uint8_t* __restrict__ ref, size_t ref_pitch, const uint8_t* __restrict__ lines, AdPairs P, int W, int H, int D,
            int dmin, uint16_t* __restrict__ A) {
    const int groups = D / (8 * CH);
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)W * H * groups;
    if (tid >= total) return;
    int gi = (int)(tid % groups);
    long long pt = tid / groups;
    // pixels are visited in TW x TH micro-tiles (a warp's 8 pixels = 2 x 4): whatever the direction of a pair, the walks of a warp
    // then fall into ~4 line-image rows instead of 8, halving the L1 wavefronts per gather
    int x, y;
    {
        const int tiles_x = W / TW;
        const long long tile = pt / (TW * TH);
        const int in = (int)(pt % (TW * TH));
        const long long full = (long long)tiles_x * (H / TH);
        if (tile < full) {
            x = (int)(tile % tiles_x) * TW + in % TW;
            y = (int)(tile / tiles_x) * TH + in / TW;
        } else {  // ragged right / bottom edges: plain raster over the leftover pixels
            long long r = pt - full * (TW * TH);
            const int wrem = W - tiles_x * TW, hfull = (H / TH) * TH;
            if (r < (long long)wrem * hfull) { x = tiles_x * TW + (int)(r % wrem); y = (int)(r / wrem); }
            else { r -= (long long)wrem * hfull; x = (int)(r % W); y = hfull + (int)(r / W); }
        }
    }
    const long long pix = (long long)y * W + x;
    uint32_t r4 = (uint32_t)ref[(size_t)y * ref_pitch + x] * 0x01010101u;
    int delta0 = dmin + 8 * gi;
    uint32_t e0[CH], o0[CH], e1[CH], o1[CH];  // per chunk: (cells 0,2) (1,3) (4,6) (5,7) as u16x2
#pragma unroll
    for (int c = 0; c < CH; c++) { e0[c] = 0; o0[c] = 0; e1[c] = 0; o1[c] = 0; }
    for (int k = 0; k < P.n; k++) {
        const uint8_t* L = lines + P.line_off[k];
        const int g = P.g[k];
        const int off = P.base[k] + x * P.alpha[k] + y * P.beta[k] + g * delta0;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            uint32_t b0, b1;
            ad_gather8(L, off + 8 * g * groups * c, g, b0, b1);
            ad_accumulate(__vabsdiffu4(b0, r4), e0[c], o0[c]);
            ad_accumulate(__vabsdiffu4(b1, r4), e1[c], o1[c]);
        }
    }
    uint4* out = reinterpret_cast<uint4*>(A + ((size_t)pix * D + 8 * gi));
#pragma unroll
    for (int c = 0; c < CH; c++) {
        uint4 v;
        v.x = __byte_perm(e0[c], o0[c], 0x5410);  // cells 0,1
        v.y = __byte_perm(e0[c], o0[c], 0x7632);  // cells 2,3
        v.z = __byte_perm(e1[c], o1[c], 0x5410);
        v.w = __byte_perm(e1[c], o1[c], 0x7632);
        out[c * groups] = v;
    }
}

static void ext_gcd(int a, int b, int& g, int& u, int& v) {  // u*a + v*b = g >= 0
    if (b == 0) { g = a < 0 ? -a : a; u = a < 0 ? -1 : 1; v = 0; return; }
    int g1, u1, v1;
    ext_gcd(b, a % b, g1, u1, v1);
    g = g1; u = v1; v = u1 - (a / b) * v1;
}

// Build the line images of all pairs for the uploaded frame.  other_imgs: n_pairs images, each H x W, pitch = W.
int sva_build_line_images(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, dmax = p.min_disp + p.num_disp - 1;
    size_t total = 0;
    for (int k = 0; k < p.n_pairs; k++) {
        PairGeom& G = ctx->geom[k];
        int gx = p.pair_gx[k], gy = p.pair_gy[k];
        if (gx == 0 && gy == 0) return ctx->fail(SVA_ERR_BAD_ARG, "pair offset (0,0)");
        int g, u, v, one;
        ext_gcd(-gx, -gy, g, u, v);                 // g = gcd(|gx|,|gy|)
        int a = -gx / g, b = -gy / g;               // source moves by (a,b)*g per unit disparity
        ext_gcd(a, b, one, u, v);                   // (a,b) coprime -> u*a + v*b = 1
        G.g = g; G.a = a; G.b = b; G.u = u; G.v = v;
        int cs[4] = {0, b * (W - 1), -a * (H - 1), b * (W - 1) - a * (H - 1)};
        int ts[4] = {0, u * (W - 1), v * (H - 1), u * (W - 1) + v * (H - 1)};
        int cmin = cs[0], cmax = cs[0], tmin = ts[0], tmax = ts[0];
        for (int i = 1; i < 4; i++) { cmin = cs[i] < cmin ? cs[i] : cmin; cmax = cs[i] > cmax ? cs[i] : cmax; tmin = ts[i] < tmin ? ts[i] : tmin; tmax = ts[i] > tmax ? ts[i] : tmax; }
        G.cmin = cmin; G.tmin = tmin;
        G.pad = 8;                                             // left slack for the aligned-down word load
        G.rows = cmax - cmin + 1;
        G.pitch = ((tmax - tmin + 1) + G.pad + g * (dmax + 8) + 24 + 15) & ~15;  // right slack: the walk + the 5-word over-read
        G.alpha = b * G.pitch + u;
        G.beta = -a * G.pitch + v;
        G.base = -cmin * G.pitch - tmin + G.pad;
        G.offset = total;
        total += (size_t)G.rows * G.pitch;
        total = (total + 255) & ~(size_t)255;
        if ((double)G.rows * G.pitch > 2.0e9) return ctx->fail(SVA_ERR_BAD_ARG, "line image exceeds 2 GB (pair offset too oblique for this image size)");
    }
    SVA_TRY(ctx->reserve(ctx->lines, total + 64));
    SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->lines.p, 0, total + 64, ctx->stream));
    for (int k = 0; k < p.n_pairs; k++) {
        const PairGeom& G = ctx->geom[k];
        dim3 grid(div_up(W, 256), H);
        LaunchScope ls(ctx, "k_build_line_image");
        k_build_line_image<<<grid, 256, 0, ctx->stream>>>(ctx->other_imgs.as<uint8_t>() + (size_t)k * W * H, W, H, (size_t)W,
                                                         ctx->lines.as<uint8_t>() + G.offset, G.a, G.b, G.u, G.v, G.cmin, G.tmin, G.pad, G.pitch);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    return SVA_OK;
}

int sva_run_ad(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    size_t cells = (size_t)W * H * D;
    SVA_TRY(ctx->reserve(ctx->A, cells * sizeof(uint16_t)));
    AdPairs P;
    P.n = ctx->pair_end - ctx->pair_begin;
    for (int i = 0; i < P.n; i++) {
        const PairGeom& G = ctx->geom[ctx->pair_begin + i];
        P.alpha[i] = G.alpha; P.beta[i] = G.beta; P.base[i] = G.base; P.g[i] = G.g; P.line_off[i] = G.offset;
    }
    {
        LaunchScope ls(ctx, "k_ad_volume");
        if (D % 32 == 0) {
            long long threads = (long long)W * H * (D / 32);
            k_ad_volume<4><<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->ref_img.as<uint8_t>(), (size_t)W, ctx->lines.as<uint8_t>(), P, W, H, D,
                                                                                     p.min_disp, ctx->A.as<uint16_t>());
        } else {
            long long threads = (long long)W * H * (D / 8);
            k_ad_volume<1><<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->ref_img.as<uint8_t>(), (size_t)W, ctx->lines.as<uint8_t>(), P, W, H, D,
                                                                                     p.min_disp, ctx->A.as<uint16_t>());
        }
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    ctx->have_ad = true;
    return SVA_OK;
}
