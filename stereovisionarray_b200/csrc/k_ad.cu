// k_ad.cu — K1a, line-image gather form: per-pixel absolute-difference volume summed over camera pairs, for ANY integer pair
// offsets.  The volume pipeline uses the image-space kernel of k_ad2.cu whenever every |gx|,|gy| <= 2 (all BASELINE grids) and
// this one otherwise; the planar output layout (ApGeom), its allocation and the unpack kernel live here.
//
//   A(y,x,d) = sum_k |R(y,x) - I_k(y - gy_k*delta, x - gx_k*delta)|,  delta = min_disp + d      (u16; stored planar, see ApGeom)
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

// One thread = 8 consecutive disparities (one 16-byte chunk of the d axis) of FOUR consecutive pixels of a row.  The lanes of a
// warp are the D/8 chunks of the same two pixel quads, so every 32-bit gather instruction of a warp reads two contiguous
// ~(D+8)-byte runs of a line image (2-4 L1 lines instead of one per lane), and the per-pair set-up is amortised over 32 cells.
// A CTA's work items form 8x8-pixel tiles so the sources of all pair directions stay L1-resident across the tile.
// Output is the planar layout ApGeom (sva_common.cuh): per disparity pair one 16-byte store of 4 consecutive columns.
__global__ void __launch_bounds__(256)
k_ad_planar(const uint8_t* __restrict__ ref, size_t ref_pitch, const uint8_t* __restrict__ lines, AdPairs P, int W, int H, int D, int dmin,
            uint32_t* __restrict__ AP, int wp, int padl, int padt) {
    const int nch = D >> 3, per_group = 2 * nch;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long G = id / per_group;               // group = 8 consecutive pixels of one row
    const int tl = (int)(id - G * per_group);
    const int half = tl / nch, chunk = tl - half * nch;
    const int tiles_x = (W + 7) >> 3;
    const long long tile = G >> 3;
    const int ty = (int)(tile / tiles_x), tx = (int)(tile - (long long)ty * tiles_x);
    const int y = ty * 8 + (int)(G & 7);
    if (y >= H) return;
    const int x0 = tx * 8 + 4 * half;
    int xc[4];
    uint32_t r4[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        xc[i] = min(x0 + i, W - 1);  // columns past the right edge compute on a clamped address and are zeroed below
        r4[i] = (uint32_t)ref[(size_t)y * ref_pitch + xc[i]] * 0x01010101u;
    }
    const int delta0 = dmin + 8 * chunk;
    uint32_t e0[4], o0[4], e1[4], o1[4];  // per pixel: cells (0,2) (1,3) (4,6) (5,7) as u16x2
#pragma unroll
    for (int i = 0; i < 4; i++) { e0[i] = 0; o0[i] = 0; e1[i] = 0; o1[i] = 0; }
    for (int k = 0; k < P.n; k++) {
        const uint8_t* L = lines + P.line_off[k];
        const int g = P.g[k], alpha = P.alpha[k];
        const int offb = P.base[k] + y * P.beta[k] + g * delta0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t b0, b1;
            ad_gather8(L, offb + xc[i] * alpha, g, b0, b1);
            ad_accumulate(__vabsdiffu4(b0, r4[i]), e0[i], o0[i]);
            ad_accumulate(__vabsdiffu4(b1, r4[i]), e1[i], o1[i]);
        }
    }
    uint32_t w[4][4];  // [pixel][d-pair of the chunk]
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool in = x0 + i < W;
        w[i][0] = in ? __byte_perm(e0[i], o0[i], 0x5410) : 0u;
        w[i][1] = in ? __byte_perm(e0[i], o0[i], 0x7632) : 0u;
        w[i][2] = in ? __byte_perm(e1[i], o1[i], 0x5410) : 0u;
        w[i][3] = in ? __byte_perm(e1[i], o1[i], 0x7632) : 0u;
    }
    uint32_t* out = AP + ((size_t)(y + padt) * (D >> 1) + 4 * chunk) * wp + padl + x0;
#pragma unroll
    for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(out + (size_t)j * wp) = make_uint4(w[0][j], w[1][j], w[2][j], w[3][j]);
}

// planar -> [H][W][D] u16 (test / download path and the RAW_U32 recomputation only)
__global__ void k_ap_unpack(const uint32_t* __restrict__ AP, int W, int H, int dph, int wp, int padl, int padt, uint32_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    for (int dp = 0; dp < dph; dp++) out[((size_t)y * W + x) * dph + dp] = AP[((size_t)(y + padt) * dph + dp) * wp + padl + x];
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

// geometry of the planar AD volume for the current parameters; (re)allocates it and establishes its zero borders
int sva_ap_prepare(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp, k = p.win_half;
    ApGeom g;
    const int kk = (k + 3) & ~3;
    g.padl = kk; g.padt = k;
    g.txo = (256 - kk - k + 1) & ~3;
    g.strips = div_up(W, g.txo);
    g.wp = g.txo * (g.strips - 1) + 256;
    const int need = kk + ((W + 7) & ~7);
    if (g.wp < need) g.wp = need;
    g.wp = (g.wp + 7) & ~7;  // 32-byte rows: k_box_planar_shfl reads 32 bytes per lane
    g.hp = H + 2 * k;
    g.row_words = (size_t)(D >> 1) * g.wp;
    g.words = (size_t)g.hp * g.row_words;
    SVA_TRY(ctx->reserve(ctx->AP, g.words * 4));
    ctx->ap = g;
    uint64_t key = (uint64_t)(uintptr_t)ctx->AP.p;
    key = key * 1000003u + (uint64_t)W; key = key * 1000003u + (uint64_t)H; key = key * 1000003u + (uint64_t)D; key = key * 1000003u + (uint64_t)k;
    if (key != ctx->ap_zero_key) {  // k_ad_planar only ever writes the interior, so the borders are zeroed once per geometry
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->AP.p, 0, g.words * 4, ctx->stream));
        ctx->ap_zero_key = key;
    }
    return SVA_OK;
}

// planar AD volume -> ctx->A as [H][W][D] u16
int sva_ap_unpack(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    SVA_TRY(ctx->reserve(ctx->A, (size_t)W * H * D * sizeof(uint16_t)));
    LaunchScope ls(ctx, "k_ap_unpack");
    k_ap_unpack<<<dim3(div_up(W, 128), H), 128, 0, ctx->stream>>>(ctx->AP.as<uint32_t>(), W, H, D >> 1, ctx->ap.wp, ctx->ap.padl, ctx->ap.padt, ctx->A.as<uint32_t>());
    SVA_CUDA_OK(ctx, cudaGetLastError());
    return SVA_OK;
}

int sva_run_ad(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    SVA_TRY(sva_ap_prepare(ctx));
    AdPairs P;
    P.n = ctx->pair_end - ctx->pair_begin;
    for (int i = 0; i < P.n; i++) {
        const PairGeom& G = ctx->geom[ctx->pair_begin + i];
        P.alpha[i] = G.alpha; P.beta[i] = G.beta; P.base[i] = G.base; P.g[i] = G.g; P.line_off[i] = G.offset;
    }
    {
        LaunchScope ls(ctx, "k_ad_planar");
        const long long threads = (long long)div_up(W, 8) * div_up(H, 8) * 8 * (2 * (D >> 3));
        k_ad_planar<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->ref_img.as<uint8_t>(), (size_t)W, ctx->lines.as<uint8_t>(), P, W, H, D, p.min_disp,
                                                                             ctx->AP.as<uint32_t>(), ctx->ap.wp, ctx->ap.padl, ctx->ap.padt);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    ctx->have_ad = true;
    return SVA_OK;
}
