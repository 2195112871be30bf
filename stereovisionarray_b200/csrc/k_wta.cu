// k_wta.cu — K3 as a tile-parallel kernel: winner-take-all (first minimum), parabolic sub-pixel fit, pixel validity, and the
// other view's WTA for the left-right check, on the fully aggregated volume S (or on C when n_paths == 0).
//
// Frozen spec: DESIGN.md §3.4, bit-exact with orc_wta() in oracle/sva_oracle.c.  No reference counterpart.
//
// A CTA stages one row segment of WTA_TX pixels x D disparities in shared memory with 16-byte cp.async copies (pixel stride
// padded to == 4 words mod 32 so both access patterns below are bank-conflict free):
//   * WTA: 4 threads per pixel scan interleaved word columns, keys (S << 16 | d) so that min == first minimum, 2 shuffles
//     combine them; the winner's neighbours S(d*-1), S(d*+1) come straight from the tile for the parabola.
//   * left-right: D_o(x') = argmin_d S(y, x' + lr_gx*delta, d) is a minimum along a diagonal of the (x, d) slice.  One thread
//     per x' walks the part of its diagonal that lies inside this tile (consecutive x' -> consecutive halfwords) and merges
//     the partial minimum into a global u32 key map with one atomicMin per (tile, x') — ~3 atomics per pixel in total.
// k_lr_check then compares d* with D_o(x - lr_gx*delta*) per pixel.
#include "sva_common.cuh"
#include "sva_vec.cuh"

#define WTA_TX 64
#define WTA_THREADS 256

struct WtaParams {
    const uint16_t* S;
    int W, H, D, dmin, k;
    int gxp, gxn, gyp, gyn;
    int lr_gx, lr_max_diff, subpixel;
    int stride_w;  // padded pixel stride in 32-bit words
    int y0, Hfull; // the volume holds image rows [y0, y0 + H) of an image Hfull rows high (row-sharded WTA); whole image: 0, H
    const uint8_t* mask;
    uint16_t* disp;
    float* sub;
    uint32_t* other_key;  // [H][W], pre-set to 0xFFFFFFFF
};

__global__ void __launch_bounds__(WTA_THREADS)
k_wta_tile(WtaParams q) {
    extern __shared__ __align__(16) uint32_t tile[];  // [WTA_TX][stride_w]
    __shared__ uint16_t s_d[WTA_TX];
    __shared__ float s_sub[WTA_TX];
    const int t = threadIdx.x, W = q.W, H = q.H, D = q.D, y = blockIdx.y, x0 = blockIdx.x * WTA_TX;
    const int npx = min(WTA_TX, W - x0);
    const int chunks_per_px = D >> 3;
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    const uint16_t* src = q.S + ((size_t)y * W + x0) * D;
    for (int c = t; c < npx * chunks_per_px; c += WTA_THREADS) {
        int px = c / chunks_per_px, part = c - px * chunks_per_px;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_s + (px * q.stride_w + part * 4) * 4), "l"(src + (size_t)px * D + part * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- WTA + sub-pixel: 4 threads per pixel ----
    {
        const int px = t >> 2, qd = t & 3;
        uint32_t best = 0xFFFFFFFFu;
        if (px < npx) {
            const uint32_t* row = tile + px * q.stride_w;
            for (int w = qd; w < (D >> 1); w += 4) {
                uint32_t v = row[w];
                uint32_t k0 = (v << 16) | (uint32_t)(2 * w), k1 = (v & 0xFFFF0000u) | (uint32_t)(2 * w + 1);
                best = min(best, min(k0, k1));
            }
        }
        best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
        best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
        if (qd == 0 && px < npx) {
            const int d = (int)(best & 0xFFFFu), x = x0 + px, k = q.k;
            const int delta = q.dmin + d;
            float f = (float)d;
            if (q.subpixel && d > 0 && d < D - 1) {
                const uint16_t* r16 = reinterpret_cast<const uint16_t*>(tile + px * q.stride_w);
                const int sl = r16[d - 1], s0 = (int)(best >> 16), sr = r16[d + 1];
                const int den = sl - 2 * s0 + sr;
                if (den > 0) f = (float)d + (float)(sl - sr) / (float)(2 * den);
            }
            const int yy = y + q.y0, HH = q.Hfull;  // image row (the volume may be a row block)
            bool ok = x >= k && x < W - k && yy >= k && yy < HH - k;
            if (ok && q.mask) ok = q.mask[(size_t)yy * W + x] != 0;
            if (ok) {
                int lim = 0x7FFFFFFF;
                if (q.gxp > 0) lim = min(lim, (x - k) / q.gxp);
                if (q.gxn > 0) lim = min(lim, (W - k - x) / q.gxn);
                if (q.gyp > 0) lim = min(lim, (yy - k) / q.gyp);
                if (q.gyn > 0) lim = min(lim, (HH - k - yy) / q.gyn);
                ok = delta <= lim;
            }
            s_d[px] = ok ? (uint16_t)delta : (uint16_t)SVA_DISP_INVALID;
            s_sub[px] = ok ? (float)q.dmin + f : SVA_SUBPIX_INVALID;
        }
    }
    // ---- other view's WTA: partial minima of the diagonals crossing this tile ----
    if (q.lr_gx != 0) {
        // x' = x - lr_gx*(dmin + d)  <=>  d = lr_gx*(x - x') - dmin
        const int g = q.lr_gx;
        const int xo_lo = g < 0 ? x0 + q.dmin : x0 - q.dmin - (D - 1);
        const int n_xo = npx + D - 1;
        const uint16_t* t16 = reinterpret_cast<const uint16_t*>(tile);
        const int stride_h = q.stride_w * 2;
        for (int i = t; i < n_xo; i += WTA_THREADS) {
            const int xo = xo_lo + i;
            if (xo < 0 || xo >= W) continue;
            uint32_t best = 0xFFFFFFFFu;
            for (int px = 0; px < npx; px++) {
                const int d = g * (x0 + px - xo) - q.dmin;
                if (d < 0 || d >= D) continue;
                best = min(best, ((uint32_t)t16[px * stride_h + d] << 16) | (uint32_t)d);
            }
            if (best != 0xFFFFFFFFu) atomicMin(&q.other_key[(size_t)y * W + xo], best);
        }
    }
    __syncthreads();
    if (t < npx) {
        q.disp[(size_t)y * W + x0 + t] = s_d[t];
        if (q.sub) q.sub[(size_t)y * W + x0 + t] = s_sub[t];
    }
}

// reject when the matching other-view pixel is outside the row, has no candidate, or disagrees by more than lr_max_diff
__global__ void k_lr_check(const uint32_t* __restrict__ other_key, int W, int H, int dmin, int lr_gx, int lr_max_diff, uint16_t* __restrict__ disp,
                           float* __restrict__ sub) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    int delta = disp[i];
    if (delta == SVA_DISP_INVALID) return;
    int xo = x - lr_gx * delta;
    bool ok = xo >= 0 && xo < W;
    if (ok) {
        uint32_t key = other_key[(size_t)y * W + xo];
        ok = key != 0xFFFFFFFFu && abs((delta - dmin) - (int)(key & 0xFFFFu)) <= lr_max_diff;
    }
    if (!ok) {
        disp[i] = (uint16_t)SVA_DISP_INVALID;
        if (sub) sub[i] = SVA_SUBPIX_INVALID;
    }
}

// ---- K3 as a register march (D = 64, 128, 192 or 256) ------------------------------------------------------------------------
// Half a warp (16 lanes x NV = D/16 consecutive disparities) marches along a row segment; the other half marches along another
// one.  Per pixel and lane: one PRMT per cell builds the (S << 16 | d) keys, VIMNMX3 folds them into the lane's best key and four
// shuffles into the pixel's winner.  The other view's WTA (left-right check) is a systolic diagonal minimum kept in the SAME key
// registers' shadow: entry x' = x - lr_gx*(dmin + d) sits in slot d while the march is at x and moves one slot per pixel, which is
// a register rotation (the step loop is unrolled NV times) plus one shuffle for the lane edge; entries leaving the volume, and
// everything still inside at the segment end, are merged into the global key map with atomicMin (segments are independent).
// The integer scan is all this kernel does: validity, LR comparison and the parabola run once per pixel in k_wta_finish.
#define WSEG_PF 7  // ring of 8 stages: the slot index is s & 7

template <int NV, int TDIR>
__global__ void __launch_bounds__(256)
k_wta_seg(const uint16_t* __restrict__ S, int W, int H, int D, int dmin, int lr_gx, int seg_len, int segs_per_row, uint16_t* __restrict__ dwin,
          uint32_t* __restrict__ other_key) {
    constexpr int NR = NV / 2, NS = WSEG_PF + 1, STAGE = 32 * NV * 2;
    extern __shared__ __align__(16) unsigned char wseg_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lin = lane & 15;
    const int item = (blockIdx.x * (blockDim.x >> 5) + warp) * 2 + (lane >> 4);  // (row, segment)
    const int n_items = H * segs_per_row;
    const bool live = item < n_items;
    const int it = live ? item : n_items - 1;  // a ragged last half-warp redoes the last item and drops its results
    const int y = it / segs_per_row, sg = it - y * segs_per_row;
    const int xa = sg * seg_len, xb = min(W, xa + seg_len), len = xb - xa;
    const int steps = seg_len;  // uniform trip count for both halves of the warp; steps beyond len are masked
    const uint32_t ring = smem_u32(wseg_smem) + warp * (NS * STAGE) + lane * (NV * 2);
    const uint16_t* src = S + ((size_t)y * W + xa) * D + lin * NV;
    uint32_t dpair[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) dpair[j] = (uint32_t)(lin * NV + 2 * j) | ((uint32_t)(lin * NV + 2 * j + 1) << 16);
    uint32_t acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0xFFFFFFFFu;
#pragma unroll
    for (int u = 0; u < WSEG_PF; u++) {
        if (u < len) Vec<NR>::cp_async(ring + u * STAGE, src + (size_t)u * D);
        cp_async_commit();
    }
    uint32_t* okrow = other_key + (size_t)y * W;
    // entry leaving the volume at pixel x: x' = x - lr_gx*dmin (TDIR < 0, leaves below d = 0) or x - lr_gx*(dmin + D - 1) (TDIR > 0)
    const int xo_off = TDIR < 0 ? -lr_gx * dmin : -lr_gx * (dmin + D - 1);

    for (int s0 = 0; s0 < steps; s0 += NV) {
#pragma unroll
        for (int u = 0; u < NV; u++) {
            const int s = s0 + u;
            if (s >= steps) break;
            cp_async_wait<WSEG_PF - 1>();
            uint32_t r[NR];
            Vec<NR>::lds(ring + (s & (NS - 1)) * STAGE, r);
            if (s + WSEG_PF < len) Vec<NR>::cp_async(ring + ((s + WSEG_PF) & (NS - 1)) * STAGE, src + (size_t)(s + WSEG_PF) * D);
            cp_async_commit();
            const bool in = s < len;
            uint32_t key[NV];  // pixels past the segment end read as S = 0xFFFF (> any real S <= 65520): keys >= 0xFFFF0000 mean "no candidate"
#pragma unroll
            for (int j = 0; j < NR; j++) {
                const uint32_t v = in ? r[j] : 0xFFFFFFFFu;
                key[2 * j] = __byte_perm(v, dpair[j], 0x1054);
                key[2 * j + 1] = __byte_perm(v, dpair[j], 0x3276);
            }
            uint32_t kb = key[0];
#pragma unroll
            for (int j = 1; j + 1 < NV; j += 2) kb = min(kb, min(key[j], key[j + 1]));
            kb = min(kb, key[NV - 1]);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) kb = min(kb, __shfl_xor_sync(0xffffffffu, kb, o, 16));
            if (lin == 0 && in && live) dwin[(size_t)y * W + xa + s] = (uint16_t)(kb & 0xFFFFu);
            if (TDIR != 0) {
                // logical slot j lives in physical register (j + u) % NV (TDIR < 0) or (j - u) mod NV (TDIR > 0): the per-pixel shift is free
                if (TDIR < 0) {
                    const uint32_t leaving = acc[u % NV];                       // logical slot 0: has seen d = 0 .. its whole diagonal
                    uint32_t carry = __shfl_down_sync(0xffffffffu, leaving, 1, 16);
                    if (lin == 15) carry = 0xFFFFFFFFu;
                    if (lin == 0 && leaving < 0xFFFF0000u && live) {
                        const int xo = xa + s - 1 + xo_off;                      // it left after the previous pixel
                        if (xo >= 0 && xo < W) atomicMin(okrow + xo, leaving);
                    }
                    acc[u % NV] = carry;                                        // becomes logical slot NV-1
#pragma unroll
                    for (int j = 0; j < NV; j++) acc[(j + u + 1) % NV] = min(acc[(j + u + 1) % NV], key[j]);
                } else {
                    const uint32_t leaving = acc[(NV - 1 - u % NV + NV) % NV];  // logical slot NV-1
                    uint32_t carry = __shfl_up_sync(0xffffffffu, leaving, 1, 16);
                    if (lin == 0) carry = 0xFFFFFFFFu;
                    if (lin == 15 && leaving < 0xFFFF0000u && live) {
                        const int xo = xa + s - 1 + xo_off;
                        if (xo >= 0 && xo < W) atomicMin(okrow + xo, leaving);
                    }
                    acc[(NV - 1 - u % NV + NV) % NV] = carry;                   // becomes logical slot 0
#pragma unroll
                    for (int j = 0; j < NV; j++) acc[(j - u - 1 + 2 * NV) % NV] = min(acc[(j - u - 1 + 2 * NV) % NV], key[j]);
                }
            }
        }
    }
    if (TDIR != 0 && live) {
        // flush: after `steps` pixels (a multiple of NV when the loop ran to completion; `steps % NV` otherwise) logical slot j is at
        // physical (j + steps) % NV resp. (j - steps) mod NV, and holds the entry x' = x_last - lr_gx*(dmin + d_j) ... relative to the last pixel
        const int rot = steps % NV;
        const int x_last = xa + steps - 1;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int pj = TDIR < 0 ? (j + rot) % NV : (j - rot + NV) % NV;
            uint32_t v = acc[0];
#pragma unroll
            for (int q = 1; q < NV; q++) if (q == pj) v = acc[q];
            const int xo = x_last - lr_gx * (dmin + lin * NV + j);
            if (v < 0xFFFF0000u && xo >= 0 && xo < W) atomicMin(okrow + xo, v);
        }
    }
}

// per pixel: border / mask / cell validity, left-right check against the other view's key map, parabolic sub-pixel fit
__global__ void k_wta_finish(const uint16_t* __restrict__ S, const uint16_t* __restrict__ dwin, const uint32_t* __restrict__ other_key, WtaParams q) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int W = q.W, H = q.H, D = q.D, k = q.k;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    const int d = dwin[i], delta = q.dmin + d;
    const int yy = y + q.y0, HH = q.Hfull;  // image row (the volume may be a row block)
    bool ok = x >= k && x < W - k && yy >= k && yy < HH - k;
    if (ok && q.mask) ok = q.mask[(size_t)yy * W + x] != 0;
    if (ok) {
        int lim = 0x7FFFFFFF;
        if (q.gxp > 0) lim = min(lim, (x - k) / q.gxp);
        if (q.gxn > 0) lim = min(lim, (W - k - x) / q.gxn);
        if (q.gyp > 0) lim = min(lim, (yy - k) / q.gyp);
        if (q.gyn > 0) lim = min(lim, (HH - k - yy) / q.gyn);
        ok = delta <= lim;
    }
    if (ok && q.lr_gx != 0) {
        const int xo = x - q.lr_gx * delta;
        ok = xo >= 0 && xo < W;
        if (ok) {
            const uint32_t key = other_key[(size_t)y * W + xo];
            ok = key != 0xFFFFFFFFu && abs(d - (int)(key & 0xFFFFu)) <= q.lr_max_diff;
        }
    }
    float f = (float)d;
    if (ok && q.sub && q.subpixel && d > 0 && d < D - 1) {
        const uint16_t* s = S + i * D + d;
        const int sl = s[-1], s0 = s[0], sr = s[1];
        const int den = sl - 2 * s0 + sr;
        if (den > 0) f = (float)d + (float)(sl - sr) / (float)(2 * den);
    }
    q.disp[i] = ok ? (uint16_t)delta : (uint16_t)SVA_DISP_INVALID;
    if (q.sub) q.sub[i] = ok ? (float)q.dmin + f : SVA_SUBPIX_INVALID;
}

template <int NV>
static cudaError_t wta_seg_launch(sva_ctx* ctx, const uint16_t* vol, const WtaParams& q, int seg_len, int segs_per_row, uint16_t* dwin) {
    const int warps = 8, items = q.H * segs_per_row;
    const int grid = div_up(div_up(items, 2), warps);
    const size_t smem = (size_t)warps * (WSEG_PF + 1) * 32 * NV * 2;
    cudaError_t e;
#define WSEG_GO(TD)                                                                                                              \
    e = cudaFuncSetAttribute(k_wta_seg<NV, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
    if (e == cudaSuccess) k_wta_seg<NV, TD><<<grid, warps * 32, smem, ctx->stream>>>(vol, q.W, q.H, q.D, q.dmin, q.lr_gx, seg_len, segs_per_row, dwin, q.other_key);
    if (q.lr_gx < 0) { WSEG_GO(-1) } else if (q.lr_gx > 0) { WSEG_GO(1) } else { WSEG_GO(0) }
#undef WSEG_GO
    return e != cudaSuccess ? e : cudaGetLastError();
}

// vol = S_total (or C when there is no aggregation); fills ctx->disp / ctx->subpix
int sva_run_wta_rows(sva_ctx* ctx, const uint16_t* vol, int y0, int rows);
int sva_run_wta(sva_ctx* ctx, const uint16_t* vol) { return sva_run_wta_rows(ctx, vol, 0, ctx->prm.height); }

// K3 on a block of `rows` image rows starting at y0 (vol = those rows of the aggregated volume); results go to the first `rows` rows of
// ctx->disp / ctx->subpix.  The left-right check is row-local, so row blocks are independent (row-sharded multi-GPU WTA).
int sva_run_wta_rows(sva_ctx* ctx, const uint16_t* vol, int y0, int rows) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = rows, D = p.num_disp;
    if (y0 < 0 || rows < 1 || y0 + rows > p.height) return ctx->fail(SVA_ERR_BAD_ARG, "row block outside the image");
    SVA_TRY(ctx->reserve(ctx->disp, (size_t)W * p.height * sizeof(uint16_t)));
    SVA_TRY(ctx->reserve(ctx->subpix, (size_t)W * p.height * sizeof(float)));
    WtaParams q{};
    q.S = vol; q.W = W; q.H = H; q.D = D; q.dmin = p.min_disp; q.k = p.win_half; q.y0 = y0; q.Hfull = p.height;
    q.lr_gx = p.lr_gx; q.lr_max_diff = p.lr_max_diff; q.subpixel = p.subpixel;
    for (int i = 0; i < p.n_pairs; i++) {
        int gx = p.pair_gx[i], gy = p.pair_gy[i];
        if (gx > 0) q.gxp = gx > q.gxp ? gx : q.gxp;
        if (gx < 0) q.gxn = -gx > q.gxn ? -gx : q.gxn;
        if (gy > 0) q.gyp = gy > q.gyp ? gy : q.gyp;
        if (gy < 0) q.gyn = -gy > q.gyn ? -gy : q.gyn;
    }
    q.stride_w = D / 2 + (((4 - D / 2) % 32) + 32) % 32;  // == 4 (mod 32): conflict-free for both access patterns, 16-byte aligned
    q.mask = ctx->has_mask ? ctx->mask.as<uint8_t>() : nullptr;
    q.disp = ctx->disp.as<uint16_t>(); q.sub = ctx->subpix.as<float>();
    if (p.lr_gx != 0) {
        SVA_TRY(ctx->reserve(ctx->other_d, (size_t)W * H * sizeof(uint32_t)));
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->other_d.p, 0xFF, (size_t)W * H * sizeof(uint32_t), ctx->stream));
        q.other_key = ctx->other_d.as<uint32_t>();
    }
    if (ctx->tune_wta_seg && (D == 64 || D == 128 || D == 192 || D == 256)) {
        // register march + per-pixel finish (the default where D splits into 16 lanes x 4 / 8 / 12 / 16 cells)
        SVA_TRY(ctx->reserve(ctx->scratch2, (size_t)W * H * sizeof(uint16_t)));
        uint16_t* dwin = ctx->scratch2.as<uint16_t>();
        // pixels per half-warp.  Short segments give more half-warps (the march is bound by the integer pipe and by latency) but pay a flush of
        // the other view's diagonals per segment: measured best 80 - 96 up to c4's size (c1 0.086 -> 0.078 ms, c2 0.119 -> 0.096 ms), 160 at c3
        int seg_len = ctx->tune_wta_seg > 0 ? ctx->tune_wta_seg : ((long long)H * div_up(W, 160) < 32768 ? 96 : 160);
        seg_len = ((seg_len + 15) / 16) * 16;  // a multiple of 16 keeps the unrolled rotation whole
        const int segs = div_up(W, seg_len);
        {
            LaunchScope ls(ctx, "k_wta_seg");
            cudaError_t e = D == 64 ? wta_seg_launch<4>(ctx, vol, q, seg_len, segs, dwin)
                          : D == 128 ? wta_seg_launch<8>(ctx, vol, q, seg_len, segs, dwin)
                          : D == 192 ? wta_seg_launch<12>(ctx, vol, q, seg_len, segs, dwin) : wta_seg_launch<16>(ctx, vol, q, seg_len, segs, dwin);
            SVA_CUDA_OK(ctx, e);
        }
        LaunchScope ls(ctx, "k_wta_finish");
        k_wta_finish<<<dim3(div_up(W, 128), H), 128, 0, ctx->stream>>>(vol, dwin, q.other_key, q);
        SVA_CUDA_OK(ctx, cudaGetLastError());
        return SVA_OK;
    }
    const size_t smem = (size_t)WTA_TX * q.stride_w * 4;
    SVA_CUDA_OK(ctx, cudaFuncSetAttribute(k_wta_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        LaunchScope ls(ctx, "k_wta_tile");
        k_wta_tile<<<dim3(div_up(W, WTA_TX), H), WTA_THREADS, smem, ctx->stream>>>(q);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    if (p.lr_gx != 0) {
        LaunchScope ls(ctx, "k_lr_check");
        k_lr_check<<<dim3(div_up(W, 256), H), 256, 0, ctx->stream>>>(q.other_key, W, H, p.min_disp, p.lr_gx, p.lr_max_diff, q.disp, q.sub);
        SVA_CUDA_OK(ctx, cudaGetLastError());
    }
    return SVA_OK;
}
