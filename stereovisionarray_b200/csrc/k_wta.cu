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

#define WTA_TX 64
#define WTA_THREADS 256

struct WtaParams {
    const uint16_t* S;
    int W, H, D, dmin, k;
    int gxp, gxn, gyp, gyn;
    int lr_gx, lr_max_diff, subpixel;
    int stride_w;  // padded pixel stride in 32-bit words
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
            bool ok = x >= k && x < W - k && y >= k && y < H - k;
            if (ok && q.mask) ok = q.mask[(size_t)y * W + x] != 0;
            if (ok) {
                int lim = 0x7FFFFFFF;
                if (q.gxp > 0) lim = min(lim, (x - k) / q.gxp);
                if (q.gxn > 0) lim = min(lim, (W - k - x) / q.gxn);
                if (q.gyp > 0) lim = min(lim, (y - k) / q.gyp);
                if (q.gyn > 0) lim = min(lim, (H - k - y) / q.gyn);
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

// vol = S_total (or C when there is no aggregation); fills ctx->disp / ctx->subpix
int sva_run_wta(sva_ctx* ctx, const uint16_t* vol) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    WtaParams q{};
    q.S = vol; q.W = W; q.H = H; q.D = D; q.dmin = p.min_disp; q.k = p.win_half;
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
