// k_box.cu — K1b: 2k x 2k box sum of the AD volume + shift/cap packing + cell validity.
//
//   C_raw(y,x,d) = sum_{j,i in [-k,k)} A(y+j, x+i, d)            (half-open window — reference src/CameraStereoVision.cpp:57,77)
//   PACK_U16: C = valid ? min(cap, C_raw >> shift) : cap           RAW_U32: C = valid ? C_raw : 0xFFFFFFFF
//
// A CTA owns a strip of TXi = 16*NC columns x 32 disparities and marches down a band of rows.  Each thread keeps the
// vertical running sums V of NC consecutive columns x 2 disparities in registers (add the row entering the window,
// subtract the row leaving it).  The horizontal window sum is a difference of row prefix sums: every thread scans its
// own NC columns in registers, segment totals are exchanged through shared memory, and the prefix row P lives in
// shared memory for the two look-ups P[x+k-1] - P[x-k-1].  All sums are exact u32.
#include "sva_common.cuh"

#define BOX_THREADS 256
#define BOX_DP 16       // d-pairs per CTA  (32 disparities)
#define BOX_CG 16       // column groups per CTA
#define BOX_MAX_BAND 512

struct BoxParams {
    const uint16_t* A;
    void* out;
    int W, H, D, k, dmin, shift, cap;
    int gxp, gxn, gyp, gyn;  // largest positive / negative |offset| over all pairs, per axis (0 = none)
    int band_rows, txo;
    int apply_validity;
};

__device__ __forceinline__ int axis_limit(int pos, int dim, int k, int gp, int gn) {
    // largest delta >= 0 such that k <= pos - g*delta <= dim-k for every pair offset g on this axis; -1 if the pixel itself is outside [k, dim-k)
    if (pos < k || pos >= dim - k) return -1;
    int lim = 0x7FFFFFFF;
    if (gp > 0) lim = min(lim, (pos - k) / gp);
    if (gn > 0) lim = min(lim, (dim - k - pos) / gn);
    return lim;
}

template <int NC, bool RAW>
__global__ void __launch_bounds__(BOX_THREADS)
k_box_cost(BoxParams q) {
    constexpr int TXI = BOX_CG * NC;
    __shared__ uint2 s_P[TXI][BOX_DP];
    __shared__ uint2 s_tot[BOX_CG][BOX_DP];
    __shared__ int s_limy[BOX_MAX_BAND];

    const int t = threadIdx.x, dp = t & (BOX_DP - 1), cg = t >> 4;
    const int W = q.W, H = q.H, D = q.D, k = q.k;
    const int d = blockIdx.y * 32 + 2 * dp;
    const bool d_ok = d < D;
    const int xs = blockIdx.x * q.txo - k;  // global x of strip column 0
    const int y0 = blockIdx.z * q.band_rows, y1 = min(H, y0 + q.band_rows);
    if (y0 >= H) return;

    for (int i = t; i < y1 - y0; i += BOX_THREADS) s_limy[i] = q.apply_validity ? axis_limit(y0 + i, H, k, q.gyp, q.gyn) : 0x7FFFFFFF;

    int limx[NC];
    bool col_in[NC];
    uint32_t v0[NC], v1[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) {
        int x = xs + cg * NC + j;
        col_in[j] = d_ok && x >= 0 && x < W;
        limx[j] = q.apply_validity ? axis_limit(x, W, k, q.gxp, q.gxn) : 0x7FFFFFFF;
        v0[j] = 0; v1[j] = 0;
    }
    const uint16_t* Acol = q.A + ((long long)(xs + cg * NC)) * D + d;  // + r*W*D + j*D
    const long long rowstride = (long long)W * D;

    auto add_row = [&](int r, bool sub) {
        if (r < 0 || r >= H) return;
        uint32_t w[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) w[j] = col_in[j] ? ldg_stream_u32(Acol + r * rowstride + (long long)j * D) : 0u;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            if (sub) { v0[j] -= w[j] & 0xFFFFu; v1[j] -= w[j] >> 16; }
            else     { v0[j] += w[j] & 0xFFFFu; v1[j] += w[j] >> 16; }
        }
    };

    for (int r = y0 - k; r <= y0 + k - 2; r++) add_row(r, false);
    __syncthreads();

    for (int y = y0; y < y1; y++) {
        add_row(y + k - 1, false);
        // prefix over this thread's columns
        uint32_t p0[NC], p1[NC];
        uint32_t a0 = 0, a1 = 0;
#pragma unroll
        for (int j = 0; j < NC; j++) { a0 += v0[j]; a1 += v1[j]; p0[j] = a0; p1[j] = a1; }
        s_tot[cg][dp] = make_uint2(a0, a1);
        __syncthreads();
        uint32_t o0 = 0, o1 = 0;
        for (int g = 0; g < cg; g++) { uint2 tt = s_tot[g][dp]; o0 += tt.x; o1 += tt.y; }
#pragma unroll
        for (int j = 0; j < NC; j++) s_P[cg * NC + j][dp] = make_uint2(p0[j] + o0, p1[j] + o1);
        __syncthreads();
        const int limy = s_limy[y - y0];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            int ci = cg * NC + j, x = xs + ci;
            if (ci < k || ci > TXI - k || x >= W || ci - k >= q.txo || !d_ok) continue;
            uint2 hi = s_P[ci + k - 1][dp];
            uint2 lo = (ci - k - 1 >= 0) ? s_P[ci - k - 1][dp] : make_uint2(0u, 0u);
            uint32_t r0 = hi.x - lo.x, r1 = hi.y - lo.y;
            int lim = min(limx[j], limy);
            bool ok0 = q.dmin + d <= lim, ok1 = q.dmin + d + 1 <= lim;
            long long o = ((long long)y * W + x) * D + d;
            if (RAW) {
                *reinterpret_cast<uint2*>((uint32_t*)q.out + o) = make_uint2(ok0 ? r0 : 0xFFFFFFFFu, ok1 ? r1 : 0xFFFFFFFFu);
            } else {
                uint32_t c0 = ok0 ? min((uint32_t)q.cap, r0 >> q.shift) : (uint32_t)q.cap;
                uint32_t c1 = ok1 ? min((uint32_t)q.cap, r1 >> q.shift) : (uint32_t)q.cap;
                *reinterpret_cast<uint32_t*>((uint16_t*)q.out + o) = c0 | (c1 << 16);
            }
        }
        add_row(y - k, true);
    }
}

// A -> out (u16 packed or u32 raw).  D must be even.  Used by the volume pipeline and (raw, no validity) by literal mode / refine.
int sva_launch_box(sva_ctx* ctx, const uint16_t* A, void* out, int W, int H, int D, int k, const sva_params* prm, bool raw, bool apply_validity) {
    if (k < 1 || k > 56) return ctx->fail(SVA_ERR_BAD_ARG, "win_half must be in 1..56");
    BoxParams q{};
    q.A = A; q.out = out; q.W = W; q.H = H; q.D = D; q.k = k;
    q.dmin = prm ? prm->min_disp : 0; q.shift = prm ? prm->cost_shift : 0; q.cap = prm ? prm->cost_cap : SVA_COST_CAP_MAX;
    q.apply_validity = apply_validity ? 1 : 0;
    if (prm && apply_validity) {
        for (int i = 0; i < prm->n_pairs; i++) {
            int gx = prm->pair_gx[i], gy = prm->pair_gy[i];
            if (gx > 0) q.gxp = gx > q.gxp ? gx : q.gxp;
            if (gx < 0) q.gxn = -gx > q.gxn ? -gx : q.gxn;
            if (gy > 0) q.gyp = gy > q.gyp ? gy : q.gyp;
            if (gy < 0) q.gyn = -gy > q.gyn ? -gy : q.gyn;
        }
    }
    const bool wide = k > 12;
    const int txi = wide ? 128 : 64;
    q.txo = txi - 2 * k + 1;
    int strips = div_up(W, q.txo), dch = div_up(D, 32);
    // enough CTAs to fill the machine twice over, bands no shorter than 4k rows (warm-up is 2k-1 rows of loads)
    int bands = div_up(2 * ctx->sm_count * 2, strips * dch);
    int min_band = 4 * k > 32 ? 4 * k : 32;
    if (bands > H / min_band) bands = H / min_band;
    if (bands < 1) bands = 1;
    if (div_up(H, bands) > BOX_MAX_BAND) bands = div_up(H, BOX_MAX_BAND);
    q.band_rows = div_up(H, bands);
    bands = div_up(H, q.band_rows);
    dim3 grid(strips, dch, bands);
    LaunchScope ls(ctx, raw ? "k_box_cost_raw" : "k_box_cost");
    if (wide) {
        if (raw) k_box_cost<8, true><<<grid, BOX_THREADS, 0, ctx->stream>>>(q);
        else k_box_cost<8, false><<<grid, BOX_THREADS, 0, ctx->stream>>>(q);
    } else {
        if (raw) k_box_cost<4, true><<<grid, BOX_THREADS, 0, ctx->stream>>>(q);
        else k_box_cost<4, false><<<grid, BOX_THREADS, 0, ctx->stream>>>(q);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    return SVA_OK;
}

int sva_run_box(sva_ctx* ctx, bool raw) {
    const sva_params& p = ctx->prm;
    size_t cells = (size_t)p.width * p.height * p.num_disp;
    if (raw) {
        SVA_TRY(ctx->reserve(ctx->Craw, cells * sizeof(uint32_t)));
        return sva_launch_box(ctx, ctx->A.as<uint16_t>(), ctx->Craw.p, p.width, p.height, p.num_disp, p.win_half, &p, true, true);
    }
    SVA_TRY(ctx->reserve(ctx->C, cells * sizeof(uint16_t)));
    SVA_TRY(sva_launch_box(ctx, ctx->A.as<uint16_t>(), ctx->C.p, p.width, p.height, p.num_disp, p.win_half, &p, false, true));
    ctx->have_cost = true;
    return SVA_OK;
}
