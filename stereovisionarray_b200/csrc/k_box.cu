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

#define BOX_STAGES 3

// 16-byte async copy global -> shared; src_bytes == 0 zero-fills (out-of-image rows / columns / disparities)
__device__ __forceinline__ void box_cp16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int NC, bool RAW>
__global__ void __launch_bounds__(BOX_THREADS, NC == 8 ? 3 : 4)
k_box_cost(BoxParams q) {
    constexpr int TXI = BOX_CG * NC;
    constexpr int ROW_BYTES = TXI * 64;                 // one tile row: TXI columns x 32 disparities x u16
    constexpr int STAGE_BYTES = 2 * ROW_BYTES;          // [entering row][leaving row]
    extern __shared__ __align__(16) unsigned char box_smem[];
    unsigned char* s_ring = box_smem;                                         // BOX_STAGES * STAGE_BYTES
    uint2 (*s_P)[BOX_DP] = reinterpret_cast<uint2 (*)[BOX_DP]>(box_smem + BOX_STAGES * STAGE_BYTES);      // [TXI][BOX_DP]
    uint2 (*s_tot)[BOX_DP] = reinterpret_cast<uint2 (*)[BOX_DP]>(box_smem + BOX_STAGES * STAGE_BYTES + TXI * BOX_DP * 8);  // [BOX_CG][BOX_DP]
    int* s_limy = reinterpret_cast<int*>(box_smem + BOX_STAGES * STAGE_BYTES + TXI * BOX_DP * 8 + BOX_CG * BOX_DP * 8);

    const int t = threadIdx.x, dp = t & (BOX_DP - 1), cg = t >> 4;
    const int W = q.W, H = q.H, D = q.D, k = q.k;
    const int d0 = blockIdx.y * 32;
    const int d = d0 + 2 * dp;
    const bool d_ok = d < D;
    const int xs = blockIdx.x * q.txo - k;  // global x of strip column 0
    const int y0 = blockIdx.z * q.band_rows, y1 = min(H, y0 + q.band_rows);
    if (y0 >= H) return;

    for (int i = t; i < y1 - y0; i += BOX_THREADS) s_limy[i] = q.apply_validity ? axis_limit(y0 + i, H, k, q.gyp, q.gyn) : 0x7FFFFFFF;

    int limx[NC];
    uint32_t v0[NC], v1[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) {
        int x = xs + cg * NC + j;
        limx[j] = q.apply_validity ? axis_limit(x, W, k, q.gxp, q.gxn) : 0x7FFFFFFF;
        v0[j] = 0; v1[j] = 0;
    }
    const long long rowstride = (long long)W * D;
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(s_ring);

    // cooperative stage fill: the tile row is TXI * 4 chunks of 16 B (8 disparities); thread t copies chunks t, t + 256, ...
    auto fill = [&](int it) {  // iteration it: entering row y0 + it + k - 1, leaving row y0 + it - k (only once it >= 0)
        const uint32_t st = ring_base + ((it + 2 * k) % BOX_STAGES) * STAGE_BYTES;
        const int ra = y0 + it + k - 1, rs = y0 + it - k;
        const bool ra_ok = ra >= 0 && ra < H, rs_ok = it >= 0 && rs >= 0 && rs < H;
#pragma unroll
        for (int c = t; c < TXI * 4; c += BOX_THREADS) {
            const int col = c >> 2, part = c & 3, x = xs + col, dd = d0 + 8 * part;
            const bool in = x >= 0 && x < W && dd < D;
            const long long off = ((long long)(in ? x : 0)) * D + (in ? dd : 0);
            box_cp16(st + c * 16, q.A + (ra_ok ? ra : 0) * rowstride + off, (in && ra_ok) ? 16 : 0);
            if (it >= 0) box_cp16(st + ROW_BYTES + c * 16, q.A + (rs_ok ? rs : 0) * rowstride + off, (in && rs_ok) ? 16 : 0);
        }
    };

    const int it_begin = -(2 * k - 1), it_end = y1 - y0;  // warm-up iterations (it < 0) only accumulate
#pragma unroll
    for (int i = 0; i < BOX_STAGES - 1; i++) {
        if (it_begin + i < it_end) fill(it_begin + i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    for (int it = it_begin; it < it_end; it++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(BOX_STAGES - 2) : "memory");
        __syncthreads();  // stage `it` is complete for every thread; everyone has finished reading stage it-1
        if (it + BOX_STAGES - 1 < it_end) fill(it + BOX_STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const unsigned char* st = s_ring + ((it + 2 * k) % BOX_STAGES) * STAGE_BYTES;
        const uint32_t* rowa = reinterpret_cast<const uint32_t*>(st) + (cg * NC) * 16 + dp;  // [col][16 u32]
#pragma unroll
        for (int j = 0; j < NC; j++) { uint32_t w = rowa[j * 16]; v0[j] += w & 0xFFFFu; v1[j] += w >> 16; }
        if (it < 0) continue;
        const int y = y0 + it;
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
        const int limy = s_limy[it];
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
        const uint32_t* rows = reinterpret_cast<const uint32_t*>(st + ROW_BYTES) + (cg * NC) * 16 + dp;
#pragma unroll
        for (int j = 0; j < NC; j++) { uint32_t w = rows[j * 16]; v0[j] -= w & 0xFFFFu; v1[j] -= w >> 16; }
    }
}

template <int NC, bool RAW>
static cudaError_t box_launch(const BoxParams& q, dim3 grid, cudaStream_t stream) {
    constexpr int TXI = BOX_CG * NC;
    const size_t smem = (size_t)BOX_STAGES * 2 * TXI * 64 + (size_t)TXI * BOX_DP * 8 + BOX_CG * BOX_DP * 8 + BOX_MAX_BAND * 4;
    cudaError_t e = cudaFuncSetAttribute(k_box_cost<NC, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_box_cost<NC, RAW><<<grid, BOX_THREADS, smem, stream>>>(q);
    return cudaGetLastError();
}

// A -> out (u16 packed or u32 raw).  D must be a multiple of 8 (16-byte async copies).  Used by the volume pipeline and (raw, no validity) by literal mode / refine.
int sva_launch_box(sva_ctx* ctx, const uint16_t* A, void* out, int W, int H, int D, int k, const sva_params* prm, bool raw, bool apply_validity) {
    if (k < 1 || k > 56) return ctx->fail(SVA_ERR_BAD_ARG, "win_half must be in 1..56");
    if (D % 8) return ctx->fail(SVA_ERR_BAD_ARG, "box filter needs a multiple of 8 planes");
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
    // enough CTAs for one full wave, bands no shorter than 4k rows (warm-up is 2k-1 rows of loads)
    int bands = div_up(2 * ctx->sm_count * 2, strips * dch);
    int min_band = 4 * k > 32 ? 4 * k : 32;
    if (bands > H / min_band) bands = H / min_band;
    if (bands < 1) bands = 1;
    if (div_up(H, bands) > BOX_MAX_BAND) bands = div_up(H, BOX_MAX_BAND);
    q.band_rows = div_up(H, bands);
    bands = div_up(H, q.band_rows);
    dim3 grid(strips, dch, bands);
    LaunchScope ls(ctx, raw ? "k_box_cost_raw" : "k_box_cost");
    cudaError_t e;
    if (wide) e = raw ? box_launch<8, true>(q, grid, ctx->stream) : box_launch<8, false>(q, grid, ctx->stream);
    else e = raw ? box_launch<4, true>(q, grid, ctx->stream) : box_launch<4, false>(q, grid, ctx->stream);
    SVA_CUDA_OK(ctx, e);
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
