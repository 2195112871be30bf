// k_ad2.cu — K1a, image-space form: per-pixel absolute-difference volume summed over camera pairs, for pair offsets |gx|,|gy| <= 2
// (every grid of the configurations; other offsets use the line-image gather kernel in k_ad.cu).
//
//   A(y,x,d) = sum_k |R(y,x) - I_k(y - gy_k*delta, x - gx_k*delta)|,  delta = min_disp + d          (planar layout, see ApGeom)
//
// getAbsDiff's |a-b| (reference src/functions.cpp:215-218) hoisted out of the window loop of src/CameraStereoVision.cpp:76-83.
//
// Bytes run along x: a thread owns FOUR consecutive pixels (one 32-bit word of the reference row) and 16 disparities.  For pair
// (gx,gy) and disparity index i the four source bytes are the word at row y - gy*delta, column x - gx*delta of the other view —
// unaligned by (-gx*i) mod 4, which is a COMPILE-TIME constant once the kernel body is instantiated per (gx,gy) and the 16
// disparities are unrolled: every source word is one or two LDS at immediate offsets plus one PRMT, then VABSDIFF4.U8 and a split
// into two packed-u16 accumulators.  No address arithmetic in the hot loop, no per-byte loads.
//
// A CTA = 8 rows x 128 columns x 32 disparities.  The views live in HBM as zero-bordered, 16-byte-pitched copies (out-of-image
// source = 0 = the spec's OOB value, so there are no bounds checks); the part of every view that the tile's disparity range can
// touch is staged once per CTA with 16-byte cp.async row copies.  The integer ALU pipe is the bound (B200: 64 int lanes per SM).
#include <algorithm>
#include <cstdlib>

#include "sva_common.cuh"

#define AD2_TH 8
#define AD2_TW 128
#define AD2_DR 32
#define AD2_THREADS 512
#define AD2_MAXG 2
#define AD2_SMEM_BUDGET (96 * 1024)

__host__ __device__ constexpr int ad2_sp(int agx) { return agx == 0 ? 160 : (agx == 1 ? 192 : 224); }  // staged row pitch, bytes
__host__ __device__ constexpr int ad2_rows(int agy) { return AD2_TH + (AD2_DR - 1) * agy; }

struct Ad2Params {
    const uint8_t* ref;    // zero-padded reference view, pitch rp
    const uint8_t* imgs;   // zero-bordered other views, one after another
    size_t img_bytes;
    int rp, pp, padx, pady;
    int W, H, D, dmin;
    uint32_t* AP;
    int wp, padl, padt;
    int n;                              // pairs handled by this launch
    int8_t gx[SVA_MAX_PAIRS], gy[SVA_MAX_PAIRS], phi[SVA_MAX_PAIRS];
    uint8_t img[SVA_MAX_PAIRS];         // index of the pair's view in imgs
    int ngroups;                        // pairs are staged in groups that fit the shared-memory budget
    uint8_t gbeg[SVA_MAX_PAIRS + 1];
    int ty0;                            // first tile row of this launch (row-block pipeline; 0 for a whole frame)
};

// accumulate one pair into the thread's 16 disparities x 4 pixels; bp = word-aligned shared pointer of (this row, this quad, disparity 0)
template <int GX, int GY>
__device__ __forceinline__ void ad2_pair(const uint32_t* __restrict__ bp, const uint32_t r, uint32_t (&ae)[16], uint32_t (&ao)[16]) {
    constexpr int SP = ad2_sp(GX < 0 ? -GX : GX);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int boff = -i * (GY * SP + GX);            // byte offset of disparity i's source word
        const int al = ((boff % 4) + 4) % 4;             // its misalignment (compile-time: SP % 4 == 0)
        const int w0 = (boff - al) / 4;
        uint32_t w = bp[w0];
        if (al != 0) w = __funnelshift_r(w, bp[w0 + 1], 8 * al);
        const uint32_t ad = __vabsdiffu4(w, r);
        ae[i] += ad & 0x00FF00FFu;                       // pixels 0 and 2
        ao[i] += __byte_perm(ad, 0, 0x4341);             // pixels 1 and 3
    }
}

template <int GX>
__device__ __forceinline__ void ad2_pair_gy(const int gy, const uint32_t* bp, const uint32_t r, uint32_t (&ae)[16], uint32_t (&ao)[16]) {
    switch (gy) {
        case -2: ad2_pair<GX, -2>(bp, r, ae, ao); break;
        case -1: ad2_pair<GX, -1>(bp, r, ae, ao); break;
        case 0: ad2_pair<GX, 0>(bp, r, ae, ao); break;
        case 1: ad2_pair<GX, 1>(bp, r, ae, ao); break;
        default: ad2_pair<GX, 2>(bp, r, ae, ao); break;
    }
}

__global__ void __launch_bounds__(AD2_THREADS, 2)
k_ad_tile(const Ad2Params q) {
    extern __shared__ __align__(16) unsigned char ad2_smem[];
    __shared__ int s_off[SVA_MAX_PAIRS];  // byte offset of each staged pair's (row 0, disparity-0 column of quad 0) in ad2_smem
    __shared__ const uint8_t* s_src[SVA_MAX_PAIRS];  // first staged byte of the pair's view in global memory
    __shared__ int4 s_geo[SVA_MAX_PAIRS];            // staged rectangle: offset in ad2_smem, pitch, rows
    const int t = threadIdx.x;
    const int quad = t & 31, sub = (t >> 5) & 1, yl = t >> 6;
    const int x0 = blockIdx.x * AD2_TW, y0 = (blockIdx.y + q.ty0) * AD2_TH, da = blockIdx.z * AD2_DR;
    const int y = y0 + yl, x = x0 + 4 * quad;
    const uint32_t r = *reinterpret_cast<const uint32_t*>(q.ref + (size_t)y * q.rp + x);
    uint32_t ae[16], ao[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { ae[i] = 0; ao[i] = 0; }
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(ad2_smem);

    for (int g = 0; g < q.ngroups; g++) {
        const int kb = q.gbeg[g], ke = q.gbeg[g + 1];
        if (g > 0) __syncthreads();  // everyone is done reading the previous group's tiles
        // ---- stage: one thread per pair works out the pair's rectangle; then thread (tr, tc) copies 16-byte chunk tc of rows
        // tr, tr + 32, ... of every pair of the group ----
        if (t < ke - kb) {
            const int k = kb + t;
            const int gx = q.gx[k], gy = q.gy[k];
            int soff = 0;
            for (int j = kb; j < k; j++) soff += ad2_rows(q.gy[j] < 0 ? -q.gy[j] : q.gy[j]) * ad2_sp(q.gx[j] < 0 ? -q.gx[j] : q.gx[j]);
            const int sp = ad2_sp(gx < 0 ? -gx : gx), rows = ad2_rows(gy < 0 ? -gy : gy);
            // first staged row / column in view coordinates (disparity index AD2_DR-1 reaches furthest towards -g)
            const int ylo = y0 - gy * (q.dmin + da) - (gy > 0 ? (AD2_DR - 1) * gy : 0);
            const int v = q.padx + q.phi[k] + x0 - gx * (q.dmin + da) - (gx > 0 ? (AD2_DR - 1) * gx : 0);
            const int c0 = v & ~15, e = v - c0;
            s_off[k] = soff + (gx > 0 ? (AD2_DR - 1) * gx : 0) + e + (gy > 0 ? (AD2_DR - 1) * gy : 0) * sp;
            s_src[k] = q.imgs + (size_t)q.img[k] * q.img_bytes + (size_t)(q.pady + ylo) * q.pp + c0;
            s_geo[k] = make_int4(soff, sp, rows, 0);
        }
        __syncthreads();
        {
            const int tc16 = (t & 15) * 16, tr = t >> 4;
            for (int k = kb; k < ke; k++) {
                const int4 geo = s_geo[k];  // dst offset, pitch, rows
                if (tc16 < geo.y) {
                    const uint8_t* src = s_src[k] + tc16 + (size_t)tr * q.pp;
                    uint32_t dst = smem_base + geo.x + tc16 + tr * geo.y;
                    for (int rr = tr; rr < geo.z; rr += AD2_THREADS / 16) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        src += (size_t)(AD2_THREADS / 16) * q.pp; dst += (AD2_THREADS / 16) * geo.y;
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        // ---- accumulate the group's pairs ----
        for (int k = kb; k < ke; k++) {
            const int gx = q.gx[k], gy = q.gy[k];
            const int sp = ad2_sp(gx < 0 ? -gx : gx);
            // (row yl, quad, disparity index 16*sub): rows move by -gy and columns by -gx per disparity index
            const int off = s_off[k] + (yl - gy * 16 * sub) * sp + 4 * quad - gx * 16 * sub;
            const uint32_t* bp = reinterpret_cast<const uint32_t*>(ad2_smem + off);
            switch (gx) {
                case -2: ad2_pair_gy<-2>(gy, bp, r, ae, ao); break;
                case -1: ad2_pair_gy<-1>(gy, bp, r, ae, ao); break;
                case 0: ad2_pair_gy<0>(gy, bp, r, ae, ao); break;
                case 1: ad2_pair_gy<1>(gy, bp, r, ae, ao); break;
                default: ad2_pair_gy<2>(gy, bp, r, ae, ao); break;
            }
        }
    }
    // ---- store: per disparity pair one 16-byte run of 4 columns in its plane ----
    const int d0 = da + 16 * sub;
    if (y >= q.H || x >= q.W || d0 >= q.D) return;
    uint32_t me = 0xFFFFFFFFu, mo = 0xFFFFFFFFu;  // columns past the right edge stay zero (they are part of the zero border)
    if (x + 3 >= q.W) {
        me = (x + 2 < q.W) ? 0xFFFFFFFFu : 0x0000FFFFu;
        mo = (x + 3 < q.W) ? 0xFFFFFFFFu : ((x + 1 < q.W) ? 0x0000FFFFu : 0u);
    }
    uint32_t* out = q.AP + ((size_t)(y + q.padt) * (q.D >> 1) + (d0 >> 1)) * q.wp + q.padl + x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (d0 + 2 * i >= q.D) break;
        const uint32_t e0 = ae[2 * i] & me, o0 = ao[2 * i] & mo, e1 = ae[2 * i + 1] & me, o1 = ao[2 * i + 1] & mo;
        // (pixel p: disparity 2i | disparity 2i+1 << 16)
        const uint4 v = make_uint4(__byte_perm(e0, e1, 0x5410), __byte_perm(o0, o1, 0x5410), __byte_perm(e0, e1, 0x7632), __byte_perm(o0, o1, 0x7632));
        *reinterpret_cast<uint4*>(out + (size_t)i * q.wp) = v;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
bool sva_ad2_usable(const sva_params& p) {
    for (int i = 0; i < p.n_pairs; i++)
        if (p.pair_gx[i] > AD2_MAXG || p.pair_gx[i] < -AD2_MAXG || p.pair_gy[i] > AD2_MAXG || p.pair_gy[i] < -AD2_MAXG) return false;
    return true;
}

// geometry of the zero-bordered device copies of the views for the current parameters; (re)allocates and zeroes them when it changes
int sva_ad2_prepare(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height;
    const int dpad = p.min_disp + ((p.num_disp + AD2_DR - 1) / AD2_DR) * AD2_DR;  // largest disparity a staged tile can reach (exclusive)
    int mgx = 0, mgy = 0;
    for (int i = 0; i < p.n_pairs; i++) {
        mgx = std::max(mgx, abs(p.pair_gx[i]));
        mgy = std::max(mgy, abs(p.pair_gy[i]));
    }
    Ad2Geom g;
    g.padx = (mgx * dpad + 64 + 15) & ~15;
    g.pady = mgy * dpad + 1;
    const int wt = div_up(W, AD2_TW) * AD2_TW, ht = div_up(H, AD2_TH) * AD2_TH;
    g.pp = (g.padx + wt + g.padx + 64 + 15) & ~15;
    g.rows = g.pady + ht + g.pady;
    g.img_bytes = ((size_t)g.pp * g.rows + 255) & ~(size_t)255;
    g.rp = wt + 16;
    g.ref_rows = ht;
    SVA_TRY(ctx->reserve(ctx->pad_imgs, g.img_bytes * p.n_pairs + 256));
    SVA_TRY(ctx->reserve(ctx->pad_ref, (size_t)g.rp * g.ref_rows + 256));
    uint64_t key = (uint64_t)(uintptr_t)ctx->pad_imgs.p * 31 + (uint64_t)(uintptr_t)ctx->pad_ref.p;
    const int parts[] = {W, H, g.padx, g.pady, g.pp, g.rows, p.n_pairs, p.min_disp};
    for (int v : parts) key = key * 1000003u + (uint64_t)v;
    for (int i = 0; i < p.n_pairs; i++) key = key * 1000003u + (uint64_t)(((p.pair_gx[i] * p.min_disp) % 4 + 4) % 4);
    if (key != ctx->ad2_zero_key) {  // uploads only ever write the interiors
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->pad_imgs.p, 0, g.img_bytes * p.n_pairs, ctx->stream));
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->pad_ref.p, 0, (size_t)g.rp * g.ref_rows, ctx->stream));
        ctx->ad2_zero_key = key;
    }
    ctx->ad2 = g;
    return SVA_OK;
}

// device address of pixel (0,0) of other view k / of the reference view inside the padded buffers (upload targets)
uint8_t* sva_ad2_view_origin(sva_ctx* ctx, int k) {
    const sva_params& p = ctx->prm;
    const int phi = ((p.pair_gx[k] * p.min_disp) % 4 + 4) % 4;
    return ctx->pad_imgs.as<uint8_t>() + (size_t)k * ctx->ad2.img_bytes + (size_t)ctx->ad2.pady * ctx->ad2.pp + ctx->ad2.padx + phi;
}

int sva_ap_prepare(sva_ctx* ctx);

int sva_run_ad2(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    SVA_TRY(sva_ap_prepare(ctx));
    Ad2Params q{};
    q.ref = ctx->pad_ref.as<uint8_t>(); q.imgs = ctx->pad_imgs.as<uint8_t>(); q.img_bytes = ctx->ad2.img_bytes;
    q.rp = ctx->ad2.rp; q.pp = ctx->ad2.pp; q.padx = ctx->ad2.padx; q.pady = ctx->ad2.pady;
    q.W = W; q.H = H; q.D = D; q.dmin = p.min_disp;
    q.AP = ctx->AP.as<uint32_t>(); q.wp = ctx->ap.wp; q.padl = ctx->ap.padl; q.padt = ctx->ap.padt;
    q.n = ctx->pair_end - ctx->pair_begin;
    size_t group_bytes = 0, max_group = 0;
    q.ngroups = 0; q.gbeg[0] = 0;
    for (int i = 0; i < q.n; i++) {
        const int k = ctx->pair_begin + i;
        q.gx[i] = (int8_t)p.pair_gx[k]; q.gy[i] = (int8_t)p.pair_gy[k]; q.img[i] = (uint8_t)k;
        q.phi[i] = (int8_t)(((p.pair_gx[k] * p.min_disp) % 4 + 4) % 4);
        const size_t bytes = (size_t)ad2_rows(abs(p.pair_gy[k])) * ad2_sp(abs(p.pair_gx[k]));
        if (group_bytes + bytes > AD2_SMEM_BUDGET && group_bytes > 0) { q.gbeg[++q.ngroups] = (uint8_t)i; group_bytes = 0; }
        group_bytes += bytes;
        max_group = std::max(max_group, group_bytes);
    }
    q.gbeg[++q.ngroups] = (uint8_t)q.n;
    if (q.n == 0) {  // empty pair range (a pair-sharded rank without pairs): the partial volume is zero
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->AP.p, 0, ctx->ap.words * 4, ctx->stream));
        ctx->have_ad = true;
        return SVA_OK;
    }
    const size_t smem = max_group + 16;
    SVA_CUDA_OK(ctx, cudaFuncSetAttribute(k_ad_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        LaunchScope ls(ctx, "k_ad_tile");
        // a row block needs A on its rows and win_half rows either side (the box window); whole tiles, clipped to the image
        int ya = 0, yb = H;
        if (ctx->win_rows > 0) { ya = std::max(0, ctx->win_y0 - p.win_half); yb = std::min(H, ctx->win_y0 + ctx->win_rows + p.win_half); }
        q.ty0 = ya / AD2_TH;
        k_ad_tile<<<dim3(div_up(W, AD2_TW), div_up(yb, AD2_TH) - q.ty0, div_up(D, AD2_DR)), AD2_THREADS, smem, ctx->stream>>>(q);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    ctx->have_ad = true;
    return SVA_OK;
}
