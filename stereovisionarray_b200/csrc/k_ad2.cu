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
// disparities are unrolled: every source word is one or two LDS at immediate offsets plus one funnel shift, then VABSDIFF4.U8.
// The four byte differences go into two accumulators WITHOUT being split first: X += ad (a plain 32-bit add, the bytes bleed into
// each other on purpose) and Y += (pixels 1 and 3 as u16x2, one PRMT); because X = E + 256 * Y (mod 2^32) with E = (pixels 0 and 2
// as u16x2), the even pixels are recovered once per tile row as X - (Y << 8).  Per source word that is VABSDIFF4 + PRMT on the
// integer ALU pipe (the bound: 2 warp-instructions / clk / SM) and two adds — written as IMADs with a run-time 1 in the kernels compiled
// per pair set, so that they stay on the FMA pipe (left to itself ptxas folds them into 3-input IADD3s on the integer pipe).
//
// A CTA = TH rows x 128 columns x 32 disparities, TH = 8 * R.  The views live in HBM as zero-bordered, 16-byte-pitched copies
// (out-of-image source = 0 = the spec's OOB value, so there are no bounds checks); the part of every view that the tile's disparity
// range can touch (TH + 31 * |gy| rows) is staged once per CTA with 16-byte cp.async row copies, and the 512 threads then walk the
// tile 8 rows at a time.  Tall tiles amortise the halo rows and the per-CTA set-up; the host picks TH (24 - 32 rows while the grid
// keeps five waves of the 2 CTAs per SM: sva_run_ad2).
//
// Three kernels: k_ad_tile_set<SET> (the pair set known at compile time: straight-line code, no dispatch — the 3 x 3 array, the
// rectified pair and, with 16 disparities per CTA, the 4 x 4 array), k_ad_tile<false> (any pair set that fits one staging group:
// a 25-way dispatch per pair) and k_ad_tile<true> (several staging groups, 8-row tiles, the accumulators live across the groups).
#include <algorithm>
#include <cstdlib>

#include "sva_common.cuh"

#define AD2_TH 8
#define AD2_TW 128
#define AD2_DR 32
#define AD2_THREADS 512
#define AD2_MAXG 2
#define AD2_SMEM_BUDGET (110 * 1024)   // two CTAs per SM

// staged row pitch, bytes: 128 columns + (dr - 1) * |gx| + up to 15 bytes of alignment + 3 of the last unaligned word, in whole 16-byte chunks
__host__ __device__ constexpr int ad2_sp(int agx, int dr = AD2_DR) { return dr == 32 ? (agx == 0 ? 160 : (agx == 1 ? 192 : 224)) : (agx == 0 ? 160 : (agx == 1 ? 176 : 192)); }
__host__ __device__ constexpr int ad2_rows(int agy, int th, int dr = AD2_DR) { return th + (dr - 1) * agy; }
#define AD2_TH_MAX 48

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
    int row0, row1;                     // image rows [row0, row1) of this launch (row0 a multiple of 8; a row block or the whole frame)
    int th;                             // rows per tile (a multiple of 8; 8 when the pairs need more than one staging group)
    uint32_t one;                       // 1, as a run-time value: acc = x * one + acc is an IMAD on the FMA pipe, where acc += x may become an
                                        // IADD3 on the integer ALU pipe — the pipe that bounds this kernel (VABSDIFF4, PRMT, SHF live there)
};

// accumulate one pair into the thread's 16 disparities x 4 pixels; bp = word-aligned shared pointer of (this row, this quad, disparity 0)
// ax[i] += the word of four byte differences (fields bleed), ao[i] += pixels 1 and 3 as u16x2; see the header
template <int GX, int GY, bool FMA = false, int DR = AD2_DR>
__device__ __forceinline__ void ad2_pair(const uint32_t* __restrict__ bp, const uint32_t r, uint32_t (&ax)[16], uint32_t (&ao)[16], const uint32_t one = 1u) {
    constexpr int SP = ad2_sp(GX < 0 ? -GX : GX, DR);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int boff = -i * (GY * SP + GX);            // byte offset of disparity i's source word
        const int al = ((boff % 4) + 4) % 4;             // its misalignment (compile-time: SP % 4 == 0)
        const int w0 = (boff - al) / 4;
        uint32_t w = bp[w0];
        if (al != 0) w = __funnelshift_r(w, bp[w0 + 1], 8 * al);
        const uint32_t ad = __vabsdiffu4(w, r);
        if (FMA) {
            ax[i] = ad * one + ax[i];
            ao[i] = __byte_perm(ad, 0, 0x4341) * one + ao[i];
        } else {
            ax[i] += ad;
            ao[i] += __byte_perm(ad, 0, 0x4341);         // pixels 1 and 3
        }
    }
}

template <int GX>
__device__ __forceinline__ void ad2_pair_gy(const int gy, const uint32_t* bp, const uint32_t r, uint32_t (&ax)[16], uint32_t (&ao)[16]) {
    switch (gy) {
        case -2: ad2_pair<GX, -2>(bp, r, ax, ao); break;
        case -1: ad2_pair<GX, -1>(bp, r, ax, ao); break;
        case 0: ad2_pair<GX, 0>(bp, r, ax, ao); break;
        case 1: ad2_pair<GX, 1>(bp, r, ax, ao); break;
        default: ad2_pair<GX, 2>(bp, r, ax, ao); break;
    }
}

#define AD2_DISPATCH(type, bp, r, ax, ao)                              \
    switch (type) {                                                    \
        case 0: ad2_pair<-2, -2>(bp, r, ax, ao); break;                \
        case 1: ad2_pair<-2, -1>(bp, r, ax, ao); break;                \
        case 2: ad2_pair<-2, 0>(bp, r, ax, ao); break;                 \
        case 3: ad2_pair<-2, 1>(bp, r, ax, ao); break;                 \
        case 4: ad2_pair<-2, 2>(bp, r, ax, ao); break;                 \
        case 5: ad2_pair<-1, -2>(bp, r, ax, ao); break;                \
        case 6: ad2_pair<-1, -1>(bp, r, ax, ao); break;                \
        case 7: ad2_pair<-1, 0>(bp, r, ax, ao); break;                 \
        case 8: ad2_pair<-1, 1>(bp, r, ax, ao); break;                 \
        case 9: ad2_pair<-1, 2>(bp, r, ax, ao); break;                 \
        case 10: ad2_pair<0, -2>(bp, r, ax, ao); break;                \
        case 11: ad2_pair<0, -1>(bp, r, ax, ao); break;                \
        case 12: ad2_pair<0, 0>(bp, r, ax, ao); break;                 \
        case 13: ad2_pair<0, 1>(bp, r, ax, ao); break;                 \
        case 14: ad2_pair<0, 2>(bp, r, ax, ao); break;                 \
        case 15: ad2_pair<1, -2>(bp, r, ax, ao); break;                \
        case 16: ad2_pair<1, -1>(bp, r, ax, ao); break;                \
        case 17: ad2_pair<1, 0>(bp, r, ax, ao); break;                 \
        case 18: ad2_pair<1, 1>(bp, r, ax, ao); break;                 \
        case 19: ad2_pair<1, 2>(bp, r, ax, ao); break;                 \
        case 20: ad2_pair<2, -2>(bp, r, ax, ao); break;                \
        case 21: ad2_pair<2, -1>(bp, r, ax, ao); break;                \
        case 22: ad2_pair<2, 0>(bp, r, ax, ao); break;                 \
        case 23: ad2_pair<2, 1>(bp, r, ax, ao); break;                 \
        default: ad2_pair<2, 2>(bp, r, ax, ao); break;                 \
    }

// per staged pair: byte offset of (tile row 0, disparity-0 column of quad 0) in ad2_smem, row pitch, 16 * (gy * pitch + gx) = what 16
// disparity indices move the source by, body index (gx + 2) * 5 + gy + 2
struct Ad2Shared {
    int4 pk[SVA_MAX_PAIRS];
    const uint8_t* src[SVA_MAX_PAIRS];  // first staged byte of the pair's view in global memory
    int2 geo[SVA_MAX_PAIRS];            // staged rectangle: offset in ad2_smem, rows
};

// stage pairs [kb, ke): one thread per pair works out the pair's rectangle, then every thread copies 16-byte chunks t, t + 512, ... of
// each rectangle (chunk c = row c / cpr, column chunk c % cpr); returns with the copies complete and visible to the CTA
template <int DR = AD2_DR>
__device__ __forceinline__ void ad2_stage(const Ad2Params& q, Ad2Shared& sh, unsigned char* smem, const int kb, const int ke, const int x0, const int y0, const int da) {
    const int t = threadIdx.x, th = q.th;
    if (t < ke - kb) {
        const int k = kb + t;
        const int gx = q.gx[k], gy = q.gy[k];
        int soff = 0;
        for (int j = kb; j < k; j++) soff += ad2_rows(q.gy[j] < 0 ? -q.gy[j] : q.gy[j], th, DR) * ad2_sp(q.gx[j] < 0 ? -q.gx[j] : q.gx[j], DR);
        const int sp = ad2_sp(gx < 0 ? -gx : gx, DR), rows = ad2_rows(gy < 0 ? -gy : gy, th, DR);
        // first staged row / column in view coordinates (disparity index AD2_DR-1 reaches furthest towards -g)
        const int ylo = y0 - gy * (q.dmin + da) - (gy > 0 ? (DR - 1) * gy : 0);
        const int v = q.padx + q.phi[k] + x0 - gx * (q.dmin + da) - (gx > 0 ? (DR - 1) * gx : 0);
        const int c0 = v & ~15, e = v - c0;
        sh.pk[k] = make_int4(soff + (gx > 0 ? (DR - 1) * gx : 0) + e + (gy > 0 ? (DR - 1) * gy : 0) * sp, sp, 16 * (gy * sp + gx), (gx + 2) * 5 + gy + 2);
        sh.src[k] = q.imgs + (size_t)q.img[k] * q.img_bytes + (size_t)(q.pady + ylo) * q.pp + c0;
        sh.geo[k] = make_int2(soff, rows);
    }
    __syncthreads();
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    for (int k = kb; k < ke; k++) {
        const int2 geo = sh.geo[k];  // dst offset, rows
        const int sp = sh.pk[k].y, cpr = sp >> 4, n = geo.y * cpr;
        const uint32_t inv = 65536u / (uint32_t)cpr + 1u;  // c / cpr == (c * inv) >> 16 for c < 4681 (cpr <= 14; n <= 110 rows x 14)
        const uint8_t* src0 = sh.src[k];
        const uint32_t dst0 = smem_base + geo.x;
        for (int c = t; c < n; c += AD2_THREADS) {
            const int rr = (int)(((uint32_t)c * inv) >> 16), cc = (c - rr * cpr) << 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + rr * sp + cc), "l"(src0 + (size_t)rr * q.pp + cc) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}

// store the thread's 4 columns x 16 disparities of image row y: per disparity pair one 16-byte run of 4 columns in its plane
__device__ __forceinline__ void ad2_store(const Ad2Params& q, const int y, const int x, const int d0, const uint32_t (&ax)[16], const uint32_t (&ao)[16]) {
    if (y >= q.H || x >= q.W || d0 >= q.D) return;
    uint32_t* out = q.AP + ((size_t)(y + q.padt) * (q.D >> 1) + (d0 >> 1)) * q.wp + q.padl + x;
    uint32_t me = 0xFFFFFFFFu, mo = 0xFFFFFFFFu;  // columns past the right edge stay zero (they are part of the zero border)
    if (x + 3 >= q.W) {
        me = (x + 2 < q.W) ? 0xFFFFFFFFu : 0x0000FFFFu;
        mo = (x + 1 < q.W) ? 0x0000FFFFu : 0u;
    }
    const int np = min(8, (q.D - d0) >> 1);  // disparity pairs of this thread inside the range
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t o0 = ao[2 * i] & mo, o1 = ao[2 * i + 1] & mo;
        const uint32_t e0 = (ax[2 * i] - (ao[2 * i] << 8)) & me, e1 = (ax[2 * i + 1] - (ao[2 * i + 1] << 8)) & me;  // pixels 0 and 2 (header)
        // (pixel p: disparity 2i | disparity 2i+1 << 16)
        const uint4 v = make_uint4(__byte_perm(e0, e1, 0x5410), __byte_perm(o0, o1, 0x5410), __byte_perm(e0, e1, 0x7632), __byte_perm(o0, o1, 0x7632));
        if (i < np) *reinterpret_cast<uint4*>(out + (size_t)i * q.wp) = v;
    }
}

// MULTI = false: all pairs fit one staging group; the tile is q.th rows, walked 8 rows at a time.
// MULTI = true:  several staging groups (c3: 15 pairs); the tile is 8 rows and the accumulators live across the groups.
template <bool MULTI>
__global__ void __launch_bounds__(AD2_THREADS, 2)
k_ad_tile(const Ad2Params q) {
    extern __shared__ __align__(16) unsigned char ad2_smem[];
    __shared__ Ad2Shared sh;
    const int t = threadIdx.x;
    const int quad = t & 31, sub = (t >> 5) & 1, yl = t >> 6;
    const int x0 = blockIdx.x * AD2_TW, y0 = q.row0 + blockIdx.y * q.th, da = blockIdx.z * AD2_DR;
    const int x = x0 + 4 * quad, d0 = da + 16 * sub;
    // this thread's constant part of every source address: (row yl, quad) and 16 * sub disparity indices folded in per pair
    uint32_t ax[16], ao[16];
    if (MULTI) {
        const int y = y0 + yl;
        const uint32_t r = *reinterpret_cast<const uint32_t*>(q.ref + (size_t)y * q.rp + x);
#pragma unroll
        for (int i = 0; i < 16; i++) { ax[i] = 0; ao[i] = 0; }
        for (int g = 0; g < q.ngroups; g++) {
            const int kb = q.gbeg[g], ke = q.gbeg[g + 1];
            if (g > 0) __syncthreads();  // everyone is done reading the previous group's tiles
            ad2_stage(q, sh, ad2_smem, kb, ke, x0, y0, da);
#pragma unroll 1
            for (int k = kb; k < ke; k++) {
                const int4 pk = sh.pk[k];
                const uint32_t* bp = reinterpret_cast<const uint32_t*>(ad2_smem + (pk.x + yl * pk.y - sub * pk.z + 4 * quad));
                AD2_DISPATCH(pk.w, bp, r, ax, ao)
            }
        }
        ad2_store(q, y, x, d0, ax, ao);
    } else {
        ad2_stage(q, sh, ad2_smem, 0, q.n, x0, y0, da);
        const int n = q.n;
        const int rows = min(q.th, q.row1 - y0);  // rows of this tile inside the launch's range, walked in whole groups of 8
#pragma unroll 1
        for (int yt = yl; yt - yl < rows; yt += AD2_TH) {
            const int y = y0 + yt;
            const uint32_t r = *reinterpret_cast<const uint32_t*>(q.ref + (size_t)y * q.rp + x);
#pragma unroll
            for (int i = 0; i < 16; i++) { ax[i] = 0; ao[i] = 0; }
#pragma unroll 1
            for (int k = 0; k < n; k++) {
                const int4 pk = sh.pk[k];
                // (row yt, quad, disparity index 16*sub): rows move by -gy and columns by -gx per disparity index
                const uint32_t* bp = reinterpret_cast<const uint32_t*>(ad2_smem + (pk.x + yt * pk.y - sub * pk.z + 4 * quad));
                AD2_DISPATCH(pk.w, bp, r, ax, ao)
            }
            ad2_store(q, y, x, d0, ax, ao);
        }
    }
}

// ---- the same kernel with the pair set known at compile time ----------------------------------------------------------------------
// SET::n pairs with grid offsets SET::gx[k], SET::gy[k], one staging group.  With the set fixed the pair loop is straight-line code: no
// dispatch, no common accumulate block that the 25 bodies of the generic kernel jump to (and that forces all 32 of a body's results to be
// live at once), and the scheduler can overlap one pair's shared-memory loads with the previous pair's arithmetic; the accumulation is
// written as IMADs (Ad2Params::one).  DR = disparities per CTA: 32 (a thread = 4 pixels x 16 disparities, two threads per quad and row,
// 8 rows per pass) or 16 (one thread per quad and row, 16 rows per pass) — the 15 pairs of a 4 x 4 array reach two baselines, and only
// with 16 disparities per CTA do their staged rows (th + 15 * |gy| instead of th + 31 * |gy|) fit one group at th = 16.
// Instantiated for the grids of the configurations (sva_run_ad2); every other pair set runs k_ad_tile.
struct Ad2Set3x3 {  // 3 x 3 array, reference in the centre, views in index order (c1, c2, c4)
    static constexpr int n = 8, dr = 32;
    static constexpr int gx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    static constexpr int gy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
};
struct Ad2SetPair {  // rectified pair, the other camera one baseline to the left (c0)
    static constexpr int n = 1, dr = 32;
    static constexpr int gx[1] = {-1};
    static constexpr int gy[1] = {0};
};
struct Ad2Set4x4 {  // 4 x 4 array, reference at (1, 1), views in index order (c3)
    static constexpr int n = 15, dr = 16;
    static constexpr int gx[15] = {-1, 0, 1, 2, -1, 1, 2, -1, 0, 1, 2, -1, 0, 1, 2};
    static constexpr int gy[15] = {-1, -1, -1, -1, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2};
};

template <class SET, int K>
struct Ad2Seq {
    static __device__ __forceinline__ void run(const Ad2Shared& sh, const unsigned char* smem, const int yt, const int sub, const int quad, const uint32_t r,
                                               uint32_t (&ax)[16], uint32_t (&ao)[16], const uint32_t one) {
        if constexpr (K < SET::n) {
            const int4 pk = sh.pk[K];
            const uint32_t* bp = reinterpret_cast<const uint32_t*>(smem + (pk.x + yt * pk.y - sub * pk.z + 4 * quad));
            ad2_pair<SET::gx[K], SET::gy[K], true, SET::dr>(bp, r, ax, ao, one);
            Ad2Seq<SET, K + 1>::run(sh, smem, yt, sub, quad, r, ax, ao, one);
        }
    }
};

template <class SET>
__global__ void __launch_bounds__(AD2_THREADS, 2)
k_ad_tile_set(const Ad2Params q) {
    constexpr int DR = SET::dr, SUBS = DR / 16, RSTEP = AD2_THREADS / (32 * SUBS);  // rows per pass: 8 (DR = 32) or 16 (DR = 16)
    extern __shared__ __align__(16) unsigned char ad2_smem[];
    __shared__ Ad2Shared sh;
    const int t = threadIdx.x;
    const int quad = t & 31, sub = (t >> 5) % SUBS, yl = t / (32 * SUBS);
    const int x0 = blockIdx.x * AD2_TW, y0 = q.row0 + blockIdx.y * q.th, da = blockIdx.z * DR;
    const int x = x0 + 4 * quad, d0 = da + 16 * sub;
    ad2_stage<DR>(q, sh, ad2_smem, 0, SET::n, x0, y0, da);
    const int rows = min(q.th, q.row1 - y0);  // rows of this tile inside the launch's range, walked in whole passes
#pragma unroll 1
    for (int yt = yl; yt - yl < rows; yt += RSTEP) {
        const int y = y0 + yt;
        const uint32_t r = *reinterpret_cast<const uint32_t*>(q.ref + (size_t)y * q.rp + x);
        uint32_t ax[16], ao[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { ax[i] = 0; ao[i] = 0; }
        Ad2Seq<SET, 0>::run(sh, ad2_smem, yt, sub, quad, r, ax, ao, q.one);
        ad2_store(q, y, x, d0, ax, ao);
    }
}

template <class SET>
static bool ad2_set_matches(const Ad2Params& q) {
    if (q.n != SET::n) return false;
    for (int i = 0; i < SET::n; i++)
        if (q.gx[i] != SET::gx[i] || q.gy[i] != SET::gy[i]) return false;
    return true;
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
    const int wt = div_up(W, AD2_TW) * AD2_TW, ht = div_up(H, AD2_TH) * AD2_TH + AD2_TH_MAX;  // the last tile of a launch may hang over by up to a tile
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
    q.one = 1u;
    for (int i = 0; i < q.n; i++) {
        const int k = ctx->pair_begin + i;
        q.gx[i] = (int8_t)p.pair_gx[k]; q.gy[i] = (int8_t)p.pair_gy[k]; q.img[i] = (uint8_t)k;
        q.phi[i] = (int8_t)(((p.pair_gx[k] * p.min_disp) % 4 + 4) % 4);
    }
    if (q.n == 0) {  // empty pair range (a pair-sharded rank without pairs): the partial volume is zero
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->AP.p, 0, ctx->ap.words * 4, ctx->stream));
        ctx->have_ad = true;
        return SVA_OK;
    }
    // a row block needs A on its rows and win_half rows either side (the box window); whole 8-row groups, clipped to the image
    int ya = 0, yb = H;
    if (ctx->win_rows > 0) { ya = std::max(0, ctx->win_y0 - p.win_half); yb = std::min(H, ctx->win_y0 + ctx->win_rows + p.win_half); }
    q.row0 = ya / AD2_TH * AD2_TH; q.row1 = yb;
    // The kernel: one compiled for the frame's pair set where there is one (it also fixes the disparities per CTA), else the generic one.
    void (*kern)(Ad2Params) = nullptr;
    int dr = AD2_DR;
    if (ctx->tune_ad_set) {
        if (ad2_set_matches<Ad2Set3x3>(q)) { kern = k_ad_tile_set<Ad2Set3x3>; dr = Ad2Set3x3::dr; }
        else if (ad2_set_matches<Ad2SetPair>(q)) { kern = k_ad_tile_set<Ad2SetPair>; dr = Ad2SetPair::dr; }
        else if (ad2_set_matches<Ad2Set4x4>(q)) { kern = k_ad_tile_set<Ad2Set4x4>; dr = Ad2Set4x4::dr; }
    }
    const int rstep = dr == 32 ? 8 : 16;  // rows per pass of the CTA's 512 threads
    // Tile height: a tile stages th + (dr - 1) * |gy| rows of every view, so tall tiles cost less staging and set-up per row — but the grid
    // should keep several waves of the 2 CTAs per SM.  Measured (B200, c1 / c4): 24 - 32 rows are best as long as the grid still has five
    // or more waves; small frames keep one pass per tile.  A pair set without a compiled kernel whose 8-row tile does not fit one staging
    // group runs the generic kernel with several groups.
    auto bytes_for = [&](int th) {
        size_t b = 0;
        for (int i = 0; i < q.n; i++) b += (size_t)ad2_rows(abs(q.gy[i]), th, dr) * ad2_sp(abs(q.gx[i]), dr);
        return b;
    };
    const int tiles_x = div_up(W, AD2_TW), tiles_d = div_up(D, dr), slots = 2 * ctx->sm_count;
    if (kern && bytes_for(rstep) > AD2_SMEM_BUDGET) { kern = nullptr; dr = AD2_DR; }
    int th = kern ? rstep : AD2_TH;
    if (ctx->tune_ad_th > 0) {
        const int want = std::min(AD2_TH_MAX, ctx->tune_ad_th / th * th);
        if (want >= th && bytes_for(want) <= AD2_SMEM_BUDGET) th = want;
    } else {
        for (int cand = 32; cand > th; cand -= (kern ? rstep : AD2_TH)) {
            if (bytes_for(cand) > AD2_SMEM_BUDGET) continue;
            if (tiles_x * tiles_d * div_up(q.row1 - q.row0, cand) >= 5 * slots) { th = cand; break; }
        }
    }
    q.th = th;
    size_t group_bytes = 0, max_group = 0;
    q.ngroups = 0; q.gbeg[0] = 0;
    for (int i = 0; i < q.n; i++) {
        const size_t bytes = (size_t)ad2_rows(abs(q.gy[i]), th, dr) * ad2_sp(abs(q.gx[i]), dr);
        if (group_bytes + bytes > AD2_SMEM_BUDGET && group_bytes > 0) { q.gbeg[++q.ngroups] = (uint8_t)i; group_bytes = 0; }
        group_bytes += bytes;
        max_group = std::max(max_group, group_bytes);
    }
    q.gbeg[++q.ngroups] = (uint8_t)q.n;
    const size_t smem = max_group + 16;
    const char* label = kern ? "k_ad_tile_set" : "k_ad_tile";
    if (!kern) kern = q.ngroups > 1 ? k_ad_tile<true> : k_ad_tile<false>;
    SVA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        LaunchScope ls(ctx, label);
        kern<<<dim3(tiles_x, div_up(q.row1 - q.row0, th), tiles_d), AD2_THREADS, smem, ctx->stream>>>(q);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    ctx->have_ad = true;
    return SVA_OK;
}
