// k_sgm.cu — K2: semi-global path aggregation (K3 — WTA / left-right check / sub-pixel — follows in k_wta.cu; the fused
// last-pass form is kept as variant A).
//
// No counterpart in the reference: this implements the frozen spec of DESIGN.md §3.3 (Hirschmueller 2008,
// fixed P1/P2), bit-exact against oracle/sva_oracle.c.
//
//   L_r(p,d) = C(p,d) + min(L_r(q,d), L_r(q,d-1)+P1, L_r(q,d+1)+P1, min_k L_r(q,k)+P2) - min_k L_r(q,k),  q = p - r
//   S = sum_r L_r (u16; C <= 4095 and P2 <= 4095 bound L by 8190 and S over 8 paths by 65520)
//
// Mapping: ONE WARP PER PATH LINE per direction.  A lane owns n = 2*NR consecutive disparities packed two per register
// (u16x2); the recurrence is DPX (VIADDMNMX.U16x2 / VIMNMX.U16x2), the d-1 / d+1 neighbours are one PRMT per register plus
// two warp shuffles for the lane edges, and min_k is a VIMNMX3 fold followed by one CREDUX.MIN.  Everything stays in
// registers along the path; C is streamed by cp.async (LDGSTS) into a per-warp ring of PF + 1 shared-memory stages, S is
// accumulated with fire-and-forget 64-bit REDs (packed u16x4 adds are carry-free by the bound above) into a zeroed volume.
// Diagonal paths use W lines of exactly H steps that wrap around the image edge and restart (L = C) where the predecessor
// is outside the image, so every line of a launch has the same length.
//
// Default schedule for 8 paths (sva_run_sgm, variant D): three launches — the three directions that sweep the rows
// downwards, the three that sweep upwards, the two horizontal ones.  All lines of a row-sweeping launch advance one row per
// step, so a row's C and S lines are touched by all three directions while they are L2-resident (DRAM sees C once and S once
// per launch); when an image row of C + S is 768 KB or more the CTAs are additionally paced against the grid-wide minimum.
//
// Variant A's last pass (k_sgm_pass<FINAL>, horizontal): S_total = S + L in registers -> packed (S<<16|d) keys -> REDUX.MIN
// gives the first-minimum winner; S(d*-1), S(d*+1) are fetched with two shuffles for the parabola; the other view's WTA
// (left-right check) is a systolic diagonal minimum: one key register per disparity slot shifts by one slot per step.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "sva_common.cuh"
#include "sva_vec.cuh"

// PF = steps of prefetch in flight per warp; the ring has PF + 1 stages (the slot refilled at step s was last read at step s-1)
#define SGM_INF2 0x7FFF7FFFu
#ifndef SGM_GROUP_REDUX
#define SGM_GROUP_REDUX 0  // REDUX with per-group masks compiles to a divergent slow path (CREDUX writes ONE uniform register per warp)
#endif
#define SGM_WARPS 8
#define SGM_FINAL_WARPS 4

enum { SGM_MODE_STORE = 0, SGM_MODE_RED = 1, SGM_MODE_FINAL = 2 };

struct SgmParams {
    const uint16_t* C;
    uint16_t* S;
    int W, H, D;
    int ndirs;          // directions processed concurrently by this launch: CTA b handles direction b % ndirs
    int dxs[8], dys[8];
    uint32_t p1p1, p2p2;
    int lanes;  // active lanes = D / (2*NR)
    // final pass only
    int dmin, k, gxp, gxn, gyp, gyn, lr_gx, lr_max_diff, subpixel, store_full, no_agg;
    int wta_only;       // the march only reads S_total (all paths already accumulated) and does K3
    unsigned int* pace_arrive;  // k_sgm_acc: [pace_rounds] CTAs that finished round r (nullptr = no global pacing)
    unsigned int* pace_min;     // rounds finished by every CTA
    int pace_rounds, pace_window, march_warps;
    int diag_split;     // k_sgm_acc, one line per warp: diagonals run the event-split march (SVA_SGM_DIAG_SPLIT, default on)
    int c_ds;           // 0: C is [H][W][D]; > 0: C is slice-major [D / c_ds][H][W][c_ds] (disparity slices gathered from several GPUs)
    int exp_no_out;     // timing experiment only (SVA_SGM_EXP=1): run the recurrence but drop the S updates
    int cta_sync;       // k_sgm_acc (balanced): named barrier among the row-sweeping warps of a CTA every round
    int balanced;       // k_sgm_acc: grid = m * SM count; warp w of CTA b handles direction w % ndirs, line b + grid * (w / ndirs)
    const uint8_t* mask;
    uint16_t* disp;
    float* sub;
};

// one step of the recurrence for this lane's 2*NR disparities; L holds L(q,.) on entry and L(p,.) on exit
// LPL = lanes per path line: 32 (one line per warp) or 16 / 8 (two / four lines per warp — more cells per lane, so the
// shuffles, the minimum reduction and the loop overhead are amortised over more cells)
// EDGE_BIAS (every lane of the group active): instead of replacing the missing d-1 / d+1 neighbour of the first / last disparity by
// +inf with two selects per step, the lane adds a per-lane P1 whose edge half is 0x7FFF (p1_up, p1_dn): the neighbour slot then holds a
// bounded real value (<= 8190), 8190 + 0x7FFF < 2^16 does not wrap and is larger than every real cost, so the minimum ignores it.
template <int NR, int LPL = 32, bool EDGE_BIAS = false>
__device__ __forceinline__ void sgm_step(uint32_t (&L)[NR], const uint32_t (&Cc)[NR], uint32_t& mm, uint32_t& mp2, uint32_t p1p1, uint32_t p2p2,
                                         bool first_lane, bool last_lane, uint32_t p1_up = 0, uint32_t p1_dn = 0) {
    uint32_t up = __shfl_up_sync(0xffffffffu, L[NR - 1], 1, LPL);
    uint32_t dn = __shfl_down_sync(0xffffffffu, L[0], 1, LPL);
    if (!EDGE_BIAS) {
        if (first_lane) up = SGM_INF2;
        if (last_lane) dn = SGM_INF2;
        p1_up = p1p1; p1_dn = p1p1;
    }
    uint32_t sh[NR + 1];  // sh[j] = values at d-1 of register j; sh[j+1] = values at d+1 of register j
    sh[0] = __byte_perm(up, L[0], 0x5432);
#pragma unroll
    for (int j = 1; j < NR; j++) sh[j] = __byte_perm(L[j - 1], L[j], 0x5432);
    sh[NR] = __byte_perm(L[NR - 1], dn, 0x5432);
    uint32_t mloc = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < NR; j++) {
        uint32_t t = __viaddmin_u16x2(sh[j], j == 0 ? p1_up : p1p1, L[j]);
        t = __viaddmin_u16x2(sh[j + 1], j == NR - 1 ? p1_dn : p1p1, t);
        t = __vminu2(t, mp2);
        L[j] = Cc[j] + t - mm;  // both halves: t >= mm, no borrow; C + t - mm <= 8190, no carry
        mloc = __vminu2(mloc, L[j]);
    }
    uint32_t m;
    if (LPL == 32) {
        m = min(mloc & 0xFFFFu, mloc >> 16);
        m = __reduce_min_sync(0xffffffffu, m);
    } else if (SGM_GROUP_REDUX) {  // one REDUX per group: the groups of a warp pass disjoint member masks
        m = min(mloc & 0xFFFFu, mloc >> 16);
        const unsigned lane = threadIdx.x & 31u;
        m = __reduce_min_sync(((1u << LPL) - 1u) << (lane & ~(unsigned)(LPL - 1)), m);
    } else {
#pragma unroll
        for (int o = LPL / 2; o > 0; o >>= 1) mloc = __vminu2(mloc, __shfl_xor_sync(0xffffffffu, mloc, o, LPL));
        m = min(mloc & 0xFFFFu, mloc >> 16);
    }
    mm = m * 0x10001u;
    mp2 = mm + p2p2;
}

int sva_run_wta(sva_ctx* ctx, const uint16_t* vol);

struct PathPos {
    int x, y;
};

template <int NR, int MODE, int SGM_PF>
__global__ void __launch_bounds__(MODE == SGM_MODE_FINAL ? SGM_FINAL_WARPS * 32 : SGM_WARPS * 32)
k_sgm_pass(SgmParams q) {
    constexpr int SGM_NS = SGM_PF + 1;
    constexpr int NV = 2 * NR;
    constexpr int WARPS = MODE == SGM_MODE_FINAL ? SGM_FINAL_WARPS : SGM_WARPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dir = blockIdx.x % q.ndirs;
    const int line = (blockIdx.x / q.ndirs) * WARPS + warp;
    const int W = q.W, H = q.H, D = q.D, dx = q.dxs[dir], dy = q.dys[dir];
    const int nlines = dy == 0 ? H : W, len = dy == 0 ? W : H;
    if (line >= nlines) return;
    const bool active = lane < q.lanes;
    const bool first_lane = lane == 0, last_lane = lane == q.lanes - 1;
    const bool diag = dx != 0 && dy != 0;

    PathPos pos, pre;  // current cell and prefetch cursor
    if (dy == 0) { pos.y = line; pos.x = dx > 0 ? 0 : W - 1; }
    else { pos.y = dy > 0 ? 0 : H - 1; pos.x = line; }
    pre = pos;
    const int lane_off = lane * NV;
    auto cell = [&](const PathPos& p) -> long long { return ((long long)p.y * W + p.x) * D + lane_off; };
    auto advance = [&](PathPos& p) -> bool {  // returns true when the new cell starts a fresh path (predecessor outside the image)
        p.y += dy; p.x += dx;
        if (dy != 0) {
            if (p.x >= W) { p.x = 0; return true; }
            if (p.x < 0) { p.x = W - 1; return true; }
        }
        return false;
    };

    // ---- streaming: cp.async (LDGSTS) into a per-warp ring of SGM_NS stages; wait_group gives FIFO completion, which the
    // register scoreboard cannot (a register prefetch ring deeper than the scoreboard slots serialises on DRAM latency) ----
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int STAGE_BYTES = 32 * NV * 2;                         // one step of one volume for the whole warp
    constexpr int NVOL = MODE == SGM_MODE_FINAL ? 2 : 1;             // final pass streams C and S
    constexpr int RING_BYTES = SGM_NS * NVOL * STAGE_BYTES;
    const uint32_t ring = smem_u32(smem_raw) + warp * RING_BYTES + lane * (NV * 2);
    const bool wta_only = MODE == SGM_MODE_FINAL && q.wta_only;
    const bool stream_s = MODE == SGM_MODE_FINAL && !q.no_agg && !wta_only;
    const uint16_t* vol0 = wta_only ? q.S : q.C;
    auto issue = [&](int slot) {  // loads for the cell under the prefetch cursor into ring slot `slot`
        if (active) {
            Vec<NR>::cp_async(ring + slot * NVOL * STAGE_BYTES, vol0 + cell(pre));
            if (stream_s) Vec<NR>::cp_async(ring + slot * NVOL * STAGE_BYTES + STAGE_BYTES, q.S + cell(pre));
        }
    };
#pragma unroll
    for (int u = 0; u < SGM_PF; u++) {
        if (u < len) { issue(u); advance(pre); }
        cp_async_commit();
    }

    uint32_t L[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) L[j] = 0;
    uint32_t mm = 0, mp2 = q.p2p2;
    bool restart = true;

    // ---- final-pass state ----
    uint16_t* pend_d = nullptr; float* pend_sub = nullptr; uint16_t* other_row = nullptr;
    uint32_t acc[NV];
    const int tdir = q.lr_gx * dx;  // +1: LR entries travel towards larger d; -1: towards smaller d
    if (MODE == SGM_MODE_FINAL) {
        unsigned char* base = smem_raw + WARPS * RING_BYTES + (size_t)warp * (8 * (size_t)W);
        pend_sub = reinterpret_cast<float*>(base);
        pend_d = reinterpret_cast<uint16_t*>(base + 4 * (size_t)W);
        other_row = reinterpret_cast<uint16_t*>(base + 6 * (size_t)W);
        for (int x = lane; x < W; x += 32) other_row[x] = 0xFFFFu;
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] = 0xFFFFFFFFu;
        __syncwarp();
    }

    for (int s0 = 0; s0 < len; s0 += SGM_NS) {
#pragma unroll
        for (int u = 0; u < SGM_NS; u++) {
            const int s = s0 + u;   // step s lives in ring slot s % SGM_NS == u
            if (s >= len) break;
            cp_async_wait<SGM_PF - 1>();
            uint32_t Cc[NR], Sp[NR];
#pragma unroll
            for (int j = 0; j < NR; j++) { Cc[j] = SGM_INF2; Sp[j] = 0; }
            if (active) {
                Vec<NR>::lds(ring + u * NVOL * STAGE_BYTES, Cc);
                if (stream_s) Vec<NR>::lds(ring + u * NVOL * STAGE_BYTES + STAGE_BYTES, Sp);
            }
            // refill the slot that was consumed one step ago with step s + PF
            if (s + SGM_PF < len) { issue((u + SGM_PF) % SGM_NS); advance(pre); }
            cp_async_commit();
            if (restart || (MODE == SGM_MODE_FINAL && q.no_agg)) {
#pragma unroll
                for (int j = 0; j < NR; j++) L[j] = 0;
                mm = 0; mp2 = q.p2p2;
            }
            if (!wta_only) sgm_step<NR>(L, Cc, mm, mp2, q.p1p1, q.p2p2, first_lane, last_lane);

            if (MODE == SGM_MODE_STORE) {
                if (active) Vec<NR>::store(q.S + cell(pos), L);
            } else if (MODE == SGM_MODE_RED) {
                if (active) Vec<NR>::red(q.S + cell(pos), L);
            } else {
                // ---- fused K3 ----
                uint32_t St[NR];
#pragma unroll
                for (int j = 0; j < NR; j++) St[j] = wta_only ? Cc[j] : (q.no_agg ? L[j] : Sp[j] + L[j]);
                if (q.store_full && !wta_only && active) Vec<NR>::store(q.S + cell(pos), St);
                uint32_t key[NV];
                uint32_t kbest = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    key[2 * j] = active ? ((St[j] << 16) | (uint32_t)(lane_off + 2 * j)) : 0xFFFFFFFFu;
                    key[2 * j + 1] = active ? ((St[j] & 0xFFFF0000u) | (uint32_t)(lane_off + 2 * j + 1)) : 0xFFFFFFFFu;
                    kbest = min(kbest, min(key[2 * j], key[2 * j + 1]));
                }
                kbest = __reduce_min_sync(0xffffffffu, kbest);
                const int dstar = (int)(kbest & 0xFFFFu);
                float fsub = (float)dstar;
                if (q.subpixel && dstar > 0 && dstar < D - 1) {  // warp-uniform
                    const int dl = dstar - 1, dr = dstar + 1;
                    uint32_t vl = St[0], vr = St[0];
#pragma unroll
                    for (int j = 1; j < NR; j++) {
                        if (((dl >> 1) % NR) == j) vl = St[j];
                        if (((dr >> 1) % NR) == j) vr = St[j];
                    }
                    vl = __shfl_sync(0xffffffffu, vl, dl / NV);
                    vr = __shfl_sync(0xffffffffu, vr, dr / NV);
                    const int sl = (dl & 1) ? (int)(vl >> 16) : (int)(vl & 0xFFFFu);
                    const int sr = (dr & 1) ? (int)(vr >> 16) : (int)(vr & 0xFFFFu);
                    const int s0v = (int)(kbest >> 16);
                    const int den = sl - 2 * s0v + sr;
                    if (den > 0) fsub = (float)dstar + (float)(sl - sr) / (float)(2 * den);
                }
                if (lane == 0) { pend_d[pos.x] = (uint16_t)dstar; pend_sub[pos.x] = fsub; }
                if (q.lr_gx != 0) {
                    if (tdir > 0) {
                        uint32_t carry = __shfl_up_sync(0xffffffffu, acc[NV - 1], 1);
                        if (first_lane) carry = 0xFFFFFFFFu;
#pragma unroll
                        for (int i = NV - 1; i > 0; i--) acc[i] = acc[i - 1];
                        acc[0] = carry;
                    } else {
                        uint32_t carry = __shfl_down_sync(0xffffffffu, acc[0], 1);
                        if (last_lane) carry = 0xFFFFFFFFu;
#pragma unroll
                        for (int i = 0; i < NV - 1; i++) acc[i] = acc[i + 1];
                        acc[NV - 1] = carry;
                    }
#pragma unroll
                    for (int i = 0; i < NV; i++) acc[i] = min(acc[i], key[i]);
                    // the entry leaving the volume this step is complete
                    if (tdir > 0 && last_lane) {
                        int xo = pos.x - q.lr_gx * (q.dmin + D - 1);
                        if (xo >= 0 && xo < W) other_row[xo] = (uint16_t)(acc[NV - 1] & 0xFFFFu);
                    }
                    if (tdir < 0 && first_lane) {
                        int xo = pos.x - q.lr_gx * q.dmin;
                        if (xo >= 0 && xo < W) other_row[xo] = (uint16_t)(acc[0] & 0xFFFFu);
                    }
                }
            }
            if (s + 1 < len) restart = advance(pos);
        }
    }

    if (MODE == SGM_MODE_FINAL) {
        // flush the LR entries still inside the volume at the end of the row (they have seen every in-image contribution)
        if (q.lr_gx != 0 && active) {
#pragma unroll
            for (int i = 0; i < NV; i++) {
                int xo = pos.x - q.lr_gx * (q.dmin + lane_off + i);
                if (xo >= 0 && xo < W && acc[i] != 0xFFFFFFFFu) other_row[xo] = (uint16_t)(acc[i] & 0xFFFFu);
            }
        }
        __syncwarp();
        const int y = pos.y, k = q.k;
        int limy = 0x7FFFFFFF;
        bool row_in = y >= k && y < H - k;
        if (q.gyp > 0) limy = min(limy, (y - k) / q.gyp);
        if (q.gyn > 0) limy = min(limy, (H - k - y) / q.gyn);
        for (int x = lane; x < W; x += 32) {
            const int d = pend_d[x];
            const int delta = q.dmin + d;
            bool ok = row_in && x >= k && x < W - k;
            if (ok && q.mask) ok = q.mask[(size_t)y * W + x] != 0;
            if (ok) {
                int lim = limy;
                if (q.gxp > 0) lim = min(lim, (x - k) / q.gxp);
                if (q.gxn > 0) lim = min(lim, (W - k - x) / q.gxn);
                ok = delta <= lim;
            }
            if (ok && q.lr_gx != 0) {
                int xo = x - q.lr_gx * delta;
                if (xo < 0 || xo >= W) ok = false;
                else {
                    int od = other_row[xo];
                    ok = od != 0xFFFF && abs(d - od) <= q.lr_max_diff;
                }
            }
            q.disp[(size_t)y * W + x] = ok ? (uint16_t)delta : (uint16_t)SVA_DISP_INVALID;
            if (q.sub) q.sub[(size_t)y * W + x] = ok ? (float)q.dmin + pend_sub[x] : SVA_SUBPIX_INVALID;
        }
    }
}

// ---- lean accumulate march (the hot kernel): same recurrence and ring as above, specialised at compile time on
// DIAG (wrap/restart logic only for diagonals), FULL (all 32 lanes active: no predicates) and STORE (plain store vs RED),
// running 32-bit element cursors instead of recomputed cell indices, and an unchecked steady-state loop with a checked tail.
template <int NR, int PF, bool FULL, bool DIAG, bool STORE, int LPL>
__device__ __forceinline__ void sgm_acc_march(const SgmParams& q, const int dx, const int dy, const int line, const bool do_out, const int lane, const uint32_t ring,
                                              const int bar_threads, volatile int* s_pace /* [0] rounds finished by this CTA, [1] rounds finished by every CTA */,
                                              const bool leader) {
    constexpr int NS = PF + 1, NV = 2 * NR, STAGE = 32 * NV * 2;
    const int W = q.W, H = q.H, D = q.D;
    const int len = dy == 0 ? W : H;
    const int x0 = dy == 0 ? (dx > 0 ? 0 : W - 1) : line, y0 = dy == 0 ? line : (dy > 0 ? 0 : H - 1);
    // Cursors are 32-bit ELEMENT indices into the volumes (W*H*D < 2^32; the 64-bit address is one IMAD.WIDE on the FMA pipe), and
    // a diagonal's wrap at the image edge is a countdown instead of two coordinate compares: the recurrence saturates the integer
    // ALU pipe, so every ALU instruction shaved off the cursor bookkeeping is time.
    const uint32_t dstep = (uint32_t)((dy * W + dx) * D), wrapfix = (uint32_t)(-dx * W * D);
    const int lin = lane % LPL;  // lane within this line's group (LPL lanes per line, 32 / LPL lines per warp)
    const uint32_t start = (uint32_t)(((long long)y0 * W + x0) * D + lin * NV);
    // the cost volume may be slice-major (a lane's cells never straddle a slice: c_ds % NV == 0): same walk, pixel stride c_ds
    const int cds = q.c_ds > 0 ? q.c_ds : D;
    const uint32_t dstep_c = (uint32_t)((dy * W + dx) * cds), wrapfix_c = (uint32_t)(-dx * W * cds);
    const uint32_t start_c = q.c_ds > 0 ? (uint32_t)((long long)((lin * NV) / cds) * H * W * cds + ((long long)y0 * W + x0) * cds + (lin * NV) % cds) : start;
    uint32_t ic = start_c, is = start;                // prefetch cursor (C), accumulate cursor (S)
    int cc = dx > 0 ? W - x0 : x0 + 1, cs = cc;       // steps until each cursor leaves the image sideways
    const bool active = FULL || lin < q.lanes;
    const bool first_lane = lin == 0, last_lane = FULL ? lin == LPL - 1 : lin == q.lanes - 1;
    const uint32_t p1_up = first_lane ? (q.p1p1 & 0xFFFF0000u) | 0x7FFFu : q.p1p1, p1_dn = last_lane ? (q.p1p1 & 0x0000FFFFu) | 0x7FFF0000u : q.p1p1;
    auto adv = [&](uint32_t& i, int& cnt, const uint32_t step, const uint32_t fix) -> bool {
        i += step;
        if (DIAG) {
            if (--cnt == 0) { cnt = W; i += fix; return true; }
        }
        return false;
    };
#pragma unroll
    for (int u = 0; u < PF; u++) {
        if (u < len) { if (active) Vec<NR>::cp_async(ring + u * STAGE, q.C + ic); adv(ic, cc, dstep_c, wrapfix_c); }
        cp_async_commit();
    }
    uint32_t L[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) L[j] = 0;
    uint32_t mm = 0, mp2 = q.p2p2;
    bool restart = false;  // step 0 starts from L = 0, mm = 0, which yields L = C
    auto step = [&](const uint32_t slot_addr, const uint32_t refill_addr, const bool refill) {
        cp_async_wait<PF - 1>();
        uint32_t Cc[NR];
#pragma unroll
        for (int j = 0; j < NR; j++) Cc[j] = SGM_INF2;
        if (active) Vec<NR>::lds(slot_addr, Cc);
        if (refill) { if (active) Vec<NR>::cp_async(refill_addr, q.C + ic); adv(ic, cc, dstep_c, wrapfix_c); }
        cp_async_commit();
        if (DIAG && restart) {
#pragma unroll
            for (int j = 0; j < NR; j++) L[j] = 0;
            mm = 0; mp2 = q.p2p2;
        }
        sgm_step<NR, LPL, FULL>(L, Cc, mm, mp2, q.p1p1, q.p2p2, first_lane, last_lane, p1_up, p1_dn);
        if (active && do_out) { if (STORE) Vec<NR>::store(q.S + is, L); else Vec<NR>::red(q.S + is, L); }
        restart = adv(is, cs, dstep, wrapfix);
    };
    int s0 = 0;
    for (; s0 + NS + PF <= len; s0 += NS) {
        // keep the row-sweeping warps of this CTA in step (one named barrier per 9 rows): directions that sweep the rows in the
        // same order then touch the C and S lines they share within microseconds of each other, i.e. while they are in L2
        if (bar_threads) {
            asm volatile("bar.sync 1, %0;" ::"r"(bar_threads) : "memory");
            if (s_pace) {
                // global pacing (asynchronous): the helper warp publishes this CTA's progress and mirrors the grid-wide minimum into
                // shared memory; a CTA that is more than pace_window rounds ahead of the slowest one naps.  Bounded spin: never hangs.
                const int round = s0 / NS;
                if (leader && lane == 0) s_pace[0] = round;
                for (int spin = 0; spin < 4096 && round > s_pace[1] + q.pace_window; spin++) __nanosleep(128);
            }
        }
#pragma unroll
        for (int u = 0; u < NS; u++) step(ring + u * STAGE, ring + ((u + PF) % NS) * STAGE, true);
    }
    if (s_pace && leader && lane == 0) s_pace[0] = s0 / NS;  // all paced rounds done (lets the helper warp finish)
    for (int s = s0; s < len; s++) step(ring + (s % NS) * STAGE, ring + ((s + PF) % NS) * STAGE, s + PF < len);
}

// Diagonal march for ONE line per warp, split at the wrap events.  A diagonal leaves the image sideways once every W steps; the
// prefetch cursor does so PF steps before the accumulate cursor.  Between two such events nothing special happens, so the step body
// carries no wrap / restart logic at all (in the generic march that logic is ~8 predicated integer instructions per step on a kernel
// bound by the integer pipe): whole ring rounds of NS steps run unrolled, the few steps around an event run one at a time.  The wrap
// step is a per-line constant, i.e. warp-uniform here.  The CTA barrier / pacing rhythm (every NS steps, same step indices in every
// warp) is kept, so the warps of a CTA still meet the same number of times.
template <int NR, int PF, bool FULL, bool STORE>
__device__ __forceinline__ void sgm_acc_march_diag32(const SgmParams& q, const int dx, const int dy, const int line, const bool do_out, const int lane,
                                                     const uint32_t ring, const int bar_threads, volatile int* s_pace, const bool leader) {
    constexpr int NS = PF + 1, NV = 2 * NR, STAGE = 32 * NV * 2;
    const int W = q.W, H = q.H, D = q.D;
    const int len = H;
    const int x0 = line, y0 = dy > 0 ? 0 : H - 1;
    const uint32_t dstep = (uint32_t)((dy * W + dx) * D), wrapfix = (uint32_t)(-dx * W * D);
    const uint32_t start = (uint32_t)(((long long)y0 * W + x0) * D + lane * NV);
    const int cds = q.c_ds > 0 ? q.c_ds : D;
    const uint32_t dstep_c = (uint32_t)((dy * W + dx) * cds), wrapfix_c = (uint32_t)(-dx * W * cds);
    const uint32_t start_c = q.c_ds > 0 ? (uint32_t)((long long)((lane * NV) / cds) * H * W * cds + ((long long)y0 * W + x0) * cds + (lane * NV) % cds) : start;
    uint32_t ic = start_c, is = start;
    const bool active = FULL || lane < q.lanes;
    const bool first_lane = lane == 0, last_lane = FULL ? lane == 31 : lane == q.lanes - 1;
    const uint32_t p1_up = first_lane ? (q.p1p1 & 0xFFFF0000u) | 0x7FFFu : q.p1p1, p1_dn = last_lane ? (q.p1p1 & 0x0000FFFFu) | 0x7FFF0000u : q.p1p1;
    int wc = dx > 0 ? W - x0 : x0 + 1;  // the prefetch cursor wraps after this many advances ...
    int ws = wc;                        // ... the accumulate cursor after this many (then every W more)
    int adv_c = 0;
#pragma unroll
    for (int u = 0; u < PF; u++) {
        if (u < len) {
            if (active) Vec<NR>::cp_async(ring + u * STAGE, q.C + ic);
            ic += dstep_c;
            if (++adv_c == wc) { ic += wrapfix_c; wc += W; }
        }
        cp_async_commit();
    }
    uint32_t L[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) L[j] = 0;
    uint32_t mm = 0, mp2 = q.p2p2;
    auto step = [&](const uint32_t slot_addr, const uint32_t refill_addr, const bool refill) {
        cp_async_wait<PF - 1>();
        uint32_t Cc[NR];
#pragma unroll
        for (int j = 0; j < NR; j++) Cc[j] = SGM_INF2;
        if (active) Vec<NR>::lds(slot_addr, Cc);
        if (refill) { if (active) Vec<NR>::cp_async(refill_addr, q.C + ic); ic += dstep_c; }
        cp_async_commit();
        sgm_step<NR, 32, FULL>(L, Cc, mm, mp2, q.p1p1, q.p2p2, first_lane, last_lane, p1_up, p1_dn);
        if (active && do_out) { if (STORE) Vec<NR>::store(q.S + is, L); else Vec<NR>::red(q.S + is, L); }
        is += dstep;
    };
    int s = 0, slot = 0, round = 0;
    while (s < len) {
        // last step before the next event: the prefetch cursor has advanced PF + s + 1 times after step s (while it refills)
        const int e_c = wc - PF - 1, e_s = ws - 1;
        const int e = min(len - 1, min(e_c, e_s));
        while (s <= e) {
            const bool paced = slot == 0 && s + NS + PF <= len;
            if (paced && bar_threads) {
                asm volatile("bar.sync 1, %0;" ::"r"(bar_threads) : "memory");
                if (s_pace) {
                    if (leader && lane == 0) s_pace[0] = round;
                    for (int spin = 0; spin < 4096 && round > s_pace[1] + q.pace_window; spin++) __nanosleep(128);
                }
            }
            if (paced) round++;
            if (paced && s + NS - 1 <= e) {  // a whole ring round without an event: every step refills
#pragma unroll
                for (int u = 0; u < NS; u++) step(ring + u * STAGE, ring + ((u + PF) % NS) * STAGE, true);
                s += NS;
            } else {
                const int rs = slot + PF >= NS ? slot + PF - NS : slot + PF;
                step(ring + slot * STAGE, ring + rs * STAGE, s + PF < len);
                s++;
                slot = slot + 1 == NS ? 0 : slot + 1;
            }
        }
        if (e == e_c) { ic += wrapfix_c; wc += W; }
        if (e == e_s) {  // the next cell starts a fresh path: L = C
            is += wrapfix; ws += W;
#pragma unroll
            for (int j = 0; j < NR; j++) L[j] = 0;
            mm = 0; mp2 = q.p2p2;
        }
    }
    if (s_pace && leader && lane == 0) s_pace[0] = round;  // all paced rounds done (lets the helper warp finish)
}

// 40 registers: 48 resident warps per SM (e.g. two 22-warp CTAs of a paced c4 launch) must fit the 64 K register file
template <int NR, int PF, bool FULL, bool STORE, int LPL>
__global__ void __maxnreg__(LPL == 32 && NR <= 4 ? 40 : 64)
k_sgm_acc(SgmParams q) {
    constexpr int LPW = 32 / LPL;  // path lines per warp
    constexpr int RING_BYTES = (PF + 1) * 32 * 2 * NR * 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ volatile int s_pace[2];
    if (q.pace_arrive) {
        if (threadIdx.x == 0) { s_pace[0] = 0; s_pace[1] = 0; }
        __syncthreads();
        if (warp == q.march_warps) {
            // helper warp: all global pacing traffic happens here, off the marching warps' critical path.  arrive[r] counts the CTAs
            // that finished round r; whoever arrives last advances the grid-wide minimum.
            if (lane == 0) {
                int published = 0;
                for (int it = 0; it < (1 << 22) && published < q.pace_rounds; it++) {
                    const int r = s_pace[0];
                    for (; published < r; published++) {
                        unsigned int old = atomicAdd(q.pace_arrive + published, 1u);
                        if (old == gridDim.x - 1) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(q.pace_min), "r"(published + 1) : "memory"); }
                    }
                    unsigned int g;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(q.pace_min) : "memory");
                    s_pace[1] = (int)g;
                    __nanosleep(1000);
                }
            }
            return;
        }
    }
    int dir, wline;  // wline = index of this warp's group of LPW consecutive lines
    if (q.balanced) {  // every CTA carries the same mix of directions and every SM the same number of CTAs -> all lines advance at the same rate
        dir = warp % q.ndirs;
        wline = blockIdx.x + gridDim.x * (warp / q.ndirs);
    } else {
        dir = blockIdx.x % q.ndirs; wline = (blockIdx.x / q.ndirs) * q.march_warps + warp;
    }
    const int dx = q.dxs[dir], dy = q.dys[dir];
    const int nlines = dy == 0 ? q.H : q.W;
    if (wline * LPW >= nlines) return;
    int line = wline * LPW + lane / LPL;
    const bool do_out = line < nlines && !q.exp_no_out;  // a ragged last group recomputes the last line and drops the result
    line = min(line, nlines - 1);
    const uint32_t ring = smem_u32(smem_raw) + warp * RING_BYTES + lane * (4 * NR);
    int bar_threads = 0;
    bool leader = false;
    if (q.balanced && q.cta_sync && dy != 0) {  // threads of this CTA that march down/up the rows (all do the same number of rounds)
        int first = -1;
        for (int w = 0; w < q.march_warps; w++)
            if (q.dys[w % q.ndirs] != 0 && (int)(blockIdx.x + gridDim.x * (w / q.ndirs)) * LPW < q.W) { bar_threads += 32; if (first < 0) first = w; }
        leader = warp == first;
    }
    volatile int* pace = (q.pace_arrive && bar_threads) ? s_pace : nullptr;
    if (dx != 0 && dy != 0) {
        if (LPL == 32 && q.diag_split) sgm_acc_march_diag32<NR, PF, FULL, STORE>(q, dx, dy, line, do_out, lane, ring, bar_threads, pace, leader);
        else sgm_acc_march<NR, PF, FULL, true, STORE, LPL>(q, dx, dy, line, do_out, lane, ring, bar_threads, pace, leader);
    } else sgm_acc_march<NR, PF, FULL, false, STORE, LPL>(q, dx, dy, line, do_out, lane, ring, bar_threads, pace, leader);
}

template <int NR, int PF, bool FULL, bool STORE, int LPL>
static int launch_acc(sva_ctx* ctx, const SgmParams& q, int nlines_all, size_t ring_smem_per_warp, const char* name) {
    const int nlines = div_up(nlines_all, 32 / LPL);  // warp-lines
    int warps = SGM_WARPS, grid = div_up(nlines, SGM_WARPS) * q.ndirs;
    SgmParams qq = q;
    qq.balanced = 0;
    qq.cta_sync = ctx->tune_sgm_cta_sync;
    qq.diag_split = ctx->tune_sgm_diag_split;
    if (ctx->tune_sgm_balanced) {
        // one wave of identical CTAs: m CTAs per SM, as few warps per CTA as cover all lines
        for (int m = 1; m <= 8; m++) {
            const int g = ctx->sm_count * m;
            const int w = div_up(nlines, g) * q.ndirs;  // lines per direction per CTA x directions
            if (w <= 32 && (size_t)w * ring_smem_per_warp * m <= 200 * 1024 && w * 32 * m <= 2048) { warps = w; grid = g; qq.balanced = 1; break; }
        }
    }
    const size_t smem = (size_t)warps * ring_smem_per_warp;
    SVA_CUDA_OK(ctx, cudaFuncSetAttribute(k_sgm_acc<NR, PF, FULL, STORE, LPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qq.march_warps = warps;
    qq.pace_arrive = nullptr;
    int threads = warps * 32;
    bool n_vert = false;
    for (int i = 0; i < q.ndirs; i++) n_vert = n_vert || q.dys[i] != 0;
    const bool pace = ctx->tune_sgm_pace < 0 ? (size_t)q.W * q.D * 4 >= 768 * 1024 : ctx->tune_sgm_pace != 0;
    if (qq.balanced && qq.cta_sync && pace && q.ndirs > 1 && n_vert && q.W >= grid && warps < 32) {
        // global pacing needs every CTA resident (the grid is one balanced wave by construction; check the occupancy anyway)
        int per_sm = 0;
        SVA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sgm_acc<NR, PF, FULL, STORE, LPL>, threads + 32, smem));
        if (getenv("SVA_DEBUG")) fprintf(stderr, "[sva] sgm_acc NR=%d ndirs=%d grid=%d threads=%d smem=%zu per_sm=%d\n", NR, q.ndirs, grid, threads + 32, smem, per_sm);
        if ((long long)per_sm * ctx->sm_count >= grid) {
            const int rounds = (q.H - PF) / (PF + 1);
            if (rounds > ctx->tune_sgm_pace_window) {
                SVA_TRY(ctx->reserve(ctx->pace_buf, ((size_t)rounds + 16) * sizeof(unsigned int)));
                SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->pace_buf.p, 0, ((size_t)rounds + 16) * sizeof(unsigned int), ctx->stream));
                qq.pace_min = ctx->pace_buf.as<unsigned int>();
                qq.pace_arrive = qq.pace_min + 16;
                qq.pace_rounds = rounds;
                qq.pace_window = ctx->tune_sgm_pace_window;
                threads += 32;  // the helper warp
            }
        }
    }
    // measured: the event-split diagonal march gains 6.6 % on unpaced launches (c1: 0.279 -> 0.261 ms) and nothing on paced ones (c2 / c4
    // move at the pace of the grid-wide minimum), where the generic march is kept
    if (qq.pace_arrive && ctx->tune_sgm_diag_split < 2) qq.diag_split = 0;
    LaunchScope ls(ctx, name);
    k_sgm_acc<NR, PF, FULL, STORE, LPL><<<grid, threads, smem, ctx->stream>>>(qq);
    return SVA_OK;
}

template <int NR, int PF>
static int launch_pass(sva_ctx* ctx, const SgmParams& q, int mode) {
    int nlines = 0;
    for (int i = 0; i < q.ndirs; i++) nlines = std::max(nlines, q.dys[i] == 0 ? q.H : q.W);
    const size_t ring_smem = (size_t)SGM_WARPS * (PF + 1) * 32 * 2 * NR * 2;
    if (mode == SGM_MODE_FINAL) {
        size_t smem = (size_t)SGM_FINAL_WARPS * 8 * q.W + (size_t)SGM_FINAL_WARPS * (PF + 1) * 2 * 32 * 2 * NR * 2;
        if (smem > 200 * 1024) return ctx->fail(SVA_ERR_BAD_ARG, "image too wide for the fused final pass");
        SVA_CUDA_OK(ctx, cudaFuncSetAttribute(k_sgm_pass<NR, SGM_MODE_FINAL, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchScope ls(ctx, q.wta_only ? "k_wta_march" : "k_sgm_final");
        k_sgm_pass<NR, SGM_MODE_FINAL, PF><<<div_up(nlines, SGM_FINAL_WARPS), SGM_FINAL_WARPS * 32, smem, ctx->stream>>>(q);
    } else if (ctx->tune_sgm_lean) {
        const bool full = q.lanes == 32;
        if (mode == SGM_MODE_STORE) {
            const char* nm = q.dys[0] == 0 ? "k_sgm_store_h" : (q.dxs[0] == 0 ? "k_sgm_store_v" : "k_sgm_store_d");
            SVA_TRY((full ? launch_acc<NR, PF, true, true, 32>(ctx, q, nlines, ring_smem / SGM_WARPS, nm) : launch_acc<NR, PF, false, true, 32>(ctx, q, nlines, ring_smem / SGM_WARPS, nm)));
        } else {
            const char* nm = q.ndirs > 1 ? "k_sgm_red_multi" : (q.dys[0] == 0 ? "k_sgm_red_h" : (q.dxs[0] == 0 ? "k_sgm_red_v" : "k_sgm_red_d"));
            SVA_TRY((full ? launch_acc<NR, PF, true, false, 32>(ctx, q, nlines, ring_smem / SGM_WARPS, nm) : launch_acc<NR, PF, false, false, 32>(ctx, q, nlines, ring_smem / SGM_WARPS, nm)));
        }
    } else if (mode == SGM_MODE_STORE) {
        SVA_CUDA_OK(ctx, cudaFuncSetAttribute(k_sgm_pass<NR, SGM_MODE_STORE, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        LaunchScope ls(ctx, q.dys[0] == 0 ? "k_sgm_store_h" : (q.dxs[0] == 0 ? "k_sgm_store_v" : "k_sgm_store_d"));
        k_sgm_pass<NR, SGM_MODE_STORE, PF><<<div_up(nlines, SGM_WARPS), SGM_WARPS * 32, ring_smem, ctx->stream>>>(q);
    } else {
        SVA_CUDA_OK(ctx, cudaFuncSetAttribute(k_sgm_pass<NR, SGM_MODE_RED, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        LaunchScope ls(ctx, q.ndirs > 1 ? "k_sgm_red_multi" : (q.dys[0] == 0 ? "k_sgm_red_h" : (q.dxs[0] == 0 ? "k_sgm_red_v" : "k_sgm_red_d")));
        k_sgm_pass<NR, SGM_MODE_RED, PF><<<div_up(nlines, SGM_WARPS) * q.ndirs, SGM_WARPS * 32, ring_smem, ctx->stream>>>(q);
    }
    SVA_CUDA_OK(ctx, cudaGetLastError());
    return SVA_OK;
}

// several path lines per warp (LPL = 16 or 8 lanes per line, every lane active): accumulate passes only
template <int NR, int LPL>
static int launch_multi(sva_ctx* ctx, const SgmParams& q, int mode) {
    constexpr int PF = 8;
    int nlines = 0;
    for (int i = 0; i < q.ndirs; i++) nlines = std::max(nlines, q.dys[i] == 0 ? q.H : q.W);
    const size_t ring_per_warp = (size_t)(PF + 1) * 32 * 2 * NR * 2;
    const char* nm = mode == SGM_MODE_STORE ? (q.dys[0] == 0 ? "k_sgm_store_h" : (q.dxs[0] == 0 ? "k_sgm_store_v" : "k_sgm_store_d"))
                                            : (q.ndirs > 1 ? "k_sgm_red_multi" : (q.dys[0] == 0 ? "k_sgm_red_h" : (q.dxs[0] == 0 ? "k_sgm_red_v" : "k_sgm_red_d")));
    static const int exp = getenv("SVA_SGM_EXP") ? atoi(getenv("SVA_SGM_EXP")) : 0;  // timing experiments (results are wrong): 1 = no S updates, 2 = plain stores
    if (exp == 1) { SgmParams qq = q; qq.exp_no_out = 1; return launch_acc<NR, PF, true, false, LPL>(ctx, qq, nlines, ring_per_warp, nm); }
    if (mode == SGM_MODE_STORE || exp == 2) SVA_TRY((launch_acc<NR, PF, true, true, LPL>(ctx, q, nlines, ring_per_warp, nm)));
    else SVA_TRY((launch_acc<NR, PF, true, false, LPL>(ctx, q, nlines, ring_per_warp, nm)));
    SVA_CUDA_OK(ctx, cudaGetLastError());
    return SVA_OK;
}

// lanes per line for the accumulate passes: 16 / 8 when D splits evenly into 2, 4, 8, 12 or 16 cells per lane, else 32
static int sgm_multi_lpl(const sva_ctx* ctx, int D) {
    const int want = ctx->tune_sgm_lpl;
    if (!ctx->tune_sgm_lean || (want != 16 && want != 8)) return 32;
    for (int lpl = want; lpl <= 16; lpl *= 2) {
        if (D % (2 * lpl)) continue;
        const int nr = D / (2 * lpl);
        if (nr == 1 || nr == 2 || nr == 4 || nr == 8 || (nr == 6 && lpl == 16)) return lpl;
    }
    return 32;
}

static int launch_pass_nr(sva_ctx* ctx, const SgmParams& q, int nr, int mode) {
    const int pf = ctx->tune_sgm_pf;
    const int lpl = mode == SGM_MODE_FINAL ? 32 : sgm_multi_lpl(ctx, q.D);
    if (lpl != 32) {
        SgmParams qq = q;
        qq.lanes = lpl;
        const int nrm = q.D / (2 * lpl);
        if (lpl == 16) {
            switch (nrm) {
                case 1: return launch_multi<1, 16>(ctx, qq, mode);
                case 2: return launch_multi<2, 16>(ctx, qq, mode);
                case 4: return launch_multi<4, 16>(ctx, qq, mode);
                case 6: return launch_multi<6, 16>(ctx, qq, mode);
                default: return launch_multi<8, 16>(ctx, qq, mode);
            }
        }
        switch (nrm) {
            case 1: return launch_multi<1, 8>(ctx, qq, mode);
            case 2: return launch_multi<2, 8>(ctx, qq, mode);
            case 4: return launch_multi<4, 8>(ctx, qq, mode);
            default: return launch_multi<8, 8>(ctx, qq, mode);
        }
    }
    switch (nr) {
        case 1: return pf >= 16 ? launch_pass<1, 16>(ctx, q, mode) : launch_pass<1, 8>(ctx, q, mode);
        case 2: return pf >= 16 ? launch_pass<2, 16>(ctx, q, mode) : launch_pass<2, 8>(ctx, q, mode);
        default: return pf >= 16 ? launch_pass<4, 16>(ctx, q, mode) : launch_pass<4, 8>(ctx, q, mode);
    }
}

// smallest n in {2,4,8} disparities per lane with D % n == 0 and D / n <= 32
int sva_sgm_regs_per_lane(int D) {
    for (int n = 2; n <= 8; n *= 2)
        if (D % n == 0 && D / n <= 32) return n / 2;
    return 0;
}

// Pass order (oracle direction indices in ORC_DIRS order: 0 v+, 1 v-, 2 h+, 3 h-, 4..7 diagonals):
//   8 paths: 0 (store), 1, 4, 5, 6, 7, 2 (RED), 3 (final)     4 paths: 0 (store), 1, 2 (RED), 3 (final)
static const int DIRS[8][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};

// Selected path directions on an EXTERNAL cost volume, optionally slice-major (multi-GPU: every rank gathers all disparity slices and
// aggregates its share of the directions; the partial sums are then reduce-scattered by rows).  The first direction stores, the others
// RED, so ctx->S ends up as the sum of exactly these directions.  bit i = direction i of {v+, v-, h+, h-, d++, d-+, d+-, d--}.
int sva_run_sgm_dirs(sva_ctx* ctx, const uint16_t* Cext, int c_ds, uint32_t dir_mask, int rows_alloc) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    const int nr = sva_sgm_regs_per_lane(D);
    if (nr == 0) return ctx->fail(SVA_ERR_BAD_ARG, "unsupported num_disp");
    if (c_ds > 0 && (D % c_ds || c_ds % (2 * nr))) return ctx->fail(SVA_ERR_BAD_ARG, "slice size must divide num_disp and hold whole lanes");
    if (!(dir_mask & 0xFFu)) return ctx->fail(SVA_ERR_BAD_ARG, "empty direction mask");
    if (rows_alloc < H) rows_alloc = H;
    const size_t cells = (size_t)W * rows_alloc * D;
    SVA_TRY(ctx->reserve(ctx->S, cells * sizeof(uint16_t) + 64));
    if (rows_alloc > H)  // padding rows (the reduce-scatter wants equal row blocks) must read as zero
        SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->S.as<uint16_t>() + (size_t)W * H * D, 0, (size_t)W * (rows_alloc - H) * D * sizeof(uint16_t), ctx->stream));
    SgmParams q{};
    q.C = Cext; q.S = ctx->S.as<uint16_t>(); q.W = W; q.H = H; q.D = D; q.c_ds = c_ds;
    q.p1p1 = (uint32_t)p.p1 * 0x10001u; q.p2p2 = (uint32_t)p.p2 * 0x10001u;
    q.lanes = D / (2 * nr);
    static const int DIRS8[8][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};
    const int lean = ctx->tune_sgm_lean, lpl = ctx->tune_sgm_lpl;
    ctx->tune_sgm_lean = 1; ctx->tune_sgm_lpl = 32;  // the slice-major cursor lives in the lean one-line-per-warp march
    bool first = true;
    int rc = SVA_OK;
    for (int i = 0; i < 8 && rc == SVA_OK; i++) {
        if (!(dir_mask & (1u << i))) continue;
        q.ndirs = 1; q.dxs[0] = DIRS8[i][0]; q.dys[0] = DIRS8[i][1];
        rc = launch_pass_nr(ctx, q, nr, first ? SGM_MODE_STORE : SGM_MODE_RED);
        first = false;
    }
    ctx->tune_sgm_lean = lean; ctx->tune_sgm_lpl = lpl;
    SVA_TRY(rc);
    ctx->have_sgm = true;
    return SVA_OK;
}

int sva_run_sgm(sva_ctx* ctx) {
    const sva_params& p = ctx->prm;
    const int W = p.width, H = p.height, D = p.num_disp;
    const int nr = sva_sgm_regs_per_lane(D);
    if (nr == 0) return ctx->fail(SVA_ERR_BAD_ARG, "num_disp must be a multiple of 2 with D/8 <= 32 (8..256)");
    size_t cells = (size_t)W * H * D;
    SVA_TRY(ctx->reserve(ctx->S, cells * sizeof(uint16_t) + 64));
    SVA_TRY(ctx->reserve(ctx->disp, (size_t)W * H * sizeof(uint16_t)));
    SVA_TRY(ctx->reserve(ctx->subpix, (size_t)W * H * sizeof(float)));
    SgmParams q{};
    q.C = ctx->C.as<uint16_t>(); q.S = ctx->S.as<uint16_t>(); q.W = W; q.H = H; q.D = D;
    q.p1p1 = (uint32_t)p.p1 * 0x10001u; q.p2p2 = (uint32_t)p.p2 * 0x10001u;
    q.lanes = D / (2 * nr);
    q.dmin = p.min_disp; q.k = p.win_half; q.lr_gx = p.lr_gx; q.lr_max_diff = p.lr_max_diff; q.subpixel = p.subpixel;
    q.store_full = ctx->debug_store_full_s ? 1 : 0;
    q.mask = ctx->has_mask ? ctx->mask.as<uint8_t>() : nullptr;
    q.disp = ctx->disp.as<uint16_t>(); q.sub = ctx->subpix.as<float>();
    for (int i = 0; i < p.n_pairs; i++) {
        int gx = p.pair_gx[i], gy = p.pair_gy[i];
        if (gx > 0) q.gxp = gx > q.gxp ? gx : q.gxp;
        if (gx < 0) q.gxn = -gx > q.gxn ? -gx : q.gxn;
        if (gy > 0) q.gyp = gy > q.gyp ? gy : q.gyp;
        if (gy < 0) q.gyn = -gy > q.gyn ? -gy : q.gyn;
    }
    if (ctx->sgm_dir_mask_override) {  // test hook: accumulate exactly these directions, no final pass
        bool first = true;
        for (int i = 0; i < 8; i++) {
            if (!(ctx->sgm_dir_mask_override & (1u << i))) continue;
            q.ndirs = 1; q.dxs[0] = DIRS[i][0]; q.dys[0] = DIRS[i][1];
            SVA_TRY(launch_pass_nr(ctx, q, nr, first ? SGM_MODE_STORE : SGM_MODE_RED));
            first = false;
        }
        ctx->have_sgm = true;
        return SVA_OK;
    }
    static const int order8[8] = {0, 1, 4, 5, 6, 7, 2, 3}, order4[4] = {0, 1, 2, 3};
    const int n = p.n_paths;
    const int* order = n == 8 ? order8 : order4;
    if (ctx->tune_sgm_fused_final) {
        // variant A: first path stores, middle paths RED (one concurrent launch), last path fused with K3 in one march
        if (n > 0) {
            q.ndirs = 1; q.dxs[0] = DIRS[order[0]][0]; q.dys[0] = DIRS[order[0]][1];
            SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_STORE));
        }
        if (ctx->tune_sgm_concurrent && n > 2) {
            q.ndirs = n - 2;
            for (int i = 1; i + 1 < n; i++) { q.dxs[i - 1] = DIRS[order[i]][0]; q.dys[i - 1] = DIRS[order[i]][1]; }
            SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_RED));
        } else {
            for (int i = 1; i + 1 < n; i++) {
                q.ndirs = 1; q.dxs[0] = DIRS[order[i]][0]; q.dys[0] = DIRS[order[i]][1];
                SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_RED));
            }
        }
        q.ndirs = 1; q.dxs[0] = -1; q.dys[0] = 0;
        q.no_agg = n == 0 ? 1 : 0;
        SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_FINAL));
    } else {
        // variant B (default): S = 0, ALL paths accumulate with REDs in ONE launch (directions interleaved over CTAs: more
        // memory-level parallelism, and paths sweeping the same rows share C and S lines in L2), then a recurrence-free
        // horizontal march does K3 (WTA + LR + sub-pixel) on S_total
        if (n == 0) {  // no aggregation: K3 straight on the cost volume
            SVA_TRY(sva_run_wta(ctx, ctx->C.as<uint16_t>()));
            ctx->have_sgm = false; ctx->have_disp = true;
            return SVA_OK;
        }
        // (zero-filling S from the box filter's store loop instead was measured: +0.06 ms there vs 0.05 ms for this memset)
        if (ctx->s_prezeroed) SVA_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_zero, 0));  // zeroed next to K1a / K1b (sva_api.cu)
        else SVA_CUDA_OK(ctx, cudaMemsetAsync(ctx->S.p, 0, cells * sizeof(uint16_t), ctx->stream));
        ctx->s_prezeroed = false;
        if (ctx->tune_sgm_split && ctx->tune_sgm_lean && n == 8) {
            // variant C: two launches, each = the three directions sweeping the rows one way + one horizontal direction.  With a
            // balanced grid (same CTAs on every SM) the same-sweep directions advance in step without any explicit pacing, so the
            // C and S lines they share are still in L2 when the next direction touches them.
            if (ctx->tune_sgm_split == 3) {  // variant E: all six row-sweeping directions in one launch (down and up sweeps cross mid-image), then the horizontals
                static const int rows6[6] = {0, 4, 5, 1, 6, 7}, hor2[2] = {2, 3};
                q.ndirs = 6;
                for (int i = 0; i < 6; i++) { q.dxs[i] = DIRS[rows6[i]][0]; q.dys[i] = DIRS[rows6[i]][1]; }
                SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_RED));
                q.ndirs = 2;
                for (int i = 0; i < 2; i++) { q.dxs[i] = DIRS[hor2[i]][0]; q.dys[i] = DIRS[hor2[i]][1]; }
                SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_RED));
                SVA_TRY(sva_run_wta(ctx, ctx->S.as<uint16_t>()));
                ctx->have_sgm = true; ctx->have_disp = true;
                return SVA_OK;
            }
            if (ctx->tune_sgm_split == 2) {
                // variant D: three launches — the three directions sweeping the rows downwards, the three sweeping upwards, and the
                // two horizontal ones.  A row-sweeping launch holds only lines that advance one row per step, so (paced) all of them
                // touch a row's C and S lines while these are in L2: DRAM sees C once and S once per launch instead of once per
                // direction.  The horizontal lines (W steps each, every row in flight at once) cannot share and get their own launch.
                static const int grp[3][3] = {{0, 4, 5}, {1, 6, 7}, {2, 3, -1}};
                // The horizontal launch is DRAM-bound and the row-sweeping launches are bound by the integer pipe, so (SVA_SGM_OVERLAP,
                // default on) the horizontal one runs on a second stream next to them: REDs commute, S was zeroed before the fork.
                cudaStream_t main_stream = ctx->stream;
                const bool overlap = ctx->tune_sgm_overlap != 0;
                if (overlap) {
                    if (!ctx->aux_stream) {
                        SVA_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
                        SVA_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                        SVA_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
                    }
                    SVA_CUDA_OK(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
                    SVA_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
                }
                // overlap 1: horizontal launch on the second stream, down / up on the main one; overlap 2: the up-sweeping launch on
                // the second stream next to the down-sweeping one, horizontal launch after the join
                const int order1[3] = {2, 0, 1}, order0[3] = {0, 1, 2}, order2[3] = {1, 0, 2};
                const int* order = ctx->tune_sgm_overlap == 1 ? order1 : (ctx->tune_sgm_overlap == 2 ? order2 : order0);
                for (int gi = 0; gi < 3; gi++) {
                    const int g = order[gi];
                    q.ndirs = g == 2 ? 2 : 3;
                    for (int i = 0; i < q.ndirs; i++) { q.dxs[i] = DIRS[grp[g][i]][0]; q.dys[i] = DIRS[grp[g][i]][1]; }
                    const bool on_aux = overlap && gi == 0;
                    if (on_aux) ctx->stream = ctx->aux_stream;
                    if (ctx->tune_sgm_overlap == 2 && gi == 2) {  // join before the horizontal launch
                        SVA_CUDA_OK(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
                        SVA_CUDA_OK(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_join, 0));
                    }
                    const int rc = launch_pass_nr(ctx, q, nr, SGM_MODE_RED);
                    ctx->stream = main_stream;
                    SVA_TRY(rc);
                }
                if (ctx->tune_sgm_overlap == 1) {
                    SVA_CUDA_OK(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
                    SVA_CUDA_OK(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_join, 0));
                }
                SVA_TRY(sva_run_wta(ctx, ctx->S.as<uint16_t>()));
                ctx->have_sgm = true; ctx->have_disp = true;
                return SVA_OK;
            }
            static const int down8[4] = {0, 4, 5, 2}, up8[4] = {1, 6, 7, 3};
            for (int half = 0; half < 2; half++) {
                static const int exp_same[4] = {0, 0, 0, 2};  // SVA_SGM_EXPERIMENT=1: L2-sharing upper bound (results are wrong)
                const int* dd = getenv("SVA_SGM_EXPERIMENT") ? exp_same : (half ? up8 : down8);
                q.ndirs = 4;
                for (int i = 0; i < 4; i++) { q.dxs[i] = DIRS[dd[i]][0]; q.dys[i] = DIRS[dd[i]][1]; }
                SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_RED));
            }
            SVA_TRY(sva_run_wta(ctx, ctx->S.as<uint16_t>()));
            ctx->have_sgm = true; ctx->have_disp = true;
            return SVA_OK;
        }
        q.ndirs = n;
        // interleave so that neighbouring CTAs sweep the same rows: down, down-diagonals, up, up-diagonals, horizontals
        static const int conc8[8] = {0, 4, 5, 1, 6, 7, 2, 3}, conc4[4] = {0, 1, 2, 3};
        const int* co = n == 8 ? conc8 : conc4;
        for (int i = 0; i < n; i++) { q.dxs[i] = DIRS[co[i]][0]; q.dys[i] = DIRS[co[i]][1]; }
        SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_RED));
        if (ctx->tune_wta_march) {  // K3 as a recurrence-free horizontal march (kept for comparison)
            q.ndirs = 1; q.dxs[0] = -1; q.dys[0] = 0;
            q.wta_only = 1;
            SVA_TRY(launch_pass_nr(ctx, q, nr, SGM_MODE_FINAL));
        } else {
            SVA_TRY(sva_run_wta(ctx, ctx->S.as<uint16_t>()));
        }
    }
    ctx->have_sgm = true; ctx->have_disp = true;
    return SVA_OK;
}
