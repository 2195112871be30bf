/*
 * sva_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's multi-camera depth path.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's CPU legs, as the checker / reported CPU
 * baseline.  The product (stereovisionarray_b200/, libsva_b200.so) never links, imports or calls it.
 *
 * Parity pinning:
 *  - "literal" functions (camera model, bresenham, pair tables, SAD, the driver loop nest, warp, refine,
 *    depth) restate the cited reference lines and are PINNED against the reference's own sources compiled
 *    here into oracle/_ref/libsva_ref.so (tests/test_oracle_vs_reference.py) and against fixtures those
 *    sources produced (tests/golden/).
 *  - "volume" functions (summed multi-camera cost volume, SGM, LR check, parabolic sub-pixel) have NO
 *    counterpart in the reference: for them parity is UNPINNED by the reference; this
 *    file IS the frozen spec, cross-checked against cv2 primitives and brute force.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (oracle/build_oracle.py).  -ffp-contract=off matters: the
 * reference was built with MSVC /fp:precise (no FMA contraction) and its int() truncations are
 * contraction-sensitive.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sva_c_api.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * Literal mode
 * ---------------------------------------------------------------------------------------------- */

/* Camera::project — src/Camera.cpp:15-22: mult = f/(Z-Cz)/ps (divide then divide), int() truncates toward 0. */
void orc_camera_project(const sva_camera* c, const double p[3], int32_t out[2]) {
    double mult = c->f / (p[2] - c->pos[2]) / c->pixel_size;
    out[0] = (int)((p[0] - c->pos[0]) * mult);
    out[1] = (int)((p[1] - c->pos[1]) * mult);
}

/* Camera::inv_project — src/Camera.cpp:25-33: (u*ps, v*ps, f) / ||.||, norm = sqrt(x*x + y*y + z*z) left to right. */
void orc_camera_inv_project(const sva_camera* c, const int32_t px[2], double out[3]) {
    double vx = px[0] * c->pixel_size, vy = px[1] * c->pixel_size, vz = c->f;
    double n = sqrt(vx * vx + vy * vy + vz * vz);
    out[0] = vx / n; out[1] = vy / n; out[2] = vz / n;
}

/* plotLineLow / plotLineHigh / bresenham — src/functions.cpp:253-321.  (ax,ay) is the FIRST call argument
 * (named point2 in the definition, :299).  Points come out sorted by increasing major-axis coordinate. */
static int line_major(int m0, int n0, int m1, int n1, int x_major, int32_t* out, int cap) {
    /* walk the major axis m from m0 to m1 (m0 <= m1), stepping the minor axis n by +-1 */
    int dm = m1 - m0, dn = n1 - n0, step = 1, cnt = 0;
    if (dn < 0) { step = -1; dn = -dn; }
    int err = 2 * dn - dm, n = n0;
    for (int m = m0; m <= m1; m++) {
        if (cnt < cap) { out[2 * cnt] = x_major ? m : n; out[2 * cnt + 1] = x_major ? n : m; }
        cnt++;
        if (err > 0) { n += step; err -= 2 * dm; }
        err += 2 * dn;
    }
    return cnt;
}
int orc_bresenham(int32_t ax, int32_t ay, int32_t bx, int32_t by, int32_t* out, int32_t cap) {
    /* definition's names: point2 = a, point1 = b */
    if (abs(ay - by) < abs(ax - bx)) {
        if (bx > ax) return line_major(ax, ay, bx, by, 1, out, cap); /* plotLineLow(point2 -> point1) */
        return line_major(bx, by, ax, ay, 1, out, cap);
    }
    if (by > ay) return line_major(ay, ax, by, bx, 0, out, cap);     /* plotLineHigh(point2 -> point1) */
    return line_major(by, bx, ay, ax, 0, out, cap);
}

/* getCameraPairs — src/functions.cpp:148-213, incl. the {cameraNum, +5} quirk at :205 and the `-5 > 0` test at :202. */
int orc_get_camera_pairs(int32_t n_cameras, int32_t type, int32_t camera_num, int32_t* out, int32_t cap) {
    int n = 0;
#define PUSH(a, b) do { if (n < cap) { out[2 * n] = (a); out[2 * n + 1] = (b); } n++; } while (0)
    if (camera_num >= 0) {
        if (type == SVA_CROSS) {
            if (camera_num - 5 > 0) PUSH(camera_num, camera_num - 5);
            if (camera_num + 5 < 25) PUSH(camera_num, 5);
            if (camera_num % 5 > 0) PUSH(camera_num, camera_num - 1);
            if (camera_num % 5 < 4) PUSH(camera_num, camera_num + 1);
        }
        return n;
    }
    switch (type) {
        case SVA_TO_CENTER: for (int i = 0; i < n_cameras; i++) if (i != 12) PUSH(12, i); break;
        case SVA_TO_CENTER_SMALL: { static const int o[8] = {6, 7, 8, 11, 13, 16, 17, 18}; for (int i = 0; i < 8; i++) PUSH(12, o[i]); } break;
        case SVA_MID_LEFT: PUSH(12, 11); break;
        case SVA_MID_TOP: PUSH(12, 7); break;
        case SVA_LINE_HORIZONTAL: for (int i = 10; i < 15; i++) if (i != 12) PUSH(12, i); break;
        case SVA_LINE_VERTICAL: for (int i = 2; i < 25; i += 5) if (i != 12) PUSH(12, i); break;
        case SVA_CROSS: PUSH(12, 11); PUSH(12, 13); PUSH(12, 7); PUSH(12, 17); break;
        case SVA_JUMP_CROSS: PUSH(12, 10); PUSH(12, 14); PUSH(12, 2); PUSH(12, 24); break;
        default: break; /* ORTHOGONAL, DIAGONAL: declared, never implemented -> empty */
    }
#undef PUSH
    return n;
}

/* getAbsDiff — src/functions.cpp:215-218: exact sum of |a-b| (abs(A-B) folds to absdiff), held in a double. */
double orc_abs_diff(const uint8_t* a, size_t astep, const uint8_t* b, size_t bstep, int w, int h) {
    long s = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) s += abs((int)a[y * astep + x] - (int)b[y * bstep + x]);
    return (double)s;
}

/* The driver loop nest — src/CameraStereoVision.cpp:49-95.  out is zero-filled first (the reference leaves the Mat
 * uninitialised, Appendix A.16).  n_threads > 1 parallelises over rows (rows are independent); the reference is 1 thread. */
int orc_match_literal(const sva_image_u8* images, const sva_camera* cams, int32_t n_images, const int32_t* pairs, int32_t n_pairs,
                      const sva_image_u8* mask, int32_t k, double ray_near, double ray_far, uint8_t* out, int32_t n_threads) {
    if (!images || !cams || !pairs || n_images < 1 || n_pairs < 1 || k < 1) return SVA_ERR_BAD_ARG;
    const int W = images[0].cols, H = images[0].rows;
    const int hx = W / 2, hy = H / 2; /* Point2i halfRes = resolution / 2  (:28) */
    memset(out, 0, (size_t)W * H);
    int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int y = k; y < H - k; y++) {
        int cap = 2 * (W + H) + 8;
        int32_t* pts = (int32_t*)malloc(sizeof(int32_t) * 2 * cap);
        for (int x = k; x < W - k; x++) {
            if (mask && mask->data[(size_t)y * mask->step + x] == 0) continue;
            for (int pi = 0; pi < n_pairs; pi++) {
                int r = pairs[2 * pi], o = pairs[2 * pi + 1];
                if (r < 0 || r >= n_images || o < 0 || o >= n_images) { bad = 1; continue; }
                const sva_image_u8 *ir = &images[r], *io = &images[o];
                int32_t q[2] = {x - hx, y - hy};
                double v[3], p1[3], p2[3];
                orc_camera_inv_project(&cams[r], q, v);                                    /* :60 */
                for (int i = 0; i < 3; i++) { p1[i] = cams[r].pos[i] + v[i] * ray_near; p2[i] = cams[r].pos[i] + v[i] * ray_far; } /* :61-62 */
                int32_t a[2], b[2];
                orc_camera_project(&cams[o], p1, a); a[0] += hx; a[1] += hy;               /* :63 */
                orc_camera_project(&cams[o], p2, b); b[0] += hx; b[1] += hy;               /* :64 */
                if (a[0] < k || a[1] < k || a[0] > W - k || a[1] > H - k) continue;        /* :66-68 */
                if (b[0] < k || b[1] < k || b[0] > W - k || b[1] > H - k) continue;        /* :69-71 */
                int n = orc_bresenham(a[0], a[1], b[0], b[1], pts, cap);                   /* :73 */
                double best = 0; int besti = -1;
                for (int i = 0; i < n; i++) {                                              /* :76-83 */
                    int cx = pts[2 * i], cy = pts[2 * i + 1];
                    double e = orc_abs_diff(io->data + (size_t)(cy - k) * io->step + (cx - k), io->step,
                                            ir->data + (size_t)(y - k) * ir->step + (x - k), ir->step, 2 * k, 2 * k);
                    if (besti < 0 || e < best) { best = e; besti = i; }                    /* first minimum, :85 */
                }
                if (besti < 0) continue;
                int ddx = pts[2 * besti] - x, ddy = pts[2 * besti + 1] - y;
                out[(size_t)y * W + x] = (uint8_t)(int)sqrt((double)ddx * ddx + (double)ddy * ddy); /* :89, narrowing */
            }
        }
        free(pts);
    }
    return bad ? SVA_ERR_BAD_ARG : SVA_OK;
}

/* shiftPerspectiveWithDisparity — src/functions.cpp:55-77. */
int orc_shift_perspective_with_disparity(const sva_camera* in_cam, const sva_camera* out_cam, const sva_image_u8* disp,
                                         const sva_image_u8* img, uint8_t* out) {
    const int W = img->cols, H = img->rows;
    double dx = in_cam->pos[0] - out_cam->pos[0], dy = in_cam->pos[1] - out_cam->pos[1], dz = in_cam->pos[2] - out_cam->pos[2];
    double dist = sqrt(dx * dx + dy * dy + dz * dz);
    double mx = dx / dist, my = dy / dist;                                             /* :61-62 */
    memset(out, 0, (size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double d = disp->data[(size_t)y * disp->step + x];
            if (d == 0) continue;                                                      /* :66 */
            int sx = (int)(d * mx + x), sy = (int)(d * my + y);                        /* :69-70 */
            if (sy >= H || sy < 0 || sx >= W || sx < 0) continue;                      /* :71 */
            out[(size_t)y * W + x] = img->data[(size_t)sy * img->step + sx];
        }
    return SVA_OK;
}

/* improveWithDisparity — src/functions.cpp:11-52.  SVA_ERR_ROI where the reference would throw from Mat::operator()(Rect). */
int orc_improve_with_disparity(const sva_image_u8* disp, const sva_image_u8* center, const sva_image_u8* images, const sva_camera* cams,
                               int32_t n, const sva_image_u8* mask, int32_t window_size, uint8_t* out) {
    const int W = center->cols, H = center->rows, k = (window_size - 1) / 2;           /* :17 */
    uint8_t* shifted = (uint8_t*)malloc((size_t)W * H);
    memset(out, 0, (size_t)W * H);
    int rc = SVA_OK;
    for (int c = 0; c < n && rc == SVA_OK; c++) {
        const sva_camera *c0 = &cams[2 * c], *c1 = &cams[2 * c + 1];
        orc_shift_perspective_with_disparity(c0, c1, disp, &images[c], shifted);       /* :22 */
        int dirx = (c0->pos[0] - c1->pos[0]) > 0.001 ? 1 : 0;                          /* :23-25: bool, negative baseline -> 0 */
        int diry = (c0->pos[1] - c1->pos[1]) > 0.001 ? 1 : 0;
        for (int y = 0; y < H && rc == SVA_OK; y++)
            for (int x = 0; x < W; x++) {
                if (mask && mask->data[(size_t)y * mask->step + x] == 0) continue;     /* :29 */
                if (x - k < 0 || y - k < 0 || x + k > W || y + k > H) { rc = SVA_ERR_ROI; break; }
                long best = 0; int besti = -1;
                for (int p = 0; p <= 10; p++) {                                        /* :32-36 */
                    int nx = x + dirx * (p - 5), ny = y + diry * (p - 5);
                    if (nx - k < 0 || ny - k < 0 || nx + k > W || ny + k > H) { rc = SVA_ERR_ROI; break; }
                    long e = (long)orc_abs_diff(shifted + (size_t)(ny - k) * W + (nx - k), W,
                                                center->data + (size_t)(y - k) * center->step + (x - k), center->step, 2 * k, 2 * k);
                    if (besti < 0 || e < best) { best = e; besti = p; }
                }
                if (rc != SVA_OK) break;
                int v = disp->data[(size_t)y * disp->step + x] + (besti - 5) * (dirx + diry);  /* :38 */
                out[(size_t)y * W + x] = (uint8_t)v;
            }
    }
    free(shifted);
    return rc;
}

/* src/CameraStereoVision.cpp:47,98-100: depth = (baseline*f) / (disp*ps); IEEE inf where disp == 0. */
int orc_disparity_to_depth(const sva_image_u8* disp, double baseline, double f, double pixel_size, double* out) {
    double num = baseline * f;
    for (int y = 0; y < disp->rows; y++)
        for (int x = 0; x < disp->cols; x++) out[(size_t)y * disp->cols + x] = num / (disp->data[(size_t)y * disp->step + x] * pixel_size);
    return SVA_OK;
}

/* ---- consumers of the depth output: pinned against the reference's code by tests/test_oracle_vs_reference.py ---- */

/* shiftPerspective2 — src/functions.cpp:79-104.  x outer, y inner, later sources overwrite earlier ones; unwritten pixels are 0 here
 * (uninitialised in the reference). */
int orc_shift_perspective2(const sva_camera* in_cam, const sva_camera* out_cam, const double* depth, int rows, int cols, double* out) {
    double pmx = (in_cam->pos[0] - out_cam->pos[0]) * in_cam->f / in_cam->pixel_size;   /* :82 */
    double pmy = (in_cam->pos[1] - out_cam->pos[1]) * in_cam->f / in_cam->pixel_size;   /* :83 */
    memset(out, 0, sizeof(double) * (size_t)rows * cols);
    for (int x = 0; x < cols; x++)                                                        /* :86 */
        for (int y = 0; y < rows; y++) {                                                  /* :87 */
            double d = depth[(size_t)y * cols + x];
            if (d < 0.5) continue;                                                        /* :89 */
            int sx = (int)(pmx / d) + x, sy = (int)(pmy / d) + y;                         /* :91-92 */
            if (sy >= rows || sy < 0 || sx >= cols || sx < 0) continue;                   /* :93 */
            out[(size_t)sy * cols + sx] = d;                                              /* :95 */
        }
    return SVA_OK;
}

/* Points3DToDepthMap — src/functions.cpp:118-133.  halfRes = resolution / 2 (integer); later points overwrite earlier ones. */
int orc_points3d_to_depth_map(const double* xyz, long long n, const sva_camera* cam, int width, int height, double* out) {
    int hx = width / 2, hy = height / 2;                                                  /* :122 */
    memset(out, 0, sizeof(double) * (size_t)width * height);
    for (long long i = 0; i < n; i++) {
        int32_t px[2];
        orc_camera_project(cam, xyz + 3 * i, px);                                         /* :125 */
        int x = px[0] + hx, y = px[1] + hy;
        if (x >= 0 && x < width && y >= 0 && y < height) out[(size_t)y * width + x] = xyz[3 * i + 2] - cam->pos[2];   /* :126-128 */
    }
    return SVA_OK;
}

/* DepthMapToPoints3D — src/functions.cpp:135-146.  u (column) outer, v (row) inner; returns the point count, writes at most cap. */
long long orc_depth_map_to_points3d(const double* depth, int rows, int cols, const sva_camera* cam, int width, int height, double* out_xyz, long long cap) {
    int hx = width / 2, hy = height / 2;                                                  /* :137 */
    long long n = 0;
    for (int u = 0; u < cols; u++)                                                        /* :139 */
        for (int v = 0; v < rows; v++) {                                                  /* :140 */
            double d = depth[(size_t)v * cols + u];
            if (d > 0.1) {                                                                /* :142 */
                int32_t px[2] = {u - hx, v - hy};
                double r[3];
                orc_camera_inv_project(cam, px, r);
                if (n < cap) { out_xyz[3 * n] = cam->pos[0] + r[0] * d; out_xyz[3 * n + 1] = cam->pos[1] + r[1] * d; out_xyz[3 * n + 2] = cam->pos[2] + r[2] * d; }  /* :143 */
                n++;
            }
        }
    return n;
}

/* getGroups — src/functions.cpp:107-116: "CHESS" = the CROSS pairs of every second camera 0, 2, ..., 24 */
int orc_get_groups(int32_t n_cameras, const char* group_type, int32_t* out_pairs, int32_t cap_pairs, int32_t* out_sizes, int32_t cap_groups) {
    int ng = 0, np = 0;
    if (strcmp(group_type, "CHESS") == 0)
        for (int i = 0; i < 25; i += 2) {                                                 /* :111 */
            int32_t tmp[128];
            int n = orc_get_camera_pairs(n_cameras, SVA_CROSS, i, tmp, 64);               /* :112 */
            if (ng < cap_groups) out_sizes[ng] = n;
            for (int j = 0; j < n; j++) { if (np < cap_pairs) { out_pairs[2 * np] = tmp[2 * j]; out_pairs[2 * np + 1] = tmp[2 * j + 1]; } np++; }
            ng++;
        }
    return ng;
}

/* ------------------------------------------------------------------------------------------------
 * Volume mode (frozen spec) — no reference counterpart: parity UNPINNED by the reference.
 * ---------------------------------------------------------------------------------------------- */

static int params_ok(const sva_params* p) {
    if (!p || p->width < 1 || p->height < 1 || p->num_disp < 1 || p->win_half < 1 || p->n_pairs < 1 || p->n_pairs > SVA_MAX_PAIRS) return 0;
    if (p->cost_cap < 1 || p->cost_cap > SVA_COST_CAP_MAX || p->cost_shift < 0 || p->cost_shift > 31) return 0;
    if (p->p1 < 0 || p->p2 < p->p1 || p->p2 > 4095) return 0;
    if (p->n_paths != 0 && p->n_paths != 4 && p->n_paths != 8) return 0;
    return 1;
}

/* cell validity: reference window inside the image (loop bounds :49-51) and every pair's window centre in
 * [k, dim-k] (the endpoint test :66-71, which allows dim-k itself). */
int orc_cell_valid(const sva_params* p, int y, int x, int d) {
    const int k = p->win_half, W = p->width, H = p->height, delta = p->min_disp + d;
    if (x < k || x >= W - k || y < k || y >= H - k) return 0;
    for (int i = 0; i < p->n_pairs; i++) {
        int cx = x - p->pair_gx[i] * delta, cy = y - p->pair_gy[i] * delta;
        if (cx < k || cx > W - k || cy < k || cy > H - k) return 0;
    }
    return 1;
}

/* ---- census cost mode (p->reserved[0] == SVA_COST_CENSUS) -----------------------------------------------------------------------------
 * north_star names a "census/SAD matching cost"; the reference has only SAD (src/functions.cpp:215-218), so this mode is PARITY UNPINNED BY
 * THE REFERENCE: the spec below is this repository's own (Zabih & Woodfill 1994 as used with SGM by Hirschmueller 2008), frozen here.
 *   signature T(y,x): 62 bits, one per neighbour (dy in [-3,3], dx in [-4,4], centre excluded, row-major, bit 0 first):
 *                     bit = 1 iff I(y+dy, x+dx) < I(y,x); a neighbour outside the image reads as 0.
 *   per-pixel cost  : A(y,x,d) = sum_k popcount(T_R(y,x) XOR T_k(y - gy_k*delta, x - gx_k*delta)); a source pixel outside the image has
 *                     signature 0.  <= 62 * 32 fits u16.
 * Everything after A (box sum over the 2k x 2k window, shift / cap, validity, SGM, WTA) is the same as for SAD. */
#define ORC_CENSUS_RX 4
#define ORC_CENSUS_RY 3
void orc_census_transform(const sva_image_u8* img, uint64_t* out) {
    const int W = img->cols, H = img->rows;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int c = img->data[(size_t)y * img->step + x];
            uint64_t t = 0;
            int bit = 0;
            for (int dy = -ORC_CENSUS_RY; dy <= ORC_CENSUS_RY; dy++)
                for (int dx = -ORC_CENSUS_RX; dx <= ORC_CENSUS_RX; dx++) {
                    if (dx == 0 && dy == 0) continue;
                    const int yy = y + dy, xx = x + dx;
                    const int v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img->data[(size_t)yy * img->step + xx] : 0;
                    if (v < c) t |= (uint64_t)1 << bit;
                    bit++;
                }
            out[(size_t)y * W + x] = t;
        }
}

static int census_ad_volume(const sva_params* p, const sva_image_u8* ref, const sva_image_u8* others, int32_t pair_begin, int32_t pair_end, uint16_t* A) {
    const int W = p->width, H = p->height, D = p->num_disp;
    const size_t px = (size_t)W * H;
    uint64_t* tr = (uint64_t*)malloc(px * sizeof(uint64_t));
    uint64_t* to = (uint64_t*)malloc(px * sizeof(uint64_t) * (size_t)(pair_end - pair_begin > 0 ? pair_end - pair_begin : 1));
    if (!tr || !to) { free(tr); free(to); return SVA_ERR_NOMEM; }
    orc_census_transform(ref, tr);
    for (int i = pair_begin; i < pair_end; i++) orc_census_transform(&others[i], to + px * (size_t)(i - pair_begin));
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint64_t r = tr[(size_t)y * W + x];
            uint16_t* a = A + ((size_t)y * W + x) * D;
            for (int d = 0; d < D; d++) {
                int delta = p->min_disp + d, s = 0;
                for (int i = pair_begin; i < pair_end; i++) {
                    int sx = x - p->pair_gx[i] * delta, sy = y - p->pair_gy[i] * delta;
                    uint64_t v = 0;
                    if (sx >= 0 && sx < W && sy >= 0 && sy < H) v = to[px * (size_t)(i - pair_begin) + (size_t)sy * W + sx];
                    s += __builtin_popcountll(r ^ v);
                }
                a[d] = (uint16_t)s;
            }
        }
    free(tr); free(to);
    return SVA_OK;
}

/* A(y,x,d) = sum over pairs [pair_begin, pair_end) of |R(y,x) - I_k(y - gy*delta, x - gx*delta)|; a source pixel outside
 * the image reads as 0.  Exact u16 (<= 255 * 32).  (Census mode: the Hamming form above.) */
int orc_ad_volume(const sva_params* p, const sva_image_u8* ref, const sva_image_u8* others, int32_t pair_begin, int32_t pair_end, uint16_t* A) {
    if (!params_ok(p) || pair_begin < 0 || pair_end > p->n_pairs) return SVA_ERR_BAD_ARG;
    if (p->reserved[0] == SVA_COST_CENSUS) return census_ad_volume(p, ref, others, pair_begin, pair_end, A);
    if (p->reserved[0] != SVA_COST_SAD) return SVA_ERR_BAD_ARG;
    const int W = p->width, H = p->height, D = p->num_disp;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int r = ref->data[(size_t)y * ref->step + x];
            uint16_t* a = A + ((size_t)y * W + x) * D;
            for (int d = 0; d < D; d++) {
                int delta = p->min_disp + d, s = 0;
                for (int i = pair_begin; i < pair_end; i++) {
                    int sx = x - p->pair_gx[i] * delta, sy = y - p->pair_gy[i] * delta, v = 0;
                    if (sx >= 0 && sx < W && sy >= 0 && sy < H) v = others[i].data[(size_t)sy * others[i].step + sx];
                    s += abs(r - v);
                }
                a[d] = (uint16_t)s;
            }
        }
    return SVA_OK;
}

/* C_raw(y,x,d) = sum_{j,i in [-k,k)} A(y+j, x+i, d) by vertical then horizontal running sums (exact in integers).
 * out16 (nullable): PACK_U16 = valid ? min(cap, raw >> shift) : cap.   out32 (nullable): RAW_U32 = valid ? raw : 0xFFFFFFFF. */
int orc_box_cost(const sva_params* p, const uint16_t* A, uint16_t* out16, uint32_t* out32) {
    if (!params_ok(p)) return SVA_ERR_BAD_ARG;
    const int W = p->width, H = p->height, D = p->num_disp, k = p->win_half;
    const size_t row = (size_t)W * D;
    uint32_t* V = (uint32_t*)calloc(row, sizeof(uint32_t)); /* vertical sums over rows [y-k, y+k) clipped to the image */
    if (!V) return SVA_ERR_NOMEM;
    /* rows [0-k, 0+k) -> rows 0..k-1 */
    for (int j = 0; j < k && j < H; j++)
        for (size_t i = 0; i < row; i++) V[i] += A[(size_t)j * row + i];
    for (int y = 0; y < H; y++) {
        if (y > 0) {
            int add = y + k - 1, sub = y - k - 1;
            if (add < H) for (size_t i = 0; i < row; i++) V[i] += A[(size_t)add * row + i];
            if (sub >= 0) for (size_t i = 0; i < row; i++) V[i] -= A[(size_t)sub * row + i];
        }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int d = 0; d < D; d++) {
            uint32_t s = 0;
            for (int i = 0; i < k && i < W; i++) s += V[(size_t)i * D + d]; /* columns [0-k, 0+k) */
            for (int x = 0; x < W; x++) {
                if (x > 0) {
                    int add = x + k - 1, sub = x - k - 1;
                    if (add < W) s += V[(size_t)add * D + d];
                    if (sub >= 0) s -= V[(size_t)sub * D + d];
                }
                int ok = orc_cell_valid(p, y, x, d);
                size_t o = ((size_t)y * W + x) * D + d;
                if (out32) out32[o] = ok ? s : SVA_COST_INVALID_U32;
                if (out16) {
                    uint32_t c = s >> p->cost_shift;
                    out16[o] = (uint16_t)(ok ? (c < (uint32_t)p->cost_cap ? c : (uint32_t)p->cost_cap) : (uint32_t)p->cost_cap);
                }
            }
        }
    }
    free(V);
    return SVA_OK;
}

/* Brute-force definition of one raw cell straight from the images (cross-check of orc_ad_volume + orc_box_cost);
 * mirrors getAbsDiff on the two 2k x 2k windows, summed over pairs. */
uint32_t orc_raw_cost_cell(const sva_params* p, const sva_image_u8* ref, const sva_image_u8* others, int y, int x, int d) {
    if (!orc_cell_valid(p, y, x, d)) return SVA_COST_INVALID_U32;
    const int k = p->win_half, delta = p->min_disp + d;
    uint32_t s = 0;
    for (int i = 0; i < p->n_pairs; i++) {
        int cx = x - p->pair_gx[i] * delta, cy = y - p->pair_gy[i] * delta;
        s += (uint32_t)orc_abs_diff(others[i].data + (size_t)(cy - k) * others[i].step + (cx - k), others[i].step,
                                    ref->data + (size_t)(y - k) * ref->step + (x - k), ref->step, 2 * k, 2 * k);
    }
    return s;
}

/* Path directions, in the order the passes run.  4 paths = the first four. */
static const int ORC_DIRS[8][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}}; /* (dx, dy) */

/* One SGM path (Hirschmueller 2008, fixed P1/P2):
 *   L(p,d) = C(p,d) + min(L(q,d), L(q,d-1)+P1, L(q,d+1)+P1, min_k L(q,k)+P2) - min_k L(q,k),   q = p - r
 *   L(p,d) = C(p,d) when q is outside the image; d-1 / d+1 outside [0,D) are +inf.
 * Adds L into S (u16; S <= 8 * (cap + P2) <= 65520). */
static void sgm_cell_row(const sva_params* p, const uint16_t* c, const uint16_t* q, uint16_t* l) {
    const int D = p->num_disp, P1 = p->p1, P2 = p->p2;
    if (!q) { for (int d = 0; d < D; d++) l[d] = c[d]; return; }
    int m = q[0];
    for (int d = 1; d < D; d++) if (q[d] < m) m = q[d];
    for (int d = 0; d < D; d++) {
        int best = q[d];
        if (d > 0 && q[d - 1] + P1 < best) best = q[d - 1] + P1;
        if (d < D - 1 && q[d + 1] + P1 < best) best = q[d + 1] + P1;
        if (m + P2 < best) best = m + P2;
        l[d] = (uint16_t)(c[d] + best - m);
    }
}
static void sgm_path_add(const sva_params* p, const uint16_t* C, int dx, int dy, uint16_t* S) {
    const int W = p->width, H = p->height, D = p->num_disp;
    const size_t row = (size_t)W * D;
    if (dy == 0) { /* horizontal: rows are independent */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int y = 0; y < H; y++) {
            uint16_t* cur = (uint16_t*)malloc(row * sizeof(uint16_t));
            for (int xi = 0; xi < W; xi++) {
                int x = dx > 0 ? xi : W - 1 - xi, px = x - dx;
                sgm_cell_row(p, C + ((size_t)y * W + x) * D, (px < 0 || px >= W) ? NULL : cur + (size_t)px * D, cur + (size_t)x * D);
            }
            uint16_t* s = S + (size_t)y * row;
            for (size_t i = 0; i < row; i++) s[i] = (uint16_t)(s[i] + cur[i]);
            free(cur);
        }
        return;
    }
    uint16_t* prev = (uint16_t*)malloc(row * sizeof(uint16_t));
    uint16_t* cur = (uint16_t*)malloc(row * sizeof(uint16_t));
    for (int yi = 0; yi < H; yi++) {
        int y = dy > 0 ? yi : H - 1 - yi, py = y - dy;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int x = 0; x < W; x++) {
            int px = x - dx;
            int border = (px < 0 || px >= W || py < 0 || py >= H);
            sgm_cell_row(p, C + ((size_t)y * W + x) * D, border ? NULL : prev + (size_t)px * D, cur + (size_t)x * D);
        }
        uint16_t* s = S + (size_t)y * row;
        for (size_t i = 0; i < row; i++) s[i] = (uint16_t)(s[i] + cur[i]);
        uint16_t* t = prev; prev = cur; cur = t;
    }
    free(prev); free(cur);
}
/* The row-sweeping part of sgm_path_add restricted to image rows [y0, y0 + rows): the restatement of the multi-GPU row-block pipeline
 * (DESIGN.md §7).  prev_in = L of the row the sweep visited just before the block ([W][D], indexed by image column; NULL when the block
 * starts the sweep), prev_out receives L of the block's last row in sweep order.  Splitting the row loop of sgm_path_add at block
 * boundaries and carrying `prev` across is all there is to it. */
int orc_sgm_rows(const sva_params* p, const uint16_t* C, int32_t dir_index, int32_t y0, int32_t rows, const uint16_t* prev_in, uint16_t* S,
                 uint16_t* prev_out) {
    if (!params_ok(p) || dir_index < 0 || dir_index > 7 || y0 < 0 || rows < 1 || y0 + rows > p->height) return SVA_ERR_BAD_ARG;
    const int W = p->width, H = p->height, D = p->num_disp, dx = ORC_DIRS[dir_index][0], dy = ORC_DIRS[dir_index][1];
    if (dy == 0) return SVA_ERR_BAD_ARG; /* horizontal paths are row-local: use orc_sgm_single_path on the rows */
    const size_t row = (size_t)W * D;
    uint16_t* prev = (uint16_t*)malloc(row * sizeof(uint16_t));
    uint16_t* cur = (uint16_t*)malloc(row * sizeof(uint16_t));
    if (prev_in) memcpy(prev, prev_in, row * sizeof(uint16_t));
    for (int yi = 0; yi < rows; yi++) {
        int y = dy > 0 ? y0 + yi : y0 + rows - 1 - yi, py = y - dy;
        for (int x = 0; x < W; x++) {
            int px = x - dx;
            int border = (px < 0 || px >= W || py < 0 || py >= H);
            if (!border && !prev_in && yi == 0) { free(prev); free(cur); return SVA_ERR_BAD_ARG; } /* the block continues a sweep: state needed */
            sgm_cell_row(p, C + ((size_t)y * W + x) * D, border ? NULL : prev + (size_t)px * D, cur + (size_t)x * D);
        }
        uint16_t* s = S + (size_t)y * row;
        for (size_t i = 0; i < row; i++) s[i] = (uint16_t)(s[i] + cur[i]);
        uint16_t* t = prev; prev = cur; cur = t;
    }
    if (prev_out) memcpy(prev_out, prev, row * sizeof(uint16_t));
    free(prev); free(cur);
    return SVA_OK;
}
/* one path on its own: L_out = L_r (for unit tests) */
int orc_sgm_single_path(const sva_params* p, const uint16_t* C, int32_t dir_index, uint16_t* L_out) {
    if (!params_ok(p) || dir_index < 0 || dir_index > 7) return SVA_ERR_BAD_ARG;
    memset(L_out, 0, (size_t)p->width * p->height * p->num_disp * sizeof(uint16_t));
    sgm_path_add(p, C, ORC_DIRS[dir_index][0], ORC_DIRS[dir_index][1], L_out);
    return SVA_OK;
}
/* S = sum over the first n_use paths (n_use <= n_paths; n_use = n_paths gives the full aggregation). */
int orc_sgm_aggregate(const sva_params* p, const uint16_t* C, int32_t n_use, uint16_t* S) {
    if (!params_ok(p) || n_use < 0 || n_use > 8) return SVA_ERR_BAD_ARG;
    memset(S, 0, (size_t)p->width * p->height * p->num_disp * sizeof(uint16_t));
    for (int i = 0; i < n_use; i++) sgm_path_add(p, C, ORC_DIRS[i][0], ORC_DIRS[i][1], S);
    return SVA_OK;
}

/* WTA (first minimum) + optional left-right check + optional parabolic sub-pixel on S (or on C when n_paths == 0).
 *  - pixel rejected (disp = SVA_DISP_INVALID, subpix = -1) when: outside [k,W-k) x [k,H-k); mask == 0; the winning cell is
 *    invalid (orc_cell_valid); or the LR check fails.
 *  - LR: D_o(x') = first argmin_d S(y, x' + lr_gx*delta, d) over the d whose source column lies in [0,W);
 *    x' = x - lr_gx*(min_disp + d*); reject when x' is outside [0,W) or |d* - D_o(x')| > lr_max_diff.
 *  - sub-pixel: for 0 < d* < D-1 and den = S(d*-1) - 2 S(d*) + S(d*+1) > 0: d* + (S(d*-1) - S(d*+1)) / (2 den) in f32.
 *  outputs carry min_disp: disp = min_disp + d*. */
int orc_wta(const sva_params* p, const uint16_t* S, const sva_image_u8* mask, uint16_t* disp, float* subpix) {
    if (!params_ok(p)) return SVA_ERR_BAD_ARG;
    const int W = p->width, H = p->height, D = p->num_disp, k = p->win_half;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int y = 0; y < H; y++) {
        int* best_d = (int*)malloc(sizeof(int) * W);
        int* other_d = (int*)malloc(sizeof(int) * W);
        for (int x = 0; x < W; x++) {
            const uint16_t* s = S + ((size_t)y * W + x) * D;
            int b = 0;
            for (int d = 1; d < D; d++) if (s[d] < s[b]) b = d;
            best_d[x] = b;
        }
        if (p->lr_gx != 0) {
            for (int xo = 0; xo < W; xo++) {
                int b = -1; uint16_t bv = 0;
                for (int d = 0; d < D; d++) {
                    int x = xo + p->lr_gx * (p->min_disp + d);
                    if (x < 0 || x >= W) continue;
                    uint16_t v = S[((size_t)y * W + x) * D + d];
                    if (b < 0 || v < bv) { b = d; bv = v; }
                }
                other_d[xo] = b;
            }
        }
        for (int x = 0; x < W; x++) {
            size_t o = (size_t)y * W + x;
            int d = best_d[x], ok = 1;
            if (x < k || x >= W - k || y < k || y >= H - k) ok = 0;
            if (ok && mask && mask->data[(size_t)y * mask->step + x] == 0) ok = 0;
            if (ok && !orc_cell_valid(p, y, x, d)) ok = 0;
            if (ok && p->lr_gx != 0) {
                int xo = x - p->lr_gx * (p->min_disp + d);
                if (xo < 0 || xo >= W || other_d[xo] < 0 || abs(d - other_d[xo]) > p->lr_max_diff) ok = 0;
            }
            if (!ok) { disp[o] = SVA_DISP_INVALID; if (subpix) subpix[o] = SVA_SUBPIX_INVALID; continue; }
            disp[o] = (uint16_t)(p->min_disp + d);
            if (subpix) {
                float f = (float)d;
                if (p->subpixel && d > 0 && d < D - 1) {
                    const uint16_t* s = S + o * D;
                    int den = (int)s[d - 1] - 2 * (int)s[d] + (int)s[d + 1];
                    if (den > 0) f = (float)d + (float)((int)s[d - 1] - (int)s[d + 1]) / (float)(2 * den);
                }
                subpix[o] = (float)p->min_disp + f;
            }
        }
        free(best_d); free(other_d);
    }
    return SVA_OK;
}

/* Whole volume-mode pipeline.  Scratch volumes are caller-provided when non-NULL (tests read them back). */
int orc_depth_from_array(const sva_params* p, const sva_image_u8* ref, const sva_image_u8* others, const sva_image_u8* mask,
                         uint16_t* A_io, uint16_t* C_io, uint16_t* S_io, uint16_t* disp, float* subpix) {
    if (!params_ok(p)) return SVA_ERR_BAD_ARG;
    size_t n = (size_t)p->width * p->height * p->num_disp;
    uint16_t* A = A_io ? A_io : (uint16_t*)malloc(n * 2);
    uint16_t* C = C_io ? C_io : (uint16_t*)malloc(n * 2);
    uint16_t* S = S_io ? S_io : (p->n_paths ? (uint16_t*)malloc(n * 2) : NULL);
    int rc = orc_ad_volume(p, ref, others, 0, p->n_pairs, A);
    if (rc == SVA_OK) rc = orc_box_cost(p, A, C, NULL);
    if (rc == SVA_OK && p->n_paths) rc = orc_sgm_aggregate(p, C, p->n_paths, S);
    if (rc == SVA_OK) rc = orc_wta(p, p->n_paths ? S : C, mask, disp, subpix);
    if (!A_io) free(A);
    if (!C_io) free(C);
    if (!S_io && S) free(S);
    return rc;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
