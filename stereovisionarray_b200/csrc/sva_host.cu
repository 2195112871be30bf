// sva_host.cu — host-side 1:1 shims of the reference's scalar helpers (no GPU work, no oracle code).
// They exist so a caller of include/functions.h / include/Camera.h finds every function behind the C ABI; the hot path
// itself never calls them per pixel (SURVEY §8b "caveat": the per-pixel loop of main is replaced by one batched call).
#include <cmath>
#include <cstdlib>

#include "sva_common.cuh"

extern "C" {

// Camera::project — reference src/Camera.cpp:15-22.  f64, divide-then-divide, truncation toward zero; this TU is built with
// -fmad=false / host -ffp-contract=off so no FMA contraction can change the truncation.
int sva_camera_project(const sva_camera* cam, const double pos3d[3], int32_t out_px[2]) {
    if (!cam || !pos3d || !out_px) return SVA_ERR_BAD_ARG;
    const double scale = cam->f / (pos3d[2] - cam->pos[2]) / cam->pixel_size;
    out_px[0] = (int32_t)((pos3d[0] - cam->pos[0]) * scale);
    out_px[1] = (int32_t)((pos3d[1] - cam->pos[1]) * scale);
    return SVA_OK;
}

// Camera::inv_project — reference src/Camera.cpp:25-33.
int sva_camera_inv_project(const sva_camera* cam, const int32_t px[2], double out_ray[3]) {
    if (!cam || !px || !out_ray) return SVA_ERR_BAD_ARG;
    const double rx = px[0] * cam->pixel_size, ry = px[1] * cam->pixel_size, rz = cam->f;
    const double len = std::sqrt(rx * rx + ry * ry + rz * rz);
    out_ray[0] = rx / len; out_ray[1] = ry / len; out_ray[2] = rz / len;
    return SVA_OK;
}

// bresenham — reference src/functions.cpp:253-321.  The first argument plays the definition's "point2" (:299); output is
// ordered by increasing major-axis coordinate whatever the argument order.
int sva_bresenham(int32_t ax, int32_t ay, int32_t bx, int32_t by, int32_t* out_xy, int32_t cap) {
    if (!out_xy && cap > 0) return SVA_ERR_BAD_ARG;
    const bool x_major = std::abs(ay - by) < std::abs(ax - bx);
    // start = the endpoint with the smaller major coordinate, ties resolved like the reference's `point1 > point2` tests
    bool start_a = x_major ? (bx > ax) : (by > ay);
    int sx = start_a ? ax : bx, sy = start_a ? ay : by, ex = start_a ? bx : ax, ey = start_a ? by : ay;
    int major0 = x_major ? sx : sy, major1 = x_major ? ex : ey, minor = x_major ? sy : sx;
    int dmaj = major1 - major0, dmin = (x_major ? ey : ex) - minor, inc = 1;
    if (dmin < 0) { inc = -1; dmin = -dmin; }
    int e = 2 * dmin - dmaj, n = 0;
    for (int m = major0; m <= major1; m++, n++) {
        if (n < cap) { out_xy[2 * n] = x_major ? m : minor; out_xy[2 * n + 1] = x_major ? minor : m; }
        if (e > 0) { minor += inc; e -= 2 * dmaj; }
        e += 2 * dmin;
    }
    return n;
}

// getCameraPairs — reference src/functions.cpp:148-213 (both overloads; quirks at :202 and :205 preserved).
int sva_get_camera_pairs(int32_t n_cameras, int32_t pair_type, int32_t camera_num, int32_t* out_pairs, int32_t cap) {
    int n = 0;
    auto push = [&](int a, int b) { if (n < cap && out_pairs) { out_pairs[2 * n] = a; out_pairs[2 * n + 1] = b; } n++; };
    if (camera_num >= 0) {
        if (pair_type == SVA_CROSS) {
            if (camera_num - 5 > 0) push(camera_num, camera_num - 5);
            if (camera_num + 5 < 25) push(camera_num, +5);
            if (camera_num % 5 > 0) push(camera_num, camera_num - 1);
            if (camera_num % 5 < 4) push(camera_num, camera_num + 1);
        }
        return n;
    }
    static const int small8[8] = {6, 7, 8, 11, 13, 16, 17, 18};
    switch (pair_type) {
        case SVA_TO_CENTER: for (int i = 0; i < n_cameras; i++) if (i != 12) push(12, i); break;
        case SVA_TO_CENTER_SMALL: for (int i : small8) push(12, i); break;
        case SVA_MID_LEFT: push(12, 11); break;
        case SVA_MID_TOP: push(12, 7); break;
        case SVA_LINE_HORIZONTAL: for (int i = 10; i < 15; i++) if (i != 12) push(12, i); break;
        case SVA_LINE_VERTICAL: for (int i = 2; i < 25; i += 5) if (i != 12) push(12, i); break;
        case SVA_CROSS: push(12, 11); push(12, 13); push(12, 7); push(12, 17); break;
        case SVA_JUMP_CROSS: push(12, 10); push(12, 14); push(12, 2); push(12, 24); break;
        default: break;
    }
    return n;
}

// Generalisation to any rows x cols grid: TO_CENTER = every other camera, TO_CENTER_SMALL = 8-neighbourhood,
// CROSS = 4-neighbourhood, MID_LEFT / MID_TOP = the left / upper neighbour.  Also returns the grid offsets the volume mode needs.
int sva_grid_pairs(int32_t grid_rows, int32_t grid_cols, int32_t ref_index, int32_t pair_type, int32_t* out_pairs, int32_t* out_gx,
                   int32_t* out_gy, int32_t cap) {
    if (grid_rows < 1 || grid_cols < 1 || ref_index < 0 || ref_index >= grid_rows * grid_cols) return SVA_ERR_BAD_ARG;
    const int rr = ref_index / grid_cols, rc = ref_index % grid_cols;
    int n = 0;
    for (int i = 0; i < grid_rows * grid_cols; i++) {
        if (i == ref_index) continue;
        int gy = i / grid_cols - rr, gx = i % grid_cols - rc;
        bool take = false;
        switch (pair_type) {
            case SVA_TO_CENTER: take = true; break;
            case SVA_TO_CENTER_SMALL: take = std::abs(gx) <= 1 && std::abs(gy) <= 1; break;
            case SVA_CROSS: take = std::abs(gx) + std::abs(gy) == 1; break;
            case SVA_MID_LEFT: take = gx == -1 && gy == 0; break;
            case SVA_MID_TOP: take = gx == 0 && gy == -1; break;
            case SVA_LINE_HORIZONTAL: take = gy == 0; break;
            case SVA_LINE_VERTICAL: take = gx == 0; break;
            default: return SVA_ERR_BAD_ARG;
        }
        if (!take) continue;
        if (n < cap) {
            if (out_pairs) { out_pairs[2 * n] = ref_index; out_pairs[2 * n + 1] = i; }
            if (out_gx) out_gx[n] = gx;
            if (out_gy) out_gy[n] = gy;
        }
        n++;
    }
    return n;
}

/* getGroups — include/functions.h:28, src/functions.cpp:107-116: "CHESS" = the CROSS pairs of cameras 0, 2, ..., 24 (13 reference views of
 * the 5x5 array); any other name yields no groups, as in the reference.  Pairs are flattened, out_sizes[g] = pairs of group g. */
int sva_get_groups(int32_t n_cameras, const char* group_type, int32_t* out_pairs, int32_t cap_pairs, int32_t* out_sizes, int32_t cap_groups) {
    if (!group_type || cap_pairs < 0 || cap_groups < 0 || (cap_pairs > 0 && !out_pairs) || (cap_groups > 0 && !out_sizes)) return SVA_ERR_BAD_ARG;
    int ng = 0, np = 0;
    if (std::string(group_type) == "CHESS") {
        for (int i = 0; i < 25; i += 2) {
            int32_t tmp[128];
            const int n = sva_get_camera_pairs(n_cameras, SVA_CROSS, i, tmp, 64);
            if (n < 0) return n;
            if (ng < cap_groups) out_sizes[ng] = n;
            for (int j = 0; j < n && j < 64; j++) {
                if (np < cap_pairs) { out_pairs[2 * np] = tmp[2 * j]; out_pairs[2 * np + 1] = tmp[2 * j + 1]; }
                np++;
            }
            ng++;
        }
    }
    return ng;
}

}  // extern "C"
