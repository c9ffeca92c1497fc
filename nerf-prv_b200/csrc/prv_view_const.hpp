// prv_view_const.hpp -- host-side per-view constants of the ray-cast kernels (prvk::ViewConst): snapped origin, inverse
// pose, and the per-view proof that lets the march drop castRay's max-range and key-overflow tests.  Used by prv_device.cu
// (prv_set_views) and, unchanged, by the CPU check of the per-ray code (tests/cpp/kernel_on_host.cpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "../../include/prv.h"
#include "../host/prv_linalg.hpp"
#include "prv_kernels.cuh"
#include "prv_keys.hpp"

namespace prvk {

// what make_view_const needs to know about the resident map and camera
struct ViewSetup {
    double resolution;
    int lo[3], n[3];        // occupancy AABB (keys)
    const uint16_t* keys;   // [n_keys][3] leaf (Morton) order
    size_t n_keys;
    double max_range, max_range_sq;
};

// castRay's d^2 for a key triple, host copy of dist_sq_at (float terms, double accumulation)
inline double host_dist_sq(const float origin[3], double res, const int k[3]) {
    double acc = 0.0;
    for (int j = 0; j < 3; j++) {
        const float e = (float)prv::key_to_coord(k[j], res);
        const float df = e - origin[j];
        acc += (double)(df * df);
    }
    return acc;
}

// per-view constants: snapped origin (main.cpp:112-114), inverse pose (main.cpp:244), fast-path proof
inline void make_view_const(const ViewSetup& ctx, const double* pose_world, const double* init_pos, uint32_t id, ViewConst& vc) {
    const prv::Matrix4d pw = prv::Matrix4d::FromRowMajor(pose_world);
    const prv::Matrix4d inv = pw.inverse();
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            vc.pose[4 * r + c] = pw(r, c);
            vc.posef[4 * r + c] = (float)pw(r, c);
            vc.inv[4 * r + c] = inv(r, c);
        }
    vc.flags = 0;
    vc.view_id = id;
    const double rf = 1.0 / ctx.resolution;
    uint16_t k[3] = {0, 0, 0};
    bool ok = true;
    for (int a = 0; a < 3; a++) ok = prv::coord_to_key_checked(init_pos[a], rf, k[a]) && ok;
    for (int a = 0; a < 3; a++) {
        vc.okey[a] = k[a];
        vc.origin[a] = ok ? (float)prv::key_to_coord(k[a], ctx.resolution) : 0.0f;
    }
    for (int a = 0; a < 3; a++) vc.tnum[a][0] = vc.tnum[a][1] = 0.0;
    for (int a = 0; a < 3; a++) vc.ovox[a] = 0.0f;
    vc.pad_ = 0.0f;
    if (!ok) return;
    for (int r = 0; r < 3; r++) vc.posef[4 * r + 3] = (float)(pw(r, 3) - (double)vc.origin[r]);
    vc.flags |= kViewInMap;
    // castRay re-derives current_key from the float origin; it is the same key (|error| << half a voxel) but recompute literally
    for (int a = 0; a < 3; a++) {
        uint16_t kk;
        if (prv::coord_to_key_checked((double)vc.origin[a], rf, kk)) vc.okey[a] = kk;
    }
    for (int a = 0; a < 3; a++) vc.ovox[a] = (float)(vc.okey[a] - ctx.lo[a]) + 0.5f;
    // castRay's voxelBorder - origin per axis and step sign (OccupancyOcTreeBase::castRay: voxelBorder = keyToCoord(key) +
    // step * resolution * 0.5; tMax = (voxelBorder - (double)origin) / direction), operation by operation
    for (int a = 0; a < 3; a++)
        for (int sgn = 0; sgn < 2; sgn++) {
            double border = prv::key_to_coord(vc.okey[a], ctx.resolution);
            border = border + ((double)(sgn == 0 ? 1 : -1) * ctx.resolution) * 0.5;
            vc.tnum[a][sgn] = border - (double)vc.origin[a];
        }
    // origin voxel occupied?
    bool inside = true;
    for (int a = 0; a < 3; a++) inside = inside && vc.okey[a] >= ctx.lo[a] && vc.okey[a] < ctx.lo[a] + ctx.n[a];
    if (inside && ctx.n[0] > 0) {
        const uint64_t code = prv::morton_code((uint16_t)vc.okey[0], (uint16_t)vc.okey[1], (uint16_t)vc.okey[2]);
        const size_t N = ctx.n_keys;
        size_t lo = 0, hi = N;
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            const uint64_t cm = prv::morton_code(ctx.keys[3 * mid], ctx.keys[3 * mid + 1], ctx.keys[3 * mid + 2]);
            if (cm < code) lo = mid + 1; else hi = mid;
        }
        if (lo < N && prv::morton_code(ctx.keys[3 * lo], ctx.keys[3 * lo + 1], ctx.keys[3 * lo + 2]) == code) vc.flags |= kViewInObject;
    }
    // fast path proof: while a ray can still hit, every key it visits lies in the box spanned by the origin key and the
    // AABB grown by one voxel; the float d^2 terms are monotone in |key - origin key| per axis, so the corner sum bounds
    // every d^2 the reference would test.  Also no key-overflow test can fire inside [1, 65534].
    bool fast = ctx.n[0] > 0;
    for (int a = 0; a < 3 && fast; a++) fast = ctx.lo[a] >= 2 && ctx.lo[a] + ctx.n[a] <= 65533;
    if (fast && ctx.max_range > 0.0) {
        double acc = 0.0;
        for (int a = 0; a < 3; a++) {
            int kl[3] = {vc.okey[0], vc.okey[1], vc.okey[2]}, kh[3] = {vc.okey[0], vc.okey[1], vc.okey[2]};
            kl[a] = ctx.lo[a] - 1;
            kh[a] = ctx.lo[a] + ctx.n[a];
            // per-axis term = d^2 with the other two axes at the origin key (their terms are exactly 0)
            const double tl = host_dist_sq(vc.origin, ctx.resolution, kl);
            const double th = host_dist_sq(vc.origin, ctx.resolution, kh);
            acc += std::max(tl, th);
        }
        // acc >= any double-accumulated d^2 in the box up to 2 roundings; keep a 1e-9 relative guard band
        fast = acc * (1.0 + 1e-9) <= ctx.max_range_sq;
    }
    if (fast) vc.flags |= kViewFastOk;
}

// The region-level cull (region_corner_outside) assumes that every pixel of a 32x32 region maps (in normalised image
// coordinates, the deprojection of Share_Data.hpp:140-196) inside the quad of the region's corners pushed 2 px outwards.
// Exact for pin-hole models; for the Brown-Conrady polynomial it is checked here for every region of the image on a 9x9
// grid of sample pixels: each must lie inside that quad with at least half a pixel to spare.
inline bool region_cull_valid(const prv_intrinsics& intr) {
    auto deproject = [&](float pu, float pv, float& x, float& y) {
        x = (pu - intr.ppx) / intr.fx;
        y = (pv - intr.ppy) / intr.fy;
        if (intr.model == 2) {
            const float r2 = x * x + y * y;
            const float f = 1 + intr.coeffs[0] * r2 + intr.coeffs[1] * r2 * r2 + intr.coeffs[4] * r2 * r2 * r2;
            const float ux = x * f + 2 * intr.coeffs[2] * x * y + intr.coeffs[3] * (r2 + 2 * x * x);
            const float uy = y * f + 2 * intr.coeffs[3] * x * y + intr.coeffs[2] * (r2 + 2 * y * y);
            x = ux;
            y = uy;
        }
    };
    bool ok = true;
    const float spare = 0.5f / std::max(intr.fx, intr.fy);
    for (int ry = 0; ry * 32 < intr.height + 1 && ok; ry++)
        for (int rx = 0; rx * 32 < intr.width + 1 && ok; rx++) {
            float qx[4], qy[4];
            const float x0 = (float)(rx * 32 - 2), x1 = (float)(rx * 32 + 33), y0 = (float)(ry * 32 - 2), y1 = (float)(ry * 32 + 33);
            deproject(x0, y0, qx[0], qy[0]);
            deproject(x1, y0, qx[1], qy[1]);
            deproject(x1, y1, qx[2], qy[2]);
            deproject(x0, y1, qx[3], qy[3]);
            float cx, cy;
            deproject(0.5f * (x0 + x1), 0.5f * (y0 + y1), cx, cy);
            for (int sy = 0; sy <= 8 && ok; sy++)
                for (int sx = 0; sx <= 8 && ok; sx++) {
                    float px, py;
                    deproject((float)(rx * 32) + 31.0f * sx / 8.0f, (float)(ry * 32) + 31.0f * sy / 8.0f, px, py);
                    for (int e = 0; e < 4; e++) {
                        const float ex = qx[(e + 1) & 3] - qx[e], ey = qy[(e + 1) & 3] - qy[e];
                        const float len = std::sqrt(ex * ex + ey * ey);
                        if (!(len > 0)) { ok = false; break; }
                        const float side_p = (ex * (py - qy[e]) - ey * (px - qx[e])) / len;
                        const float side_c = (ex * (cy - qy[e]) - ey * (cx - qx[e])) / len;
                        // the sample must be on the same side of the edge as the region centre, at least `spare` inside
                        if (!(side_c != 0 && side_p * (side_c > 0 ? 1.0f : -1.0f) > spare)) { ok = false; break; }
                    }
                }
        }
    return ok;
}

}  // namespace prvk
