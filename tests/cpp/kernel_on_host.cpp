// kernel_on_host.cpp -- TEST INFRASTRUCTURE (a checker, not a product path and not a CPU fallback).
//
// Compiles the per-ray device code of the ray-cast kernels -- nerf-prv_b200/csrc/prv_kernels.cuh, unchanged -- with g++ and
// runs it one ray at a time on the CPU, so the `-m "not gpu"` suite can hold the kernels' arithmetic and the exactness of
// the three conservative culls against the oracle without a GPU:
//   * region cull (region_corner_outside / region_skip_from_ballot), slab + brick cull (ray_direction_approx + coarse_miss)
//     and coarse brick cull (coarse_miss) exactly as cull_kernel / coarse_kernel chain them;
//   * the exact set-up (ray_direction + ray_init) and the three march variants (march_plain / march_fast / march_axis);
//   * the per-view constants and the "fast path" proof (prv_view_const.hpp, the code prv_set_views runs);
//   * voxel-driven mode: project_point_to_pixel, truncated pixel, per-voxel gather (project_voxels_kernel, precept_points_kernel).
// The lookup tables (occupancy bitmap, shell-padded bitmap, coarse grid, prefix, rank table) are rebuilt here on the host
// from the layout documented in prv_kernels.cuh / DESIGN.md section 3 -- an independent statement of what the
// map_*_kernel write.  What this cannot check: warp-level plumbing (ballots, queues, tickets, atomics), the PTX of
// dda_step (restated below in C++) and the approximate device intrinsics used by the culls only (__fdividef, rsqrtf,
// __frcp_rn are the IEEE operations here; the culls' margins are 3+ orders of magnitude above either error).
// Those stay with the GPU parity tests (tests/test_gpu_parity.py, tests/test_gpu_edge.py).
//
// Build (tests/test_kernel_on_host.py): g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -I/usr/local/cuda/include
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// ---- the CUDA device intrinsics the header uses, as the IEEE operations they denote -----------------------------------
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }  // culls only
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }  // culls only
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline double __longlong_as_double(long long v) {
    double d;
    std::memcpy(&d, &v, 8);
    return d;
}
static inline long long __double_as_longlong(double d) {
    long long v;
    std::memcpy(&v, &d, 8);
    return v;
}
using std::max;
using std::min;

// ---- dda_step: the one function of the header written in PTX; same statement in C++ ------------------------------------
#define PRVK_HOST_CHECK 1
namespace prvk {
struct DdaMasks;
static inline void dda_masks_init(DdaMasks&) {}
static inline uint32_t dda_step(double& t0, double& t1, double& t2, double d0, double d1, double d2, uint32_t inc0, uint32_t inc1, uint32_t inc2, DdaMasks&) {
    const bool c01 = t0 < t1, c02 = t0 < t2, c12 = t1 < t2;
    const bool p0 = c01 && c02, p1 = !c01 && c12, p2 = !(p0 || p1);
    t0 = std::fma(p0 ? 1.0 : 0.0, d0, t0);
    t1 = std::fma(p1 ? 1.0 : 0.0, d1, t1);
    t2 = std::fma(p2 ? 1.0 : 0.0, d2, t2);
    return p0 ? inc0 : (p1 ? inc1 : inc2);
}
}  // namespace prvk

#include "../../nerf-prv_b200/csrc/prv_kernels.cuh"
#include "../../nerf-prv_b200/csrc/prv_view_const.hpp"

using namespace prvk;

namespace {

// the HBM tables of one map, built on the host from the documented layout
struct HostMap {
    DevMap m{};
    std::vector<uint32_t> bitmap, pad, coarse, prefix, leaf_of_raster;
    std::vector<uint16_t> keys;
    std::vector<uint8_t> rgb;
    ViewSetup setup{};
};

bool build_map(HostMap& hm, const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution, double max_range, int brick_cs = kCoarseDefault) {
    if (N == 0) return false;
    hm.keys.assign(keys, keys + 3 * (size_t)N);
    if (rgb) hm.rgb.assign(rgb, rgb + 3 * (size_t)N); else hm.rgb.assign(3 * (size_t)N, 0);
    int lo[3] = {65536, 65536, 65536}, hi[3] = {-1, -1, -1};
    for (uint32_t i = 0; i < N; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], (int)keys[3 * i + a]);
            hi[a] = std::max(hi[a], (int)keys[3 * i + a]);
        }
    DevMap& m = hm.m;
    m.resolution = resolution;
    for (int a = 0; a < 3; a++) {
        m.lo[a] = lo[a];
        m.n[a] = hi[a] - lo[a] + 1;
        m.nc[a] = (m.n[a] + brick_cs - 1) / brick_cs;
        m.nhi[a] = (float)m.n[a] + 1.0f;
        // AABB grown by 2 voxels, metres
        m.bmin[a] = (float)((double)(lo[a] - 2 - prv::kTreeMaxVal) * resolution);
        m.bmax[a] = (float)((double)(lo[a] + m.n[a] + 2 - prv::kTreeMaxVal) * resolution);
    }
    m.wx = (m.n[0] + 31) / 32;
    m.n_occ = N;
    uint32_t w64 = (N + 63) / 64;
    if (w64 == 0) w64 = 1;
    m.words64 = (w64 + 1) & ~1u;
    // dense occupancy [n2][n1][wx]
    hm.bitmap.assign((size_t)m.wx * m.n[1] * m.n[2], 0u);
    // shell-padded copy: rows of 1 << row_log2 bits, 3 planes + 2 rows of slack on both sides
    int row_log2 = 5;
    while ((1 << row_log2) < m.n[0] + 2) row_log2++;
    const size_t pad_rows = (size_t)(m.n[1] + 2) * (m.n[2] + 2);
    const size_t slack_bits = ((size_t)3 * (m.n[1] + 2) + 2) << row_log2;
    hm.pad.assign(((pad_rows << row_log2) + 2 * slack_bits) / 32, 0u);
    m.pad_row_log2 = row_log2;
    m.pad_words = (uint32_t)hm.pad.size();
    m.pad_bit_offset = (uint32_t)slack_bits;
    auto set_pad = [&](int c0, int c1, int c2) {  // padded cell coordinates (0..n+1)
        const size_t L = slack_bits + (((size_t)c2 * (m.n[1] + 2) + c1) << row_log2) + (size_t)c0;
        hm.pad[L >> 5] |= 1u << (L & 31);
    };
    for (int c2 = 0; c2 < m.n[2] + 2; c2++)
        for (int c1 = 0; c1 < m.n[1] + 2; c1++)
            for (int c0 = 0; c0 < m.n[0] + 2; c0++)
                if (c0 == 0 || c0 == m.n[0] + 1 || c1 == 0 || c1 == m.n[1] + 1 || c2 == 0 || c2 == m.n[2] + 1) set_pad(c0, c1, c2);
    hm.coarse.assign(((size_t)m.nc[0] * m.nc[1] * m.nc[2] + 31) / 32 + 1, 0u);
    m.cs = brick_cs;  // brick edge (prv_set_brick_cull)
    m.inv_cs = 1.0f / (float)brick_cs;
    {
        const uint64_t n1p = (uint64_t)m.n[1] + 2;
        const uint64_t magic = (((uint64_t)1 << 32) + n1p - 1) / n1p;
        m.n1p_magic = (pad_rows + 8) * n1p < ((uint64_t)1 << 32) && magic < ((uint64_t)1 << 32) ? (uint32_t)magic : 0u;
    }
    for (uint32_t i = 0; i < N; i++) {
        const int q[3] = {keys[3 * i] - lo[0], keys[3 * i + 1] - lo[1], keys[3 * i + 2] - lo[2]};
        hm.bitmap[((size_t)q[2] * m.n[1] + q[1]) * m.wx + (q[0] >> 5)] |= 1u << (q[0] & 31);
        set_pad(q[0] + 1, q[1] + 1, q[2] + 1);
        // every coarse cell that the voxel grown by one voxel touches
        for (int d2 = -1; d2 <= 1; d2++)
            for (int d1 = -1; d1 <= 1; d1++)
                for (int d0 = -1; d0 <= 1; d0++) {
                    const int p0 = q[0] + d0, p1 = q[1] + d1, p2 = q[2] + d2;
                    if (p0 < 0 || p1 < 0 || p2 < 0 || p0 >= m.n[0] || p1 >= m.n[1] || p2 >= m.n[2]) continue;
                    const uint32_t c = (uint32_t)(((p2 / brick_cs) * m.nc[1] + p1 / brick_cs) * m.nc[0] + p0 / brick_cs);
                    hm.coarse[c >> 5] |= 1u << (c & 31);
                }
    }
    hm.prefix.assign(hm.bitmap.size(), 0u);
    uint32_t run = 0;
    for (size_t w = 0; w < hm.bitmap.size(); w++) {
        hm.prefix[w] = run;
        run += (uint32_t)__builtin_popcount(hm.bitmap[w]);
    }
    if (run != N) return false;  // duplicate keys
    hm.leaf_of_raster.assign(N, 0u);
    for (uint32_t i = 0; i < N; i++) {
        const int q0 = keys[3 * i] - lo[0], q1 = keys[3 * i + 1] - lo[1], q2 = keys[3 * i + 2] - lo[2];
        const size_t w = ((size_t)q2 * m.n[1] + q1) * m.wx + (q0 >> 5);
        hm.leaf_of_raster[hm.prefix[w] + (uint32_t)__builtin_popcount(hm.bitmap[w] & ((1u << (q0 & 31)) - 1u))] = i;
    }
    m.bitmap = hm.bitmap.data();
    m.bitmap_pad = hm.pad.data();
    m.coarse = hm.coarse.data();
    m.prefix = hm.prefix.data();
    m.leaf_of_raster = hm.leaf_of_raster.data();
    m.keys = hm.keys.data();
    m.rgb = hm.rgb.data();
    hm.setup.resolution = resolution;
    for (int a = 0; a < 3; a++) {
        hm.setup.lo[a] = m.lo[a];
        hm.setup.n[a] = m.n[a];
    }
    hm.setup.keys = hm.keys.data();
    hm.setup.n_keys = N;
    hm.setup.max_range = max_range;
    hm.setup.max_range_sq = max_range * max_range;
    return true;
}

DevCam make_cam(const prv_intrinsics& it, double max_range, int force_region_cull) {
    DevCam c{};
    c.W = it.width;
    c.H = it.height;
    c.ppx = it.ppx;
    c.ppy = it.ppy;
    c.fx = it.fx;
    c.fy = it.fy;
    c.model = it.model;
    for (int i = 0; i < 5; i++) c.c[i] = it.coeffs[i];
    c.max_range = max_range;
    c.max_range_sq = max_range * max_range;
    c.inv_fx = 1.0f / it.fx;
    c.inv_fy = 1.0f / it.fy;
    c.region_cull_ok = force_region_cull >= 0 ? force_region_cull : (region_cull_valid(it) ? 1 : 0);
    return c;
}

enum { S_RAYS = 0, S_REGION_CULLED, S_LOOSE_CULLED, S_COARSE_CULLED, S_MARCHED, S_PROBES, S_STEPS, S_HITS, S_FLAGS, S_REGION_OK, S_BOX_ENTRIES, S_BOX_FALLBACKS, S_N };

// one pixel through stages 2..3 of the AXIS pipeline (coarse_kernel, march_kernel)
void axis_pixel(const HostMap& hm, const DevCam& cam, const ViewConst& vc, int px, int py, CastResult& res, uint64_t* st, bool entry = true) {
    res.rank = kNone;
    res.steps = res.probes = 0;
    res.k0 = res.k1 = res.k2 = 0;
    const bool fast = (vc.flags & kViewFastOk) != 0;
    uint32_t cell = kNone;  // what coarse_kernel hands to march_kernel through queue2b
    if (fast) {
        float dx, dy, dz;
        ray_direction_approx(cam, vc, (float)px, (float)py, dx, dy, dz);
        if (coarse_miss(hm.m, vc, dx, dy, dz, cell)) {
            st[S_COARSE_CULLED]++;
            return;
        }
    }
    st[S_MARCHED]++;
    if (!entry) cell = kNone;  // cast_impl passes no queue2b
    if (cell != kNone) st[S_BOX_ENTRIES]++;
    for (;;) {  // march_kernel's loop
        RayState r;
        float dx, dy, dz;
        ray_direction(cam, vc, px, py, dx, dy, dz);
        if (!ray_init(vc, hm.m.resolution, dx, dy, dz, r)) break;
        if (!fast) {
            march_plain(hm.m, cam, vc, r, res);
            break;
        }
        if (march_axis(hm.m, vc, r, cell, res)) break;
        st[S_BOX_FALLBACKS]++;
        cell = kNone;
    }
}

float hit_depth(const HostMap& hm, const ViewConst& vc, const CastResult& res) {
    return res.rank == kNone ? 0.0f : (float)__dsqrt_rn(dist_sq_at(vc, hm.m.resolution, res.k0, res.k1, res.k2));
}

}  // namespace

extern "C" {

// Dense cast of one view.  variant: 0 PLAIN, 1 FAST (raycast_kernel), 2 AXIS pipeline (cull / coarse / march kernels).
// force_region_cull: -1 = as prv_set_camera decides, 0 / 1 = force off / on.  brick_cs, brick_entry: prv_set_brick_cull.
// hit_rank, depth: [H][W]; stats: S_N counters.  Returns 0, or -1 for bad input.
int koh_cast_view_dense(const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution, const prv_intrinsics* intr, double max_range,
                        const double* pose_world, const double* init_pos, int variant, int force_region_cull, int brick_cs, int brick_entry,
                        uint32_t* hit_rank, float* depth, uint64_t* stats) {
    HostMap hm;
    if (!keys || !intr || !pose_world || !init_pos || !hit_rank || !depth || !stats) return -1;
    if (brick_cs != 4 && brick_cs != 8 && brick_cs != 16) return -1;
    if (!build_map(hm, keys, rgb, N, resolution, max_range, brick_cs)) return -1;
    const bool entry = brick_entry != 0;
    const DevCam cam = make_cam(*intr, max_range, force_region_cull);
    ViewConst vc;
    std::memset(&vc, 0, sizeof(vc));
    make_view_const(hm.setup, pose_world, init_pos, 0, vc);
    for (int i = 0; i < S_N; i++) stats[i] = 0;
    stats[S_FLAGS] = vc.flags;
    stats[S_REGION_OK] = (uint64_t)cam.region_cull_ok;
    const int W = cam.W, H = cam.H;
    for (size_t i = 0; i < (size_t)W * H; i++) {
        hit_rank[i] = kNone;
        depth[i] = 0.0f;
    }
    const bool view_ok = (vc.flags & kViewInMap) && !(vc.flags & kViewInObject);
    if (!view_ok) return 0;
    const bool fast = (vc.flags & kViewFastOk) != 0;
    if (variant != 2) {
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++) {
                CastResult res;
                res.rank = kNone;
                res.steps = res.probes = 0;
                res.k0 = res.k1 = res.k2 = 0;
                RayState r;
                stats[S_RAYS]++;
                stats[S_MARCHED]++;
                if (setup_ray(cam, vc, hm.m.resolution, px, py, r)) {
                    if (variant == 0 || !fast)
                        march_plain(hm.m, cam, vc, r, res);
                    else
                        march_fast(hm.m, vc, r, res);
                }
                stats[S_PROBES] += res.probes;
                stats[S_STEPS] += res.steps;
                if (res.rank != kNone) stats[S_HITS]++;
                hit_rank[(size_t)py * W + px] = res.rank;
                depth[(size_t)py * W + px] = hit_depth(hm, vc, res);
            }
        return 0;
    }
    const int regions_x = (W + 31) >> 5, regions_y = (H + 31) >> 5;
    for (int ry = 0; ry < regions_y; ry++)
        for (int rx = 0; rx < regions_x; rx++) {
            const int x1 = std::min(W, (rx + 1) << 5), y1 = std::min(H, (ry + 1) << 5);
            const uint64_t npix = (uint64_t)(x1 - (rx << 5)) * (uint64_t)(y1 - (ry << 5));
            stats[S_RAYS] += npix;
            if (fast && cam.region_cull_ok) {
                uint32_t bal = 0;
                for (int lane = 0; lane < 32; lane++)
                    if (region_corner_outside(hm.m, cam, vc, rx, ry, lane)) bal |= 1u << lane;
                if (region_skip_from_ballot(bal)) {
                    stats[S_REGION_CULLED] += npix;
                    continue;
                }
            }
            for (int py = ry << 5; py < y1; py++)
                for (int px = rx << 5; px < x1; px++) {
                    CastResult res;
                    axis_pixel(hm, cam, vc, px, py, res, stats, entry);
                    stats[S_PROBES] += res.probes;
                    stats[S_STEPS] += res.steps;
                    if (res.rank != kNone) stats[S_HITS]++;
                    hit_rank[(size_t)py * W + px] = res.rank;
                    depth[(size_t)py * W + px] = hit_depth(hm, vc, res);
                }
        }
    return 0;
}

// Perception_3D::precept of one view (voxel-driven mode): project every occupied voxel, cast the ray through its truncated
// pixel on the (W+1) x (H+1) grid (masked pixels only: no region cull, as cull_kernel<true>), give voxel i its pixel's
// result.  points_out: [N] pcl::PointXYZRGB images; voxel_hit_out: [N] hit rank or PRV_NONE.  *in_map_out as prv_precept.
int koh_precept(const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution, const prv_intrinsics* intr, double max_range,
                const double* pose_world, const double* init_pos, prv_point_xyzrgb* points_out, uint32_t* voxel_hit_out, int* in_map_out) {
    HostMap hm;
    if (!keys || !intr || !pose_world || !init_pos || !points_out || !voxel_hit_out) return -1;
    if (!build_map(hm, keys, rgb, N, resolution, max_range)) return -1;
    const DevCam cam = make_cam(*intr, max_range, -1);
    ViewConst vc;
    std::memset(&vc, 0, sizeof(vc));
    make_view_const(hm.setup, pose_world, init_pos, 0, vc);
    if (in_map_out) *in_map_out = (vc.flags & kViewInMap) ? 1 : 0;
    const bool view_ok = (vc.flags & kViewInMap) && !(vc.flags & kViewInObject);
    const int GW = cam.W + 1, GH = cam.H + 1;
    std::vector<uint32_t> pix_hit((size_t)GW * GH, kNone);
    std::vector<uint8_t> done((size_t)GW * GH, 0);
    uint64_t st[S_N] = {0};
    for (uint32_t i = 0; i < N; i++) {
        uint32_t h = kNone;
        if (view_ok) {  // project_voxels_kernel
            const float ex = (float)key_to_coord_d(keys[3 * i + 0], resolution);
            const float ey = (float)key_to_coord_d(keys[3 * i + 1], resolution);
            const float ez = (float)key_to_coord_d(keys[3 * i + 2], resolution);
            const float vx = (float)row_apply(vc.inv + 0, (double)ex, (double)ey, (double)ez);
            const float vy = (float)row_apply(vc.inv + 4, (double)ex, (double)ey, (double)ez);
            const float vz = (float)row_apply(vc.inv + 8, (double)ex, (double)ey, (double)ez);
            float u, v;
            project_point_to_pixel(cam, vx, vy, vz, u, v);
            if (u >= 0.0f && u <= (float)cam.W && v >= 0.0f && v <= (float)cam.H) {
                const int ix = (int)u, iy = (int)v;
                const size_t pid = (size_t)iy * GW + ix;
                if (!done[pid]) {
                    CastResult res;
                    axis_pixel(hm, cam, vc, ix, iy, res, st);
                    pix_hit[pid] = res.rank;
                    done[pid] = 1;
                }
                h = pix_hit[pid];
            }
        }
        voxel_hit_out[i] = h;
        prv_point_xyzrgb pt;  // precept_points_kernel
        std::memset(&pt, 0, sizeof(pt));
        pt.w = 1.0f;
        pt.a = 255;
        if (h != kNone) {
            pt.x = (float)key_to_coord_d(keys[3 * h + 0], resolution);
            pt.y = (float)key_to_coord_d(keys[3 * h + 1], resolution);
            pt.z = (float)key_to_coord_d(keys[3 * h + 2], resolution);
            pt.r = hm.rgb[3 * h + 0];
            pt.g = hm.rgb[3 * h + 1];
            pt.b = hm.rgb[3 * h + 2];
        }
        points_out[i] = pt;
    }
    return 0;
}

int koh_num_stats() { return S_N; }

}  // extern "C"
