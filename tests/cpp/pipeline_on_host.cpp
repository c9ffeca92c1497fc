// pipeline_on_host.cpp -- TEST INFRASTRUCTURE.  The ray-cast KERNELS of kernels_cast.cuh -- cull_kernel, coarse_kernel,
// march_kernel, and the voxel-mode kernels, source unchanged -- run on the CPU by
// the SIMT emulator of simt_on_host.hpp: queues, tickets, ballots, block / warp barriers, atomics and the coverage-row
// scatter execute as written, so what kernel_on_host.cpp cannot see (the plumbing between the per-ray functions) is held
// against the oracle without a GPU too.  The launch sequence below is cast_impl's (prv_device.cu); the lookup tables are the
// host-built ones that map_kernels_on_host.cpp shows the map kernels to produce.  Small images only: one OS thread per CUDA
// thread.
//
// Build (tests/test_kernel_on_host.py): g++ -O2 -std=c++20 -ffp-contract=off -shared -fPIC -pthread -I/usr/local/cuda/include
#include "kernel_on_host.cpp"
#include "simt_on_host.hpp"

#include "../../nerf-prv_b200/csrc/kernels_common.cuh"
#include "../../nerf-prv_b200/csrc/kernels_cast.cuh"

// splat_resolve_kernel's tile is dynamic shared memory (extern __shared__): here an ordinary global of the largest tile size
#undef __shared__
#define __shared__
unsigned long long s_tile[(32 + 31 + 32) * (8 + 31)];  // tile + row minima
static inline uint32_t __float_as_uint(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
#include "../../nerf-prv_b200/csrc/kernels_ensemble.cuh"
#include "../../nerf-prv_b200/csrc/kernels_splat.cuh"
#undef __shared__
#define __shared__ static
#include "../../nerf-prv_b200/csrc/kernels_greedy.cuh"  // popcount_rows_kernel, greedy_init_kernel, greedy_iter_kernel (PRVK_HOST_CHECK hides the rest)

extern "C" {

// render_impl (prv_device.cu): corner buffer cleared to all ones, splat_points_kernel, splat_resolve_kernel
int poh_render_views(const float* xyz, const uint8_t* rgb, uint64_t P, const prv_intrinsics* intr, const double* pose_world, uint32_t V, int point_size,
                     uint8_t* rgba_out, float* depth_out) {
    if (!xyz || !rgb || !intr || !pose_world || !rgba_out || P == 0 || V == 0 || point_size < 1 || point_size > 32) return -1;
    const DevCam cam = make_cam(*intr, 1.0, 0);
    std::vector<ViewConst> views(V);
    for (uint32_t v = 0; v < V; v++) {  // prv_render_views: only the inverse pose is needed
        std::memset(&views[v], 0, sizeof(ViewConst));
        const prv::Matrix4d pw = prv::Matrix4d::FromRowMajor(pose_world + 16 * (size_t)v);
        const prv::Matrix4d inv = pw.inverse();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) {
                views[v].pose[4 * r + c] = pw(r, c);
                views[v].inv[4 * r + c] = inv(r, c);
            }
    }
    const int W = cam.W, H = cam.H, Wc = W + point_size - 1, Hc = H + point_size - 1;
    std::vector<unsigned long long> corner((size_t)V * Wc * Hc, ~0ull);
    const float focal = (float)((double)intr->height * (double)intr->fy / (2.0 * (double)(int)intr->ppy));  // prv_splat_focal
    simt_launch(dim3((unsigned)((P + 255) / 256), V), dim3(256),
                [&] { splat_points_kernel(xyz, P, cam, views.data(), 0u, focal, point_size, corner.data(), Wc, Hc); });
    simt_launch(dim3((W + 31) / 32, (H + 7) / 8, V), dim3(256),
                [&] { splat_resolve_kernel(corner.data(), Wc, Hc, W, H, point_size, rgb, rgba_out, depth_out, 0u); });
    return 0;
}

// popcount_rows_kernel (coverage counts, always on the path) and the one-launch-per-iteration greedy of greedy_impl
// (greedy_init_kernel + greedy_iter_kernel: the path for coverage masks too large for the cluster / grid-barrier kernels).
// rows [V][words] u64, ids[V] = global view ids.  seq / gains need room for max_iter + 1 entries; returns their number.
int poh_counts_and_greedy(const uint64_t* rows, uint32_t V, uint32_t words, const uint32_t* ids, uint32_t first_view, uint32_t max_iter,
                          uint32_t* counts_out, uint32_t* seq_out, uint32_t* gains_out, uint64_t* covered_out) {
    if (!rows || !ids || !counts_out || !seq_out || !gains_out || V == 0 || (words & 1u)) return -1;
    simt_launch(dim3(V), dim3(256), [&] { popcount_rows_kernel(rows, words, counts_out); });
    uint32_t max_id = 0;
    for (uint32_t v = 0; v < V; v++) max_id = std::max(max_id, ids[v]);
    std::vector<uint32_t> row_of_id((size_t)max_id + 1, kNone);
    for (uint32_t v = 0; v < V; v++) row_of_id[ids[v]] = v;
    if (first_view > max_id || row_of_id[first_view] == kNone) return -1;
    std::vector<unsigned long long> best((size_t)max_iter + 2, 0ull);
    std::vector<uint64_t> cov[2] = {std::vector<uint64_t>(words, 0), std::vector<uint64_t>(words, 0)};
    simt_launch(dim3(1), dim3(256), [&] { greedy_init_kernel(rows, words, row_of_id[first_view], first_view, best.data()); });
    for (uint32_t k = 1; k <= max_iter + 1; k++) {
        const int cover_only = k == max_iter + 1;
        simt_launch(dim3(cover_only ? 1 : V), dim3(256), [&] {
            greedy_iter_kernel(rows, words, ids, row_of_id.data(), k, best.data(), cov[(k - 1) & 1].data(), cov[k & 1].data(), cover_only);
        });
    }
    uint32_t n = 0;  // prv_get_greedy
    for (uint32_t k = 0; k <= max_iter; k++) {
        const uint32_t g = (uint32_t)(best[k] >> 32);
        if (k > 0 && g == 0) break;
        seq_out[n] = 0xFFFFFFFFu - (uint32_t)(best[k] & 0xFFFFFFFFull);
        gains_out[n] = g;
        n++;
    }
    if (covered_out) std::memcpy(covered_out, cov[n & 1].data(), (size_t)words * 8);
    return (int)n;
}

// prv_score_ensemble's two kernels (method 2 with E == 2 uses the host-libm table of logs, as the library does)
int poh_score_ensemble(const uint8_t* images, uint32_t V, uint32_t E, int W, int H, int method, double* scores_out) {
    if (!images || !scores_out || V == 0 || E == 0 || (method != 2 && method != 3)) return -1;
    const uint32_t npix = (uint32_t)W * (uint32_t)H;
    std::vector<double> terms((size_t)V * npix * 3, 0.0);
    double lut[256];
    lut[0] = 0.0;
    for (int d = 1; d < 256; d++) lut[d] = std::log(((double)d * (double)d) / 4.0);
    const double* lut_p = (method == 2 && E == 2) ? lut : nullptr;
    simt_launch(dim3((npix + 255) / 256, V), dim3(256), [&] { ensemble_terms_kernel(images, E, npix, method, lut_p, terms.data()); });
    simt_launch(dim3((V + 31) / 32), dim3(32), [&] { ensemble_sum_kernel(terms.data(), V, npix, scores_out); });
    return 0;
}

// mode: PRV_MODE_DENSE / PRV_MODE_VOXEL.  grid_blocks: size of the persistent grids (sm_count * occupancy on the device).
// Outputs: bitsets [V][words64] u64, pix_hit / pix_depth [V][GH][GW] (GW x GH = W x H dense, (W+1) x (H+1) voxel; depth only
// dense), stats [V][4] (rays, probes, hits, steps), marched [V], voxel_hit [V][N] (voxel mode only, else may be null).
int poh_cast_views(const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution, const prv_intrinsics* intr, double max_range,
                   const double* pose_world, const double* init_pos, uint32_t V, int mode, int brick_cs, int brick_entry, int grid_blocks,
                   uint64_t* bitsets_out, uint32_t* pix_hit_out, float* pix_depth_out, unsigned long long* stats_out, uint32_t* marched_out,
                   uint32_t* voxel_hit_out) {
    HostMap hm;
    if (!keys || !intr || !pose_world || !init_pos || !bitsets_out || !pix_hit_out || !stats_out || !marched_out || V == 0 || V > (uint32_t)kMaxViewsPerLaunch)
        return -1;
    if (brick_cs != 4 && brick_cs != 8 && brick_cs != 16) return -1;
    if (!build_map(hm, keys, rgb, N, resolution, max_range, brick_cs)) return -1;
    const DevCam cam = make_cam(*intr, max_range, -1);
    std::vector<ViewConst> views(V);
    for (uint32_t v = 0; v < V; v++) {
        std::memset(&views[v], 0, sizeof(ViewConst));
        make_view_const(hm.setup, pose_world + 16 * (size_t)v, init_pos + 3 * (size_t)v, v, views[v]);
    }
    const bool voxel = mode == PRV_MODE_VOXEL;
    CastParams p{};
    p.map = hm.m;
    p.cam = cam;
    p.views = views.data();
    p.GW = voxel ? cam.W + 1 : cam.W;
    p.GH = voxel ? cam.H + 1 : cam.H;
    p.pix_stride = (unsigned long long)p.GW * p.GH;
    const uint32_t words = hm.m.words64;
    std::vector<uint32_t> bitsets32((size_t)V * words * 2, 0u), queue((size_t)V * (size_t)(((p.GW + 31) / 32) * ((p.GH + 31) / 32))), queue2((size_t)V * p.pix_stride), qcount(2 * (size_t)V, 0u),
        tickets(2, 0u), mask, voxel_pix;
    std::vector<unsigned long long> stats((size_t)V * 4, 0ull);
    std::vector<uint32_t> queue2b;
    p.bitsets32 = bitsets32.data();
    p.stats = stats.data();
    p.pix_hit = pix_hit_out;
    p.pix_depth = voxel ? nullptr : pix_depth_out;
    for (size_t i = 0; i < (size_t)V * p.pix_stride; i++) pix_hit_out[i] = 0xABABABABu;  // every ray of a live view must be written by some kernel
    p.queue = queue.data();
    p.qcount = qcount.data();
    p.queue2 = queue2.data();
    p.qcount2 = qcount.data() + V;
    p.queue_cap = p.pix_stride;
    p.rqueue_cap = (uint32_t)(((p.GW + 31) / 32) * ((p.GH + 31) / 32));
    p.tickets = tickets.data();
    p.view_base = 0;
    p.nviews = V;
    if (brick_entry) {
        queue2b.assign((size_t)V * p.pix_stride, 0u);
        p.queue2b = queue2b.data();
    }
    if (voxel) {
        p.mask_words = (uint32_t)((p.pix_stride + 31) / 32);
        mask.assign((size_t)V * p.mask_words, 0u);
        voxel_pix.assign((size_t)V * N, 0u);
        p.mask = mask.data();
        simt_launch(dim3((N + 255) / 256, V), dim3(256), [&] { project_voxels_kernel(hm.m, cam, views.data(), 0u, mask.data(), p.mask_words, voxel_pix.data()); });
    }
    const dim3 rgrid((unsigned)(((p.GW + 31) / 32 + kCullRegions - 1) / kCullRegions), (unsigned)((p.GH + 31) / 32), V);
    if (voxel)
        simt_launch(rgrid, dim3(256), [&] { cull_kernel<true>(p); });
    else
        simt_launch(rgrid, dim3(256), [&] { cull_kernel<false>(p); });
    if (voxel)
        simt_launch(dim3((unsigned)grid_blocks), dim3(256), [&] { coarse_kernel<8, true>(p); });
    else
        simt_launch(dim3((unsigned)grid_blocks), dim3(256), [&] { coarse_kernel<8, false>(p); });
    simt_launch(dim3((unsigned)grid_blocks), dim3(kMarchBlock), [&] { march_kernel<kMarchBlock, kMarchMinBlocks>(p); });
    if (voxel && voxel_hit_out)
        simt_launch(dim3((N + 255) / 256, V), dim3(256), [&] { gather_voxel_hits_kernel(N, voxel_pix.data(), pix_hit_out, p.pix_stride, voxel_hit_out); });
    std::memcpy(bitsets_out, bitsets32.data(), (size_t)V * words * 8);
    std::memcpy(stats_out, stats.data(), (size_t)V * 4 * 8);
    for (uint32_t v = 0; v < V; v++) marched_out[v] = qcount[V + v];
    return 0;
}

int poh_words(uint32_t N) {
    uint32_t w = (N + 63) / 64;
    if (w == 0) w = 1;
    return (int)((w + 1) & ~1u);
}

}  // extern "C"
