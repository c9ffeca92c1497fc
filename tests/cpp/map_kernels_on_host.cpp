// map_kernels_on_host.cpp -- TEST INFRASTRUCTURE.  Runs the barrier-free map-build kernels of kernels_map.cuh
// (map_scatter_kernel, map_shell_kernel, map_rank_kernel) -- the kernel source, unchanged -- on the CPU:
// blocks and threads are executed one after the other with threadIdx / blockIdx as plain globals and atomics as plain
// read-modify-writes, and the tables they leave must equal, word for word, the tables kernel_on_host.cpp builds from the
// layout documented in DESIGN.md section 3 (which the per-ray checks run on).  Closes the loop for the lookup tables
// without a GPU, for every brick size of prv_set_brick_cull, and runs clean under ASan.
#include "kernel_on_host.cpp"

#include <cstdlib>

// ---- just enough of the CUDA execution model for kernels without barriers ----------------------------------------------
static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
template <class F>
static void launch(dim3 grid, dim3 block, F kernel) {
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++)
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++) {
                            blockIdx = {bx, by, bz};
                            threadIdx = {tx, ty, tz};
                            kernel();
                        }
}
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p = old | v;
    return old;
}
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p = std::min(old, v);
    return old;
}
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p = std::max(old, v);
    return old;
}
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
// the prefix kernels need warp shuffles and block barriers: never launched here (the prefix is a plain scan, done on the host)
static inline uint32_t __shfl_down_sync(unsigned, uint32_t, int) { std::abort(); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t, int) { std::abort(); }
static inline void __syncthreads() { std::abort(); }
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

// kernels_cast.cuh is included for its three barrier-free voxel-mode kernels; the cull / coarse / march kernels in the same
// file need warp intrinsics, which only have to compile here
static inline uint32_t __ballot_sync(unsigned, bool) { std::abort(); }
static inline uint32_t __shfl_sync(unsigned, uint32_t, int) { std::abort(); }
static inline uint32_t __reduce_add_sync(unsigned, uint32_t) { std::abort(); }
static inline void __syncwarp() { std::abort(); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p = old + v;
    return old;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = old + v;
    return old;
}

#include "../../nerf-prv_b200/csrc/kernels_common.cuh"
#include "../../nerf-prv_b200/csrc/kernels_cast.cuh"
#include "../../nerf-prv_b200/csrc/kernels_map.cuh"

extern "C" {

// returns 0 when every table written by the kernels equals the host-built one; otherwise the number of the first differing table
int mkh_check_map_kernels(const uint16_t* keys, uint32_t N, double resolution, int brick_cs) {
    HostMap hm;
    if (!build_map(hm, keys, nullptr, N, resolution, 1.0, brick_cs)) return -1;
    const DevMap& m = hm.m;
    // zeroed tables of exactly the sizes prv_set_map allocates
    std::vector<uint32_t> bitmap(hm.bitmap.size(), 0u), pad(hm.pad.size(), 0u), coarse(hm.coarse.size(), 0u);
    std::vector<uint32_t> leaf_of_raster(N, 0xDEADBEEFu);
    MapBuild b{};
    for (int a = 0; a < 3; a++) {
        b.lo[a] = m.lo[a];
        b.n[a] = m.n[a];
        b.nc[a] = m.nc[a];
    }
    b.wx = m.wx;
    b.row_log2 = m.pad_row_log2;
    b.n_occ = N;
    b.slack_bits = m.pad_bit_offset;
    b.nwords = bitmap.size();
    b.bitmap = bitmap.data();
    b.pad = pad.data();
    b.coarse = coarse.data();
    b.prefix = hm.prefix.data();  // host scan of the kernel-built bitmap, checked below
    b.leaf_of_raster = leaf_of_raster.data();
    b.keys = hm.keys.data();
    b.cs = brick_cs;
    {   // map_check_kernel: leaf order + AABB of the key table (prv_set_map's validation)
        uint32_t chk[7] = {N, 65535u, 65535u, 65535u, 0u, 0u, 0u};
        launch(dim3((N + 255) / 256), dim3(256), [&] { map_check_kernel(hm.keys.data(), N, chk); });
        if (chk[0] != N) return 6;
        for (int a = 0; a < 3; a++)
            if ((int)chk[1 + a] != m.lo[a] || (int)chk[4 + a] != m.lo[a] + m.n[a] - 1) return 7;
        if (N >= 3) {  // a swapped pair must be reported at the index of the first key that is not above its predecessor
            std::vector<uint16_t> bad(hm.keys);
            for (int a = 0; a < 3; a++) std::swap(bad[3 * (size_t)(N / 2) + a], bad[3 * (size_t)(N / 2 + 1) + a]);
            uint32_t chk2[7] = {N, 65535u, 65535u, 65535u, 0u, 0u, 0u};
            launch(dim3((N + 255) / 256), dim3(256), [&] { map_check_kernel(bad.data(), N, chk2); });
            if (chk2[0] != N / 2 + 1) return 8;
        }
    }
    launch(dim3((N + 255) / 256), dim3(256), [&] { map_scatter_kernel(b); });
    const size_t pad_rows = (size_t)(m.n[1] + 2) * (m.n[2] + 2);
    launch(dim3((unsigned)((pad_rows + 255) / 256)), dim3(256), [&] { map_shell_kernel(b); });
    if (bitmap != hm.bitmap) return 1;
    if (pad != hm.pad) return 2;
    if (coarse != hm.coarse) return 3;
    launch(dim3((N + 255) / 256), dim3(256), [&] { map_rank_kernel(b); });
    if (leaf_of_raster != hm.leaf_of_raster) return 5;
    return 0;
}

// Perception_3D::precept of one view with the three voxel-mode kernels run as kernels (project_voxels_kernel ->
// [per masked pixel: the AXIS chain, as koh_precept] -> gather_voxel_hits_kernel -> precept_points_kernel).
int mkh_precept(const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution, const prv_intrinsics* intr, double max_range,
                const double* pose_world, const double* init_pos, prv_point_xyzrgb* points_out, uint32_t* voxel_hit_out) {
    HostMap hm;
    if (!build_map(hm, keys, rgb, N, resolution, max_range)) return -1;
    const DevCam cam = make_cam(*intr, max_range, -1);
    ViewConst vc;
    std::memset(&vc, 0, sizeof(vc));
    make_view_const(hm.setup, pose_world, init_pos, 0, vc);
    const int GW = cam.W + 1, GH = cam.H + 1;
    const unsigned long long pix_stride = (unsigned long long)GW * GH;
    const uint32_t mask_words = (uint32_t)((pix_stride + 31) / 32);
    std::vector<uint32_t> mask(mask_words, 0u), voxel_pix(N, 0u), pix_hit(pix_stride, kNone);
    launch(dim3((N + 255) / 256, 1), dim3(256), [&] { project_voxels_kernel(hm.m, cam, &vc, 0u, mask.data(), mask_words, voxel_pix.data()); });
    uint64_t st[S_N] = {0};
    for (unsigned long long pid = 0; pid < pix_stride; pid++)
        if ((mask[pid >> 5] >> (pid & 31)) & 1u) {
            CastResult res;
            axis_pixel(hm, cam, vc, (int)(pid % GW), (int)(pid / GW), res, st);
            pix_hit[pid] = res.rank;
        }
    launch(dim3((N + 255) / 256, 1), dim3(256), [&] { gather_voxel_hits_kernel(N, voxel_pix.data(), pix_hit.data(), pix_stride, voxel_hit_out); });
    launch(dim3((N + 255) / 256), dim3(256), [&] { precept_points_kernel(hm.m, voxel_hit_out, points_out); });
    return 0;
}

// GPU ingest (prv_set_map_from_cloud) with its three kernels run as kernels and the two cub primitives between them replaced
// by their definitions (stable sort of (code, index) pairs on the 49-bit code; exclusive prefix sum).
// keys_out / rgb_out need room for P entries; returns the number of voxels, or -1.
int mkh_ingest(const float* xyz, const uint8_t* rgb, uint32_t P, double resolution, uint16_t* keys_out, uint8_t* rgb_out) {
    if (!xyz || !rgb || P == 0) return -1;
    std::vector<unsigned long long> codes(P), sorted_codes(P);
    std::vector<uint32_t> index(P), sorted_index(P), head(P + 1), pos(P + 1);
    launch(dim3((P + 255) / 256), dim3(256), [&] { ingest_keys_kernel(xyz, P, 1.0 / resolution, codes.data(), index.data()); });
    std::vector<uint32_t> order(P);
    for (uint32_t i = 0; i < P; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
    for (uint32_t i = 0; i < P; i++) {
        sorted_codes[i] = codes[order[i]];
        sorted_index[i] = index[order[i]];
    }
    launch(dim3((P + 1 + 255) / 256), dim3(256), [&] { ingest_heads_kernel(sorted_codes.data(), P, head.data()); });
    uint32_t run = 0;
    for (uint32_t i = 0; i <= P; i++) {
        pos[i] = run;
        run += head[i];
    }
    launch(dim3((P + 255) / 256), dim3(256),
           [&] { ingest_compact_kernel(sorted_codes.data(), sorted_index.data(), head.data(), pos.data(), P, rgb, keys_out, rgb_out); });
    return (int)pos[P];
}

}  // extern "C"
