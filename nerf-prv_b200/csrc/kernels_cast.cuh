// kernels_cast.cuh -- ray-cast kernels: cull / coarse / march pipeline (variant AXIS), single-kernel PLAIN/FAST variants, voxel-mode projection and gathers
// Part of the single translation unit prv_device.cu (included there, in order); see DESIGN.md section 4.
#pragma once

// deterministic counters, one set per view: block-reduce, then one atomic per counter per block (per-warp atomics on
// one address serialise in L2 and were measured to bound the whole kernel)
__device__ __forceinline__ void commit_stats(unsigned long long* view_stats, uint32_t rays, uint32_t probes, uint32_t hits, uint32_t steps) {
    __shared__ uint32_t s_cnt[4][8];
    uint32_t c[4] = {rays, probes, hits, steps};
#pragma unroll
    for (int i = 0; i < 4; i++)
        for (int o = 16; o > 0; o >>= 1) c[i] += __shfl_down_sync(0xFFFFFFFFu, c[i], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int i = 0; i < 4; i++) s_cnt[i][warp] = c[i];
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_cnt[threadIdx.x][w];
        if (t) atomicAdd(view_stats + threadIdx.x, t);
    }
}

// warp-level variant: shuffle-reduce, lanes 1..3 post probes / hits / steps (slot 0 = rays belongs to the cull kernel)
__device__ __forceinline__ void warp_commit_stats(unsigned long long* view_stats, uint32_t probes, uint32_t hits, uint32_t steps) {
    probes = __reduce_add_sync(0xFFFFFFFFu, probes);
    hits = __reduce_add_sync(0xFFFFFFFFu, hits);
    steps = __reduce_add_sync(0xFFFFFFFFu, steps);
    const int lane = threadIdx.x & 31;
    const uint32_t v = lane == 1 ? probes : (lane == 2 ? hits : steps);
    if (lane >= 1 && lane <= 3 && v) atomicAdd(view_stats + lane, (unsigned long long)v);
}

__device__ __forceinline__ void load_view_const(ViewConst& dst, const ViewConst* src) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(&dst);
    for (int i = threadIdx.x; i < (int)(sizeof(ViewConst) / 4); i += blockDim.x) d32[i] = s32[i];
    __syncthreads();
}

// the cull / coarse kernels only need the first kViewCullWords words (float pose, origin, origin key, flags)
__device__ __forceinline__ void load_view_prefix(ViewConst& dst, const ViewConst* src) {
    if (threadIdx.x < kViewCullWords) reinterpret_cast<uint32_t*>(&dst)[threadIdx.x] = reinterpret_cast<const uint32_t*>(src)[threadIdx.x];
    __syncthreads();
}

__device__ __forceinline__ void write_hit(const CastParams& p, const ViewConst& vc, uint32_t view, unsigned long long pid, const CastResult& res) {
    if (res.rank != kNone) {
        uint32_t* row = p.bitsets32 + (size_t)view * (2u * p.map.words64);
        const uint32_t bit = 1u << (res.rank & 31);
        uint32_t* w = row + (res.rank >> 5);
        if (!(*reinterpret_cast<volatile uint32_t*>(w) & bit)) atomicOr(w, bit);
    }
    if (p.pix_hit) {
        p.pix_hit[(size_t)view * p.pix_stride + pid] = res.rank;
        if (p.pix_depth) {
            float d = 0.0f;
            if (res.rank != kNone) d = (float)__dsqrt_rn(dist_sq_at(vc, p.map.resolution, res.k0, res.k1, res.k2));
            p.pix_depth[(size_t)view * p.pix_stride + pid] = d;
        }
    }
}

// ---- AXIS pipeline ------------------------------------------------------------------------------------------------
// kernel 1 (cull_kernel):   every 32x32-pixel region against the AABB; dismissed regions are filled with "no hit", the rest -> region queue
// kernel 2 (coarse_kernel): every pixel of the queued regions, conservative slab test + coarse-brick walk; survivors -> queue 2
// kernel 3 (march_kernel):  dense warps over queue 2, the exact castRay march
// Kernels 2 and 3 are persistent: kernel 2's blocks pull 256-pixel row-tiles, kernel 3's warps 32-ray chunks of a flattened
// (view, chunk) list with an atomic ticket, so expensive and cheap chunks balance across the 148 SMs and there is no
// partial last wave.

// One block per kCullRegions consecutive 32x32-pixel regions of one row of one view (blockIdx = (region group, region
// row, view)): the view constants are fetched once and the region tests of the group run side by side (one warp each),
// so the ~1.5 us of serial latency at the start of a block (constants from L2, barrier, region test) is paid once per
// group instead of once per region (ncu had half of all stall samples there).
// Region test: the region's rays lie inside the pyramid spanned by the four corner rays taken 2 px outside the region
// (the pixel->direction map is affine up to the lens distortion, whose deviation inside any region was verified on the
// host to stay within that margin).  If all eight corners of the AABB grown by 2 voxels lie outside one side plane of the
// pyramid, no ray of the region can touch the AABB: its pixels get "no hit" with 128-bit stores and nothing else.
// Every other region is appended to the view's REGION queue: its pixels are the coarse kernel's work.  (Round 1 ran a float
// slab test per pixel here and queued the surviving pixels; the coarse kernel then recomputed the direction and a tighter slab
// test for each of them -- on C3 three quarters of this kernel's instructions were that duplicate per-pixel work, for a test
// that removed 20 % of the rays it looked at.)
constexpr int kCullRegions = 4;
template <bool MASKED>
__global__ void __launch_bounds__(256, 8) cull_kernel(const CastParams p) {
    __shared__ ViewConst s_vc;  // cull prefix only
    __shared__ int s_skip[kCullRegions];
    const uint32_t view = blockIdx.z + p.view_base;
    load_view_prefix(s_vc, p.views + view);
    const ViewConst& vc = s_vc;
    const int regions_x = (p.GW + 31) >> 5;
    const int region_y = blockIdx.y, rx0 = blockIdx.x * kCullRegions;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool view_ok = (vc.flags & kViewInMap) && !(vc.flags & kViewInObject);
    const bool fast = (vc.flags & kViewFastOk) != 0;

    if (warp < kCullRegions) {
        // warp r tests region rx0 + r; lane = plane (0..3) * 8 + box corner (0..7): 32 dot products, one ballot
        bool skip = false;
        const int region_x = rx0 + warp;
        if (!MASKED && region_x < regions_x && view_ok && fast && p.cam.region_cull_ok) {
            const bool outside = region_corner_outside(p.map, p.cam, vc, region_x, region_y, lane);
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, outside);
            skip = region_skip_from_ballot(bal);
        }
        if (lane == 0) {
            s_skip[warp] = skip ? 1 : 0;
            // a live region of a live view goes to the coarse kernel (a dead view's pixels keep whatever the caller reads as
            // "view out of map": nothing is cast, the tables are written below as no hit)
            if (!skip && region_x < regions_x && view_ok) {
                const uint32_t pos = atomicAdd(p.qcount + view, 1u);
                p.queue[(size_t)view * p.rqueue_cap + pos] = ((uint32_t)region_y << 16) | (uint32_t)region_x;
            }
        }
    }
    __syncthreads();
    const bool vec_ok = ((p.GW & 3) == 0) && ((p.pix_stride & 3ull) == 0ull);
    uint32_t skipped_rays = 0;  // thread 0: rays of the regions skipped as a whole
    for (int r = 0; r < kCullRegions; r++) {
        const int region_x = rx0 + r;
        if (region_x >= regions_x) break;  // uniform
        if (s_skip[r] == 0 && view_ok) continue;  // queued
        // the whole region provably misses (or the view casts nothing): record "no hit" for its pixels
        if (p.pix_hit && !MASKED) {
            if (vec_ok) {  // thread -> row t/8, 4 consecutive pixels: one 128-bit store per array
                const int px = (region_x << 5) + ((threadIdx.x & 7) << 2), py = (region_y << 5) + (threadIdx.x >> 3);
                if (px < p.GW && py < p.GH) {
                    const size_t o = (size_t)view * p.pix_stride + (size_t)py * p.GW + px;
                    *reinterpret_cast<uint4*>(p.pix_hit + o) = make_uint4(kNone, kNone, kNone, kNone);
                    if (p.pix_depth) *reinterpret_cast<float4*>(p.pix_depth + o) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            } else {
                const int px = (region_x << 5) + ((warp & 3) << 3) + (lane & 7);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int py = (region_y << 5) + (t << 3) + ((warp >> 2) << 2) + (lane >> 3);
                    if (px < p.GW && py < p.GH) {
                        const size_t o = (size_t)view * p.pix_stride + (size_t)py * p.GW + px;
                        p.pix_hit[o] = kNone;
                        if (p.pix_depth) p.pix_depth[o] = 0.0f;
                    }
                }
            }
        }
        if (view_ok) skipped_rays += (uint32_t)(min(32, p.GW - (region_x << 5)) * min(32, p.GH - (region_y << 5)));
    }
    if (threadIdx.x == 0 && skipped_rays) atomicAdd(p.stats + 4 * (size_t)view, (unsigned long long)skipped_rays);
}

// exclusive prefix of chunk counts over the views of this launch -> s_prefix[0..nviews]; a view has ceil(count / chunk) chunks,
// or count * per_item chunks when per_item > 0 (the coarse kernel: 4 row-tiles of 256 pixels per queued region)
__device__ __forceinline__ void build_chunk_prefix(const uint32_t* counts, uint32_t nviews, uint32_t* s_prefix, uint32_t chunk, uint32_t per_item = 0) {
    __shared__ uint32_t s_part[8];
    // each thread owns a contiguous run of views
    const uint32_t per = (nviews + blockDim.x - 1) / blockDim.x;
    const uint32_t b = threadIdx.x * per, e = min(nviews, b + per);
    uint32_t sum = 0;
    for (uint32_t v = b; v < e; v++) sum += per_item ? counts[v] * per_item : (counts[v] + chunk - 1u) / chunk;
    // block exclusive scan of the per-thread sums
    uint32_t incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_part[w];
    uint32_t run = woff + incl - sum;
    for (uint32_t v = b; v < e; v++) {
        s_prefix[v] = run;
        run += per_item ? counts[v] * per_item : (counts[v] + chunk - 1u) / chunk;
    }
    if (threadIdx.x == blockDim.x - 1) s_prefix[nviews] = woff + incl;
    __syncthreads();
}

// Persistent WARPS over the (view, region row-tile) list: per pixel the approximate direction and the conservative brick walk
// (coarse_miss; its own slab test against the AABB grown by one voxel is the first thing it does).  Survivors -> queue 2.
// A warp takes a whole 32x8 row-tile (or the four row-tiles of a region, when every warp still gets 32+ tickets: on the 1024-view
// workload a warp that takes single tiles changes view on most of them), walks its eight 8x4 pixel patches one after the other,
// collects the survivors in its own 256-entry shared-memory stage and appends them to queue 2 as ONE run per tile with one atomic
// -- no block barrier after the chunk-prefix table is built.  Round 2 went through three forms of this kernel
// (profiles/r2_march_ab.md): blocks of eight warps per tile with a block-wide compaction (2-4 barriers per tile: ~30 % of the
// kernel's stall samples sat at those barriers); per-warp appends of single patches (no barriers, but the runs of different
// blocks interleave in the queue, a 32-ray chunk of the march is no longer one patch, and march_kernel lost 13 % on C2 / 19 % on
// C3 to divergence); and this one, which keeps the run-per-tile order of the first and the independence of the second (C3 cull +
// coarse 5.34 -> 4.95 ms).  A warp needs a tile's time for its last ticket, though: with few tiles per warp (C2: 3.5) the tail costs
// more than the barriers did (C2 0.190 -> 0.222 ms), so small casts keep the block form -- coarse_kernel picks by the tile count.
#ifndef PRV_COARSE_WARP_TILES
#define PRV_COARSE_WARP_TILES 12
#endif
#ifndef PRV_TICKET_SPREAD
#define PRV_TICKET_SPREAD 32  // tickets every warp must still get before a ticket grows beyond one tile / one chunk
#endif
// (both overridable so that the CPU checker can force the large-cast paths on its small scenes: tests/test_kernel_on_host.py)
constexpr uint32_t kCoarseWarpTiles = PRV_COARSE_WARP_TILES;  // tiles per warp from which the warps work on their own
constexpr uint32_t kTicketSpread = PRV_TICKET_SPREAD;
#ifndef PRV_COARSE_MINB
#define PRV_COARSE_MINB 8
#endif
constexpr int kCoarseMinBlocks = PRV_COARSE_MINB;  // resident blocks per SM the register budget is set for (8: 32 registers)
constexpr int kCoarseStage = 256;
template <bool MASKED>
__device__ __forceinline__ void coarse_tiles_by_warp(const CastParams& p, const uint32_t* s_prefix, const uint32_t total) {
    __shared__ __align__(16) uint32_t s_vcw[8][kViewCullWords];  // the cull prefix of each warp's current view (96 B rows)
    __shared__ uint32_t s_pid[8][kCoarseStage], s_cell[8][kCoarseStage];
    const uint32_t per = total / (gridDim.x * 8u * kTicketSpread) >= 4u ? 4u : 1u;  // (a view's tiles start at a multiple of four)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ViewConst& vc = *reinterpret_cast<const ViewConst*>(s_vcw[warp]);  // (prefix words only)
    uint32_t* const stage_pid = s_pid[warp];
    uint32_t* const stage_cell = s_cell[warp];
    uint32_t cur_view = 0xFFFFFFFFu;
    uint32_t vl = 0;
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(p.tickets + 0, 1u);
    next = __shfl_sync(0xFFFFFFFFu, next, 0);
    for (;;) {
        const uint32_t g0 = next * per;
        if (g0 >= total) break;
        if (lane == 0) next = atomicAdd(p.tickets + 0, 1u);  // consumed after this ticket: the round trip overlaps the walk
        while (s_prefix[vl + 1] <= g0) vl++;                 // a warp's tickets grow monotonically: amortised O(1)
        const uint32_t view = vl + p.view_base;
        if (view != cur_view) {
            __syncwarp();
            if (lane < kViewCullWords) s_vcw[warp][lane] = reinterpret_cast<const uint32_t*>(p.views + view)[lane];
            __syncwarp();
            cur_view = view;
        }
#pragma unroll 1
        for (uint32_t g = g0; g < g0 + per; g++) {
            // one 32x8 row-tile of a queued region, as eight 8x4 patches
            const uint32_t c = g - s_prefix[vl];
            const uint32_t region = p.queue[(size_t)view * p.rqueue_cap + (c >> 2)];
            const int x0 = (int)((region & 0xFFFFu) << 5), y0 = (int)((region >> 16) << 5) + (int)((c & 3u) << 3);
            uint32_t n = 0, nrays = 0;
#pragma unroll 1
            for (int patch = 0; patch < 8; patch++) {
                const int px = x0 + ((patch & 3) << 3) + (lane & 7);
                const int py = y0 + ((patch >> 2) << 2) + (lane >> 3);
                bool active = px < p.GW && py < p.GH;
                if (MASKED && active) {
                    const unsigned long long lin = (unsigned long long)py * p.GW + px;
                    const uint32_t w = __ldg(p.mask + (size_t)view * p.mask_words + (uint32_t)(lin >> 5));
                    active = (w >> (lin & 31)) & 1u;
                }
                bool keep = false;
                uint32_t cell = kNone;
                if (active) {
                    if (!(vc.flags & kViewFastOk)) {
                        keep = true;  // this view needs the literal march (max-range test): no cull
                    } else {
                        float dx, dy, dz;
                        ray_direction_approx_px(p.cam, vc, px, py, dx, dy, dz);
                        keep = !coarse_miss(p.map, vc, dx, dy, dz, cell);
                    }
                    if (!keep && p.pix_hit) {
                        const size_t o = (size_t)view * p.pix_stride + (size_t)py * p.GW + px;
                        p.pix_hit[o] = kNone;
                        if (p.pix_depth) p.pix_depth[o] = 0.0f;
                    }
                }
                if (MASKED) nrays += __popc(__ballot_sync(0xFFFFFFFFu, active));
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    const uint32_t i = n + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                    stage_pid[i] = ((uint32_t)py << 16) | (uint32_t)px;  // queues carry (y,x) packed: no division downstream
                    stage_cell[i] = cell;
                }
                n += (uint32_t)__popc(bal);
            }
            if (!MASKED) nrays = (uint32_t)(max(0, min(32, p.GW - x0)) * max(0, min(8, p.GH - y0)));  // dense mode: by geometry
            __syncwarp();
            // the tile's survivors -> queue 2 (+ entry brick): one run, one atomic
            uint32_t base = 0;
            if (lane == 0) {
                if (nrays) atomicAdd(p.stats + 4 * (size_t)view, (unsigned long long)nrays);
                if (n) base = atomicAdd(p.qcount2 + view, n);
            }
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            for (uint32_t i = (uint32_t)lane; i < n; i += 32u) {
                const size_t pos = (size_t)view * p.queue_cap + base + i;
                p.queue2[pos] = stage_pid[i];
                if (p.queue2b) p.queue2b[pos] = stage_cell[i];
            }
            __syncwarp();  // (the stage is rewritten by the next tile)
        }
        next = __shfl_sync(0xFFFFFFFFu, next, 0);
    }
}

template <bool MASKED>
__device__ __forceinline__ void coarse_tiles_by_block(const CastParams& p, const uint32_t* s_prefix, const uint32_t total) {
    __shared__ ViewConst s_vc;
    // Small casts (a few tiles per warp: C2 has 3.5): the eight warps of a block share a tile, one 8x4 patch each, so that the tail of
    // the kernel is a patch long, not a tile.  Two block barriers per tile.  The ticket of the next tile is fetched by thread 0 while the block works on this one and is
    // published by this tile's barriers (double-buffered slots); the survivors of a tile are appended to queue 2 as ONE run in
    // warp order, with one atomic per tile.  (Round 2 measured per-warp appends -- one barrier per tile, one atomic per warp:
    // coarse_kernel -2 %, but the runs of different blocks interleave in the queue, a 32-ray chunk of the march is no longer
    // one 8x4 pixel patch, and march_kernel lost 13 % on C2 / 19 % on C3 to divergence: profiles/r2_march_ab.md.)
    __shared__ uint32_t s_ticket[2], s_vl[2], s_wcnt[2][8], s_base[2][8];
    uint32_t cur_view = 0xFFFFFFFFu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(p.tickets + 0, 1u);
        uint32_t v = 0;
        if (t < total)
            while (s_prefix[v + 1] <= t) v++;
        s_ticket[0] = t;
        s_vl[0] = v;
    }
    __syncthreads();
    for (uint32_t it = 0;; it++) {
        const uint32_t slot = it & 1u;
        const uint32_t g = s_ticket[slot];
        if (g >= total) break;
        const uint32_t vl = s_vl[slot];
        const uint32_t view = vl + p.view_base;
        if (view != cur_view) {  // (every warp passed the previous tile's second barrier: s_vc is no longer read)
            load_view_prefix(s_vc, p.views + view);
            cur_view = view;
        }
        if (threadIdx.x == 0) {  // next tile's ticket: its round trip overlaps this tile's work
            const uint32_t t = atomicAdd(p.tickets + 0, 1u);
            uint32_t v = vl;
            if (t < total)
                while (s_prefix[v + 1] <= t) v++;  // tickets grow monotonically within a block: amortised O(1)
            s_ticket[slot ^ 1u] = t;
            s_vl[slot ^ 1u] = v;
        }
        const ViewConst& vc = s_vc;
        // chunk = one 32x8 row-tile of a queued region; a warp covers an 8x4 patch of it
        const uint32_t c = g - s_prefix[vl];
        const uint32_t region = p.queue[(size_t)view * p.rqueue_cap + (c >> 2)];
        const int x0 = (int)((region & 0xFFFFu) << 5), y0 = (int)((region >> 16) << 5) + (int)((c & 3u) << 3);
        const int px = x0 + ((warp & 3) << 3) + (lane & 7);
        const int py = y0 + ((warp >> 2) << 2) + (lane >> 3);
        const uint32_t pid = ((uint32_t)py << 16) | (uint32_t)px;  // queues carry (y,x) packed: no division downstream
        bool active = px < p.GW && py < p.GH;
        if (MASKED && active) {
            const unsigned long long lin = (unsigned long long)py * p.GW + px;
            const uint32_t w = __ldg(p.mask + (size_t)view * p.mask_words + (uint32_t)(lin >> 5));
            active = (w >> (lin & 31)) & 1u;
        }
        bool keep = false;
        uint32_t cell = kNone;
        if (active) {
            if (!(vc.flags & kViewFastOk)) {
                keep = true;  // this view needs the literal march (max-range test): no cull
            } else {
                float dx, dy, dz;
                ray_direction_approx_px(p.cam, vc, px, py, dx, dy, dz);
                keep = !coarse_miss(p.map, vc, dx, dy, dz, cell);
            }
            if (!keep && p.pix_hit) {
                const size_t o = (size_t)view * p.pix_stride + (size_t)py * p.GW + px;
                p.pix_hit[o] = kNone;
                if (p.pix_depth) p.pix_depth[o] = 0.0f;
            }
        }
        // rays of the tile: dense mode by geometry (one atomic per tile), voxel mode by counting the masked pixels per warp
        if (MASKED) {
            const uint32_t nact = __popc(__ballot_sync(0xFFFFFFFFu, active));
            if (lane == 0 && nact) atomicAdd(p.stats + 4 * (size_t)view, (unsigned long long)nact);
        } else if (threadIdx.x == 32) {
            const int w = min(32, p.GW - x0), h = min(8, p.GH - y0);
            if (w > 0 && h > 0) atomicAdd(p.stats + 4 * (size_t)view, (unsigned long long)(w * h));
        }
        // survivors -> queue 2 (+ entry brick): one run per tile, warp after warp
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_wcnt[slot][warp] = (uint32_t)__popc(bal);
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the eight warp counts, one atomic for the tile
            const uint32_t n = lane < 8 ? s_wcnt[slot][lane] : 0u;
            uint32_t incl = n;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 7);
            uint32_t base = 0;
            if (lane == 0 && tot) base = atomicAdd(p.qcount2 + view, tot);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (lane < 8) s_base[slot][lane] = base + incl - n;
        }
        __syncthreads();
        if (keep) {
            const size_t pos = (size_t)view * p.queue_cap + s_base[slot][warp] + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            p.queue2[pos] = pid;
            if (p.queue2b) p.queue2b[pos] = cell;
        }
    }
}


template <int MINB, bool MASKED>
__global__ void __launch_bounds__(256, MINB) coarse_kernel(const CastParams p) {
    __shared__ uint32_t s_prefix[kMaxViewsPerLaunch + 1];
    build_chunk_prefix(p.qcount + p.view_base, p.nviews, s_prefix, 256u, 4u);
    const uint32_t total = s_prefix[p.nviews];
    if (total >= gridDim.x * 8u * kCoarseWarpTiles)
        coarse_tiles_by_warp<MASKED>(p, s_prefix, total);
    else
        coarse_tiles_by_block<MASKED>(p, s_prefix, total);
}

// Persistent WARPS: every warp pulls 32-ray chunks of the flattened (view, chunk) list with its own atomic ticket and
// shares nothing with the other warps of its block after the chunk-prefix table is built -- no block barrier in the loop
// (with 64-ray block chunks ncu showed ~11 % of warp time parked at the ticket barrier waiting for the sibling warp).
// The ticket for the NEXT chunk is requested before the current chunk is marched, so the ~1 us atomic round trip is
// hidden (un-prefetched warp tickets had measured 6 % slower than block tickets).  Measured on C2: 56 registers / 32 warps
// per SM (no spills) 0.514 ms, 48 / 40 (16 B spilled) 0.521 ms, 40 / 48 (88 B spilled) 0.530 ms; round 2 (profiles/
// r2_staging_ab.md): 9 blocks of 128 threads (36 warps, 56 registers, 16 B spilled) +1 %, two probes in flight instead of
// four +2 % on C2 / -0.5 % on C3.
constexpr int kMarchBlock = 256, kMarchMinBlocks = 4;
#ifndef PRV_MARCH_TICKET_MAX
#define PRV_MARCH_TICKET_MAX 8
#endif
constexpr int kMarchTicketMax = PRV_MARCH_TICKET_MAX;
template <int BS, int MINB, bool SMEM = false>
__global__ void __launch_bounds__(BS, MINB) march_kernel(const CastParams p) {
    __shared__ ViewConst s_vcw[BS / 32];
    __shared__ uint32_t s_prefix[kMaxViewsPerLaunch + 1];
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < p.map.pad_words; i += BS) s_dyn_pad[i] = __ldg(p.map.bitmap_pad + i);
        // (build_chunk_prefix's barriers order these stores before the first probe)
    }
    build_chunk_prefix(p.qcount2 + p.view_base, p.nviews, s_prefix, 32u);
    const uint32_t total = s_prefix[p.nviews];
    // A ticket is `per` consecutive chunks of the flattened list: 1 for small casts (C2: 23 chunks per warp, balance matters), up to
    // kMarchTicketMax when every warp still gets 32+ tickets -- on the 1024-view workload a warp that takes single chunks changes
    // view (constants reloaded, counters flushed) on four chunks out of five (C3 march 12.07 -> 11.66 ms; C2 0.424 -> 0.537 ms if
    // forced there, profiles/r2_march_ab.md).  Every block derives the same value from the same table.
    // (kept in shared memory and re-read per ticket: two more live registers put MOVs back into the four-probe loop)
    __shared__ uint32_t s_per;
    if (threadIdx.x == 0) s_per = min((uint32_t)kMarchTicketMax, max(1u, total / (gridDim.x * (uint32_t)(BS / 32) * kTicketSpread)));
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ViewConst& vc = s_vcw[warp];
    uint32_t cur_view = 0xFFFFFFFFu;
    uint32_t c_probes = 0, c_hits = 0, c_steps = 0;
    uint32_t vl = 0;
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(p.tickets + 1, 1u);
    next = __shfl_sync(0xFFFFFFFFu, next, 0);
    for (;;) {
        const uint32_t per = *reinterpret_cast<volatile uint32_t*>(&s_per);
        uint32_t g = next * per;
        if (g >= s_prefix[p.nviews]) break;
        const uint32_t g_end = min(s_prefix[p.nviews], g + per);
        if (lane == 0) next = atomicAdd(p.tickets + 1, 1u);  // consumed after this ticket: the round trip overlaps the march
#pragma unroll 1
        for (; g < g_end; g++) {
        while (s_prefix[vl + 1] <= g) vl++;                  // a warp's chunks grow monotonically: amortised O(1)
        const uint32_t view = vl + p.view_base;
        if (view != cur_view) {
            if (cur_view != 0xFFFFFFFFu) {  // flush the finished view's counters: one atomic per counter per (warp, view)
                warp_commit_stats(p.stats + 4 * (size_t)cur_view, c_probes, c_hits, c_steps);
                c_probes = c_hits = c_steps = 0;
            }
            __syncwarp();
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(p.views + view);
            uint32_t* d32 = reinterpret_cast<uint32_t*>(&s_vcw[warp]);
            for (int i = lane; i < (int)(sizeof(ViewConst) / 4); i += 32) d32[i] = s32[i];
            __syncwarp();
            cur_view = view;
        }
        const uint32_t count = p.qcount2[view];
        const uint32_t idx = (g - s_prefix[vl]) * 32u + (uint32_t)lane;
        if (idx < count) {
            const uint32_t packed = p.queue2[(size_t)view * p.queue_cap + idx];
            uint32_t cell = p.queue2b ? p.queue2b[(size_t)view * p.queue_cap + idx] : kNone;  // brick entry (prv_set_brick_cull)
            const int py = (int)(packed >> 16), px = (int)(packed & 0xFFFFu);
            const uint32_t pid = (uint32_t)py * (uint32_t)p.GW + (uint32_t)px;
            CastResult res;
            res.rank = kNone;
            res.steps = 0;
            res.probes = 0;
            res.k0 = res.k1 = res.k2 = 0;
#pragma unroll 1
            for (;;) {  // a second pass only when a brick box could not be entered (never observed): start over from the AABB face
                RayState r;
                float dx, dy, dz;
                ray_direction(p.cam, vc, px, py, dx, dy, dz);
                if (!ray_init(vc, p.map.resolution, dx, dy, dz, r)) break;
                if (!(vc.flags & kViewFastOk)) {
                    march_plain(p.map, p.cam, vc, r, res);
                    break;
                }
                if (march_axis<SMEM>(p.map, vc, r, cell, res)) break;
                cell = kNone;
            }
            write_hit(p, vc, view, pid, res);
            c_probes += res.probes;
            c_hits += res.rank != kNone ? 1u : 0u;
            c_steps += res.steps;
        }
        }
        next = __shfl_sync(0xFFFFFFFFu, next, 0);
    }
    if (cur_view != 0xFFFFFFFFu) warp_commit_stats(p.stats + 4 * (size_t)cur_view, c_probes, c_hits, c_steps);
}

// ---- PLAIN / FAST variants: one kernel, one thread per pixel of a 32x8 tile ----------------------------------------
template <int VARIANT, bool MASKED>
__global__ void __launch_bounds__(256) raycast_kernel(const CastParams p) {
    __shared__ ViewConst s_vc;
    const uint32_t view = blockIdx.y + p.view_base;
    load_view_const(s_vc, p.views + view);
    const ViewConst& vc = s_vc;
    const int tiles_x = (p.GW + 31) >> 5;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = (tile_x << 5) + ((warp & 3) << 3) + (lane & 7);
    const int py = (tile_y << 3) + ((warp >> 2) << 2) + (lane >> 3);
    const bool in_grid = px < p.GW && py < p.GH;
    const unsigned long long pid = (unsigned long long)py * p.GW + px;
    bool active = in_grid && (vc.flags & kViewInMap) && !(vc.flags & kViewInObject);
    if (MASKED && active) {
        const uint32_t w = __ldg(p.mask + (size_t)view * p.mask_words + (uint32_t)(pid >> 5));
        active = (w >> (pid & 31)) & 1u;
    }
    CastResult res;
    res.rank = kNone;
    res.steps = 0;
    res.probes = 0;
    res.k0 = res.k1 = res.k2 = 0;
    if (active) {
        RayState r;
        const bool plain = VARIANT == PRV_VARIANT_PLAIN || !(vc.flags & kViewFastOk);
        if (setup_ray(p.cam, vc, p.map.resolution, px, py, r)) {
            if (plain)
                march_plain(p.map, p.cam, vc, r, res);
            else
                march_fast(p.map, vc, r, res);
        }
    }
    if (in_grid && (!MASKED || active)) write_hit(p, vc, view, pid, res);
    commit_stats(p.stats + 4 * (size_t)view, active ? 1u : 0u, res.probes, res.rank != kNone ? 1u : 0u, res.steps);
}

// rs2_deproject_pixel_to_point of every integer pixel of the (W+1) x (H+1) grid at depth 1 (models 0 / 2 / 4), with the very
// function the march would otherwise evaluate per ray: DevCam::deproj_exact
__global__ void __launch_bounds__(256) deproj_table_kernel(DevCam cam, float2* out) {
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px > cam.W || py > cam.H) return;
    float x, y;
    deproject_pixel(cam, (float)px, (float)py, x, y);
    out[(size_t)py * (size_t)(cam.W + 1) + (size_t)px] = make_float2(x, y);
}

// voxel-driven mode, stage 1 (main.cpp:243-251 of the reference): project every occupied voxel centre, mark its
// TRUNCATED pixel in the (W+1)x(H+1) mask (pixel == W or == H passes the reference's '>' test).
__global__ void __launch_bounds__(256) project_voxels_kernel(DevMap map, DevCam cam, const ViewConst* views, uint32_t view_base,
                                                             uint32_t* mask, uint32_t mask_words, uint32_t* voxel_pix) {
    const uint32_t view = blockIdx.y + view_base;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= map.n_occ) return;
    const ViewConst& vc = views[view];
    uint32_t pid = kNone;
    if ((vc.flags & kViewInMap) && !(vc.flags & kViewInObject)) {
        const float ex = (float)key_to_coord_d(map.keys[3 * i + 0], map.resolution);
        const float ey = (float)key_to_coord_d(map.keys[3 * i + 1], map.resolution);
        const float ez = (float)key_to_coord_d(map.keys[3 * i + 2], map.resolution);
        const float vx = (float)row_apply(vc.inv + 0, (double)ex, (double)ey, (double)ez);
        const float vy = (float)row_apply(vc.inv + 4, (double)ex, (double)ey, (double)ez);
        const float vz = (float)row_apply(vc.inv + 8, (double)ex, (double)ey, (double)ez);
        float u, v;
        project_point_to_pixel(cam, vx, vy, vz, u, v);
        // reject: pixel<0 || pixel>W (resp. H); NaN is rejected too (float->int of NaN is UB in the reference)
        if (u >= 0.0f && u <= (float)cam.W && v >= 0.0f && v <= (float)cam.H) {
            const int ix = (int)u, iy = (int)v;  // truncation at the int-parameter call, main.cpp:253
            pid = (uint32_t)iy * (uint32_t)(cam.W + 1) + (uint32_t)ix;
            atomicOr(mask + (size_t)view * mask_words + (pid >> 5), 1u << (pid & 31));
        }
    }
    voxel_pix[(size_t)view * map.n_occ + i] = pid;
}

// voxel-driven mode, stage 3: voxel i takes the result of its pixel's ray
__global__ void __launch_bounds__(256) gather_voxel_hits_kernel(uint32_t n_occ, const uint32_t* voxel_pix, const uint32_t* pix_hit,
                                                                unsigned long long pix_stride, uint32_t* out) {
    const uint32_t view = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_occ) return;
    const uint32_t pid = voxel_pix[(size_t)view * n_occ + i];
    out[(size_t)view * n_occ + i] = pid == kNone ? kNone : pix_hit[(size_t)view * pix_stride + pid];
}

// cloud->points image of Perception_3D::precept for one view (main.cpp:240-283)
__global__ void __launch_bounds__(256) precept_points_kernel(DevMap map, const uint32_t* voxel_hit, prv_point_xyzrgb* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= map.n_occ) return;
    prv_point_xyzrgb pt;
    pt.x = pt.y = pt.z = 0.0f;
    pt.w = 1.0f;
    pt.b = pt.g = pt.r = 0;
    pt.a = 255;
    pt.pad[0] = pt.pad[1] = pt.pad[2] = 0.0f;
    const uint32_t h = voxel_hit[i];
    if (h != kNone) {
        pt.x = (float)key_to_coord_d(map.keys[3 * h + 0], map.resolution);
        pt.y = (float)key_to_coord_d(map.keys[3 * h + 1], map.resolution);
        pt.z = (float)key_to_coord_d(map.keys[3 * h + 2], map.resolution);
        pt.r = map.rgb[3 * h + 0];
        pt.g = map.rgb[3 * h + 1];
        pt.b = map.rgb[3 * h + 2];
    }
    out[i] = pt;
}
