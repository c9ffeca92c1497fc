// prv_device.cu -- device side of include/prv.h: context, HBM layout, kernel launches, NCCL glue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off (see build.py).
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/prv.h"
#include "../host/prv_linalg.hpp"
#include "prv_kernels.cuh"
#include "prv_view_const.hpp"

using namespace prvk;

// =====================================================================================================
// kernels
// =====================================================================================================

#include "kernels_common.cuh"
#include "kernels_cast.cuh"
#include "kernels_map.cuh"
#include "kernels_greedy.cuh"
#include "kernels_ensemble.cuh"
#include "kernels_splat.cuh"

// =====================================================================================================
// context
// =====================================================================================================
namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

enum KClass { K_CAST = 0, K_CULL, K_MARCH, K_PROJECT, K_COUNT, K_GREEDY, K_SPLAT, K_RESOLVE, K_OTHER, K_GATHER, K_FLUSH, K_NCLASS };

struct TimedSpan {
    int cls;
    uint32_t launches;
    cudaEvent_t a, b;
};

struct Id128 {
    char b[128];
};

// NCCL is resolved at run time (dlopen) so the library has no link-time dependency on it.
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

}  // namespace

struct prv_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t mem_bytes = 0;
    // Two streams: `stream` carries uploads, map build, the cast and the splat; `score_stream` carries what consumes the
    // coverage rows of a finished cast -- the all-gather and the greedy selection -- so that the scoring of cast k overlaps
    // the cull / march of cast k+1 when the caller enqueues several steps without synchronising in between.  The rows are
    // double-buffered (d_rows[2]); events order producer and consumer (rows_ready / rows_free).
    cudaStream_t stream = nullptr;
    cudaStream_t score_stream = nullptr;
    cudaEvent_t ev_rows_ready[2] = {nullptr, nullptr}, ev_rows_free[2] = {nullptr, nullptr}, ev_join = nullptr, ev_flush = nullptr;
    bool rows_free_pending[2] = {false, false};
    int cur = 0;  // row buffer of the latest cast
    char err[512] = {0};
    int variant = PRV_VARIANT_AXIS;
    bool last_cast_axis = false;  // the latest cast ran the AXIS pipeline (queues / qcount valid for prv_get_cast_stats)
    uint32_t last_cast_views = 0;
    int occ_coarse = 0, occ_march = 0, occ_greedy = 0;
    uint32_t occ_greedy_words = 0;
    bool greedy_persistent = false;
    int greedy_path = 0;  // 0 cluster kernel, 1 grid-barrier kernel, 2 one launch per iteration (prv_greedy_path)
    // PRV_GREEDY_CLUSTER=0 forces the grid-barrier kernel (the fallback for tables larger than one cluster's shared memory)
    bool greedy_no_cluster = false;
    int greedy_blocks_per_sm = 2;
    DevBuf d_arrive;
    DevBuf d_ens_images, d_ens_terms, d_ens_scores;
    DevBuf d_ing_xyz, d_ing_rgb, d_ing_k0, d_ing_k1, d_ing_v0, d_ing_v1, d_ing_pos, d_ing_tmp;

    // map
    bool have_map = false;
    DevMap map{};
    double resolution = 0;
    int lo[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    DevBuf d_coarse;
    // A/B switches for the north-star's staging claims (prv_set_staging): padded bitmap in shared memory; L2 persisting window
    int stage_smem = 1, stage_l2 = 0;
    int occ_march_smem = 0;
    uint32_t occ_march_smem_words = 0;
    int brick_cs = 0;  // brick edge requested for the NEXT prv_set_map (0 = by map size); ctx->map.cs is what the resident map carries
    bool brick_entry = true;        // start the exact march at the first set brick of the coarse walk (prv_set_brick_cull)
    DevBuf d_bitmap, d_bitmap_pad, d_prefix, d_leaf_of_raster, d_keys, d_rgb, d_tilesum, d_check, d_deproj;
    std::vector<uint16_t> h_keys;

    // camera
    bool have_cam = false;
    DevCam cam{};
    prv_intrinsics intr{};
    prv_intrinsics intr_checked{};  // intrinsics the region-cull validation was last run for

    // views
    uint32_t V = 0;
    DevBuf d_views, d_view_ids, d_row_of_id;
    std::vector<ViewConst> h_views;
    std::vector<uint32_t> h_view_ids;
    bool ids_user = false;                  // h_view_ids came from prv_set_view_ids (else 0..V-1, regenerated by every prv_set_views)
    std::vector<uint32_t> h_row_of_id;      // view id -> local row (kNone = absent)
    std::vector<uint32_t> h_row_of_id_all;  // view id -> row of the all-gathered table
    DevBuf d_row_of_id_all;
    bool gather_ids_valid = false;
    uint32_t id_space = 0;

    // cast outputs
    DevBuf d_queue, d_queue2, d_queue2b;
    DevBuf d_zero;  // stats [V][4] u64 | qcount [2][V] u32 | tickets: everything a cast needs zeroed, one memset
    DevBuf d_rows[2], d_counts, d_pix_hit, d_pix_depth, d_mask, d_voxel_pix, d_voxel_hit, d_points;
    unsigned long long* p_stats = nullptr;
    uint32_t* p_qcount = nullptr;
    uint32_t* p_tickets = nullptr;
    int last_mode = -1;
    bool have_pixels = false;
    bool cast_done = false;
    unsigned long long pix_stride = 0;
    uint32_t mask_words = 0;

    // greedy
    DevBuf d_best, d_cov[2];
    uint32_t greedy_max_iter = 0;
    bool greedy_done = false;
    // table the greedy runs over (local rows, or the all-gathered table)
    const uint64_t* g_rows = nullptr;
    const uint32_t* g_ids = nullptr;
    uint32_t g_nrows = 0;
    DevBuf d_all_rows, d_all_ids;
    bool gathered = false;

    // splat
    DevBuf d_cloud_xyz, d_cloud_rgb, d_corner, d_rgba, d_depth_img;
    uint64_t P = 0;
    uint32_t rendered_views = 0;

    // timing
    uint32_t spans_dropped = 0;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> free_events;
    cudaEvent_t slots[16] = {};
    DevBuf d_flush;
    // pinned staging for small device->host results and the event that marks "host inputs consumed"
    void* h_stage = nullptr;
    size_t h_stage_cap = 0;
    cudaEvent_t ev_copy = nullptr;

    // counters (since prv_reset_counters)
    uint64_t n_launches = 0, h2d_bytes = 0, d2h_bytes = 0;

    // peer-memory exchange of the coverage rows (prv_comm_p2p_*): arena = flags | error | 4 x {rows, ids}
    DevBuf d_arena;
    void* peer_base[kMaxPeers] = {};
    bool peer_opened[kMaxPeers] = {};
    bool p2p_ready = false;
    size_t arena_bytes = 0, arena_buf_bytes = 0;
    uint32_t p2p_step = 0;          // published casts so far
    bool cur_published = false;     // the resident cast published its rows (prv_cast_async with PRV_CAST_PUBLISH)
    DevBuf d_done, d_counts_scratch;  // d_done[0]: block counter of the cast stream's count kernel, [1]: of the scoring stream's

    // nccl
    NcclApi nccl;
    void* comm = nullptr;
    int rank = 0, nranks = 1;
};

namespace {

char g_create_error[512] = {0};

int fail(prv_ctx* c, int code, const char* fmt, ...) {  // (fixed buffers: reporting an allocation failure must not allocate)
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

// nothing may throw across the C ABI: bodies that allocate host memory run inside these
#define PRV_GUARD_BEGIN try {
#define PRV_GUARD_END(ctx_, what_)                                                                  \
    }                                                                                               \
    catch (const std::bad_alloc&) { return fail(ctx_, PRV_ERR_OOM, "%s: host allocation failed", what_); } \
    catch (...) { return fail(ctx_, PRV_ERR_INVALID, "%s: unexpected exception", what_); }

#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? PRV_ERR_OOM : PRV_ERR_CUDA, \
                                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int ensure(prv_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return PRV_OK;
    if (b.p) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaStreamSynchronize(ctx->score_stream));
        CU(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 255) & ~(size_t)255;
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return PRV_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

cudaEvent_t get_event(prv_ctx* ctx) {
    if (!ctx->free_events.empty()) {
        cudaEvent_t e = ctx->free_events.back();
        ctx->free_events.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

constexpr size_t kMaxSpans = 65536;  // prv_get_timing reports how many spans were not recorded beyond this (prv_timing::dropped)
struct Span {
    prv_ctx* ctx;
    size_t idx;
    bool on;
    cudaStream_t st;
    Span(prv_ctx* c, int cls, uint32_t launches, cudaStream_t stream = nullptr) : ctx(c), idx(0), on(c->spans.size() < kMaxSpans), st(stream ? stream : c->stream) {
        c->n_launches += launches;
        if (!on) {
            c->spans_dropped++;
            return;
        }
        TimedSpan s;
        s.cls = cls;
        s.launches = launches;
        s.a = get_event(c);
        s.b = get_event(c);
        cudaEventRecord(s.a, st);
        idx = c->spans.size();
        c->spans.push_back(s);
    }
    ~Span() {
        if (on) cudaEventRecord(ctx->spans[idx].b, st);
    }
};

// everything queued on the scoring stream so far happens before whatever is queued on the main stream from here on
cudaError_t join_score(prv_ctx* ctx) {
    cudaError_t e = cudaEventRecord(ctx->ev_join, ctx->score_stream);
    if (e != cudaSuccess) return e;
    return cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
}
cudaError_t sync_all(prv_ctx* ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ctx->score_stream);
}

cudaError_t h2d(prv_ctx* ctx, void* dst, const void* src, size_t bytes) {
    ctx->h2d_bytes += bytes;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
}
cudaError_t d2h(prv_ctx* ctx, void* dst, const void* src, size_t bytes) {
    ctx->d2h_bytes += bytes;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
}

// Small results (coverage rows, counts, the greedy sequence) are copied into a ctx-owned pinned buffer and memcpy'd to
// the caller: a device->host copy straight into pageable memory is staged by the driver piecewise and costs ~0.15 ms of
// latency for 0.5 MB.  Pinned / registered destinations and large results are copied directly.  Synchronous.
constexpr size_t kStageMax = (size_t)8 << 20;
cudaError_t d2h_result(prv_ctx* ctx, void* dst, const void* src, size_t bytes) {
    cudaError_t e;
    bool direct = bytes > kStageMax;
    if (!direct) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, dst) == cudaSuccess)
            direct = attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
        else
            cudaGetLastError();
    }
    if (!direct && bytes > ctx->h_stage_cap) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_cap = 0;
        const size_t cap = std::max<size_t>((bytes + 4095) & ~(size_t)4095, (size_t)1 << 20);
        if (cudaMallocHost(&ctx->h_stage, cap) == cudaSuccess)
            ctx->h_stage_cap = cap;
        else {
            cudaGetLastError();
            direct = true;
        }
    }
    if (direct) {
        if ((e = d2h(ctx, dst, src, bytes)) != cudaSuccess) return e;
        return cudaStreamSynchronize(ctx->stream);
    }
    if ((e = d2h(ctx, ctx->h_stage, src, bytes)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return e;
    memcpy(dst, ctx->h_stage, bytes);
    return cudaSuccess;
}

template <typename T>
T* ptr(const DevBuf& b) {
    return reinterpret_cast<T*>(b.p);
}

constexpr uint32_t kMaxViewId = 1u << 24;

uint32_t words_for(uint32_t n_occ) {
    uint32_t w = (n_occ + 63) / 64;
    if (w == 0) w = 1;
    return (w + 1) & ~1u;
}

// per-view constants (prv_view_const.hpp) for the resident map and camera
void make_view_const(const prv_ctx* ctx, const double* pose_world, const double* init_pos, uint32_t id, ViewConst& vc) {
    ViewSetup s;
    s.resolution = ctx->resolution;
    for (int a = 0; a < 3; a++) {
        s.lo[a] = ctx->lo[a];
        s.n[a] = ctx->n[a];
    }
    s.keys = ctx->h_keys.data();
    s.n_keys = ctx->h_keys.size() / 3;
    s.max_range = ctx->cam.max_range;
    s.max_range_sq = ctx->cam.max_range_sq;
    prvk::make_view_const(s, pose_world, init_pos, id, vc);
}

int upload_views(prv_ctx* ctx, const double* pose_world, const double* init_pos, uint32_t V) {
    if (!ctx->have_map || !ctx->have_cam) return fail(ctx, PRV_ERR_INVALID, "prv_set_views: call prv_set_map and prv_set_camera first");
    if (!pose_world || !init_pos || V == 0) return fail(ctx, PRV_ERR_INVALID, "prv_set_views: null/empty input");
    if (ctx->ids_user && ctx->h_view_ids.size() != V)
        return fail(ctx, PRV_ERR_INVALID, "prv_set_views: %u views but prv_set_view_ids gave %zu ids (pass the same V, or clear them with prv_set_view_ids(ctx, NULL, 0))",
                    V, ctx->h_view_ids.size());
    CU(join_score(ctx));  // a selection still running on the scoring stream reads the id tables this call replaces
    PRV_GUARD_BEGIN
    ctx->h_views.resize(V);
    const bool ids_ok = ctx->ids_user && ctx->h_view_ids.size() == V;
    for (uint32_t v = 0; v < V; v++) make_view_const(ctx, pose_world + 16 * (size_t)v, init_pos + 3 * (size_t)v, ids_ok ? ctx->h_view_ids[v] : v, ctx->h_views[v]);
    if (!ids_ok) {
        ctx->h_view_ids.resize(V);
        for (uint32_t v = 0; v < V; v++) ctx->h_view_ids[v] = v;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->d_views, sizeof(ViewConst) * (size_t)V))) return rc;
    CU(h2d(ctx, ctx->d_views.p, ctx->h_views.data(), sizeof(ViewConst) * (size_t)V));
    // id tables for the greedy
    uint32_t max_id = 0;
    for (uint32_t v = 0; v < V; v++) max_id = std::max(max_id, ctx->h_view_ids[v]);
    std::vector<uint32_t>& row_of_id = ctx->h_row_of_id;
    row_of_id.assign((size_t)max_id + 1, kNone);
    for (uint32_t v = 0; v < V; v++) row_of_id[ctx->h_view_ids[v]] = v;
    ctx->id_space = max_id + 1;
    if ((rc = ensure(ctx, ctx->d_view_ids, 4 * (size_t)V))) return rc;
    if ((rc = ensure(ctx, ctx->d_row_of_id, 4 * (size_t)ctx->id_space))) return rc;
    CU(h2d(ctx, ctx->d_view_ids.p, ctx->h_view_ids.data(), 4 * (size_t)V));
    CU(h2d(ctx, ctx->d_row_of_id.p, row_of_id.data(), 4 * (size_t)ctx->id_space));
    // no stream sync: the three sources are pageable ctx-owned vectors, which cudaMemcpyAsync consumes (stages) before it
    // returns, so the copies simply queue behind whatever the stream is still doing (e.g. prv_set_map's table kernels)
    ctx->V = V;
    ctx->cast_done = false;
    ctx->greedy_done = false;
    ctx->gathered = false;
    ctx->gather_ids_valid = false;
    PRV_GUARD_END(ctx, "prv_set_views")
    return PRV_OK;
}

// ---- peer arena (prv_comm_p2p_*): header (flags, error), then kPeerBuffers x { rows [G*V][words] u64 | ids [G*V] u32 }
constexpr size_t kArenaHeader = 4096;
uint32_t* arena_flags(void* base) { return reinterpret_cast<uint32_t*>(base); }
uint32_t* arena_error(void* base) { return reinterpret_cast<uint32_t*>(base) + 16; }
bool peer_arena_for(prv_ctx* ctx, uint32_t step, uint32_t V, uint32_t words, PeerArena& pa) {
    const size_t G = (size_t)ctx->nranks;
    const size_t rows_bytes = G * V * words * 8, ids_bytes = G * V * 4;
    if (rows_bytes + ids_bytes > ctx->arena_buf_bytes) return false;
    const size_t off = kArenaHeader + (size_t)(step % kPeerBuffers) * ctx->arena_buf_bytes;
    for (int r = 0; r < ctx->nranks; r++) {
        char* base = reinterpret_cast<char*>(r == ctx->rank ? ctx->d_arena.p : ctx->peer_base[r]);
        pa.flags[r] = arena_flags(base);
        pa.rows[r] = reinterpret_cast<uint64_t*>(base + off);
        pa.ids[r] = reinterpret_cast<uint32_t*>(base + off + rows_bytes);
    }
    pa.nranks = (uint32_t)ctx->nranks;
    pa.rank = (uint32_t)ctx->rank;
    return true;
}

template <int VARIANT>
void launch_cast(prv_ctx* ctx, const CastParams& p, bool masked, dim3 grid) {
    if (masked)
        raycast_kernel<VARIANT, true><<<grid, 256, 0, ctx->stream>>>(p);
    else
        raycast_kernel<VARIANT, false><<<grid, 256, 0, ctx->stream>>>(p);
}

int cast_impl(prv_ctx* ctx, int mode, int want_pixels, bool publish = false) {
    if (!ctx->have_map || !ctx->have_cam || ctx->V == 0) return fail(ctx, PRV_ERR_INVALID, "prv_cast: map, camera and views must be set");
    if (mode != PRV_MODE_VOXEL && mode != PRV_MODE_DENSE) return fail(ctx, PRV_ERR_INVALID, "prv_cast: bad mode %d", mode);
    const uint32_t V = ctx->V;
    const uint32_t words = ctx->map.words64;
    const int W = ctx->cam.W, H = ctx->cam.H;
    int rc;
    // the other row buffer: the scoring stream may still be reading the rows of the previous cast
    const int cur = ctx->cur ^ 1;
    if ((rc = ensure(ctx, ctx->d_rows[cur], (size_t)V * words * 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_counts, (size_t)V * 4))) return rc;
    const uint32_t nlaunch_axis = (V + kMaxViewsPerLaunch - 1) / kMaxViewsPerLaunch;
    const size_t zero_bytes = (size_t)V * 4 * 8 + (size_t)V * 2 * 4 + (size_t)nlaunch_axis * 2 * 4;
    if ((rc = ensure(ctx, ctx->d_zero, zero_bytes))) return rc;
    ctx->p_stats = ptr<unsigned long long>(ctx->d_zero);
    ctx->p_qcount = reinterpret_cast<uint32_t*>(ctx->p_stats + (size_t)V * 4);
    ctx->p_tickets = ctx->p_qcount + (size_t)V * 2;
    if (ctx->rows_free_pending[cur]) {  // ... and the cast before that must have been scored
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_rows_free[cur], 0));
        ctx->rows_free_pending[cur] = false;
    }
    CU(cudaMemsetAsync(ctx->d_rows[cur].p, 0, (size_t)V * words * 8, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_zero.p, 0, zero_bytes, ctx->stream));

    CastParams p{};
    p.map = ctx->map;
    p.cam = ctx->cam;
    p.views = ptr<ViewConst>(ctx->d_views);
    p.bitsets32 = ptr<uint32_t>(ctx->d_rows[cur]);
    p.stats = ctx->p_stats;
    const bool voxel = mode == PRV_MODE_VOXEL;
    p.GW = voxel ? W + 1 : W;
    p.GH = voxel ? H + 1 : H;
    ctx->pix_stride = (unsigned long long)p.GW * p.GH;
    p.pix_stride = ctx->pix_stride;
    const bool pixels = voxel || want_pixels;
    if (pixels) {
        if ((rc = ensure(ctx, ctx->d_pix_hit, (size_t)V * ctx->pix_stride * 4))) return rc;
        p.pix_hit = ptr<uint32_t>(ctx->d_pix_hit);
        if (!voxel) {
            if ((rc = ensure(ctx, ctx->d_pix_depth, (size_t)V * ctx->pix_stride * 4))) return rc;
            p.pix_depth = ptr<float>(ctx->d_pix_depth);
        }
    }
    if (voxel) {
        ctx->mask_words = (uint32_t)((ctx->pix_stride + 31) / 32);
        p.mask_words = ctx->mask_words;
        if ((rc = ensure(ctx, ctx->d_mask, (size_t)V * ctx->mask_words * 4))) return rc;
        if ((rc = ensure(ctx, ctx->d_voxel_pix, (size_t)V * ctx->map.n_occ * 4))) return rc;
        CU(cudaMemsetAsync(ctx->d_mask.p, 0, (size_t)V * ctx->mask_words * 4, ctx->stream));
        p.mask = ptr<uint32_t>(ctx->d_mask);
    }
    if (ctx->variant == PRV_VARIANT_AXIS) {
        const uint32_t nregions = (uint32_t)(((p.GW + 31) / 32) * ((p.GH + 31) / 32));
        if ((rc = ensure(ctx, ctx->d_queue, (size_t)V * nregions * 4))) return rc;
        if ((rc = ensure(ctx, ctx->d_queue2, (size_t)V * ctx->pix_stride * 4))) return rc;
        p.rqueue_cap = nregions;
        p.queue = ptr<uint32_t>(ctx->d_queue);
        p.qcount = ctx->p_qcount;
        p.queue2 = ptr<uint32_t>(ctx->d_queue2);
        p.qcount2 = ctx->p_qcount + V;
        p.queue_cap = ctx->pix_stride;
        if (ctx->brick_entry) {  // (a grid too large to pack its brick coordinates reports no brick: coarse_miss)
            if ((rc = ensure(ctx, ctx->d_queue2b, (size_t)V * ctx->pix_stride * 4))) return rc;
            p.queue2b = ptr<uint32_t>(ctx->d_queue2b);
        }
    }
    const uint32_t tiles = (uint32_t)(((p.GW + 31) / 32) * ((p.GH + 7) / 8));
    const uint32_t vstep = ctx->variant == PRV_VARIANT_AXIS ? (uint32_t)kMaxViewsPerLaunch : 32768u;
    for (uint32_t vb = 0, li = 0; vb < V; vb += vstep, li++) {
        const uint32_t vn = std::min<uint32_t>(vstep, V - vb);
        if (voxel && ctx->cam.deproj_table) {
            // models 3 / 5: main.cpp:243-251 on the host (atan is the host libm's), then mask + pixel table uploaded
            PRV_GUARD_BEGIN
            const uint32_t N = ctx->map.n_occ;
            std::vector<uint32_t> h_mask((size_t)vn * ctx->mask_words, 0u), h_pix((size_t)vn * N, kNone);
            for (uint32_t v = 0; v < vn; v++) {
                const ViewConst& vc = ctx->h_views[vb + v];
                if (!(vc.flags & kViewInMap) || (vc.flags & kViewInObject)) continue;
                for (uint32_t i = 0; i < N; i++) {
                    float e[3];
                    for (int a = 0; a < 3; a++) e[a] = (float)prv::key_to_coord(ctx->h_keys[3 * (size_t)i + a], ctx->resolution);
                    float pc[3];
                    for (int r = 0; r < 3; r++) {  // row_apply: m0*x; m1*y + acc; m2*z + acc; m3 + acc
                        double acc = vc.inv[4 * r + 0] * (double)e[0];
                        acc = vc.inv[4 * r + 1] * (double)e[1] + acc;
                        acc = vc.inv[4 * r + 2] * (double)e[2] + acc;
                        acc = vc.inv[4 * r + 3] + acc;
                        pc[r] = (float)acc;
                    }
                    float uv[2];
                    prv_host_project_point_to_pixel(&ctx->intr, pc, uv);
                    if (uv[0] >= 0.0f && uv[0] <= (float)W && uv[1] >= 0.0f && uv[1] <= (float)H) {
                        const uint32_t pid = (uint32_t)(int)uv[1] * (uint32_t)(W + 1) + (uint32_t)(int)uv[0];
                        h_mask[(size_t)v * ctx->mask_words + (pid >> 5)] |= 1u << (pid & 31);
                        h_pix[(size_t)v * N + i] = pid;
                    }
                }
            }
            CU(h2d(ctx, ptr<uint32_t>(ctx->d_mask) + (size_t)vb * ctx->mask_words, h_mask.data(), h_mask.size() * 4));
            CU(h2d(ctx, ptr<uint32_t>(ctx->d_voxel_pix) + (size_t)vb * N, h_pix.data(), h_pix.size() * 4));
            CU(cudaStreamSynchronize(ctx->stream));
            PRV_GUARD_END(ctx, "prv_cast (voxel mode, model 3 / 5)")
        } else if (voxel) {
            Span s(ctx, K_PROJECT, 1);
            project_voxels_kernel<<<dim3((ctx->map.n_occ + 255) / 256, vn), 256, 0, ctx->stream>>>(
                ctx->map, ctx->cam, ptr<ViewConst>(ctx->d_views), vb, ptr<uint32_t>(ctx->d_mask), ctx->mask_words, ptr<uint32_t>(ctx->d_voxel_pix));
        }
        p.view_base = vb;
        p.nviews = vn;
        const dim3 grid(tiles, vn);
        if (ctx->variant == PRV_VARIANT_AXIS) {
            p.tickets = ctx->p_tickets + 2 * li;
            if (ctx->occ_coarse == 0) {  // persistent grids = exactly one resident wave
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_coarse, coarse_kernel<kCoarseMinBlocks, false>, 256, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_march, march_kernel<kMarchBlock, kMarchMinBlocks>, kMarchBlock, 0);
                ctx->occ_coarse = std::max(1, ctx->occ_coarse);
                ctx->occ_march = std::max(1, ctx->occ_march);
            }
            {
                Span s(ctx, K_CULL, 2);
                const dim3 rgrid((uint32_t)(((p.GW + 31) / 32 + kCullRegions - 1) / kCullRegions), (uint32_t)((p.GH + 31) / 32), vn);
                if (voxel)
                    cull_kernel<true><<<rgrid, 256, 0, ctx->stream>>>(p);
                else
                    cull_kernel<false><<<rgrid, 256, 0, ctx->stream>>>(p);
                const uint32_t cgrid = (uint32_t)(ctx->sm_count * ctx->occ_coarse);
                if (voxel)
                    coarse_kernel<kCoarseMinBlocks, true><<<cgrid, 256, 0, ctx->stream>>>(p);
                else  // (8 blocks/SM at 32 registers with 28 B spilled measured 1-3 % faster than 6 blocks at 38 without)
                    coarse_kernel<kCoarseMinBlocks, false><<<cgrid, 256, 0, ctx->stream>>>(p);
            }
            Span s(ctx, K_MARCH, 1);
            const size_t pad_bytes = (size_t)p.map.pad_words * 4;
            bool smem_map = ctx->stage_smem && pad_bytes <= (size_t)44 * 1024;  // 4 blocks/SM stay resident (10.8 KB static each)
            if (smem_map && (ctx->occ_march_smem == 0 || ctx->occ_march_smem_words != p.map.pad_words)) {
                if (cudaFuncSetAttribute(march_kernel<kMarchBlock, kMarchMinBlocks, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad_bytes) != cudaSuccess) {
                    cudaGetLastError();
                    smem_map = false;
                } else {
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_march_smem, march_kernel<kMarchBlock, kMarchMinBlocks, true>, kMarchBlock, pad_bytes);
                    ctx->occ_march_smem = std::max(1, ctx->occ_march_smem);
                    ctx->occ_march_smem_words = p.map.pad_words;
                }
            }
            if (smem_map)
                march_kernel<kMarchBlock, kMarchMinBlocks, true><<<(uint32_t)(ctx->sm_count * ctx->occ_march_smem), kMarchBlock, pad_bytes, ctx->stream>>>(p);
            else
                march_kernel<kMarchBlock, kMarchMinBlocks><<<(uint32_t)(ctx->sm_count * ctx->occ_march), kMarchBlock, 0, ctx->stream>>>(p);
        } else {
            Span s(ctx, K_CAST, 1);
            if (ctx->variant == PRV_VARIANT_PLAIN)
                launch_cast<PRV_VARIANT_PLAIN>(ctx, p, voxel, grid);
            else
                launch_cast<PRV_VARIANT_FAST>(ctx, p, voxel, grid);
        }
    }
    ctx->cur_published = false;
    if (publish) {
        // multi-GPU: the count pass also stores every row into every rank's gathered table over NVLink peer memory
        if (!ctx->p2p_ready) return fail(ctx, PRV_ERR_INVALID, "prv_cast_async: PRV_CAST_PUBLISH needs prv_comm_p2p_import first");
        PeerArena pa;
        if (!peer_arena_for(ctx, ctx->p2p_step + 1, V, words, pa))
            return fail(ctx, PRV_ERR_INVALID, "prv_cast_async: the gathered table (%d x %u rows of %u words) does not fit the peer arena", ctx->nranks, V, words);
        ctx->p2p_step++;
        Span s(ctx, K_COUNT, 1);
        count_publish_kernel<<<V, 256, 0, ctx->stream>>>(ptr<uint64_t>(ctx->d_rows[cur]), words, V, ptr<uint32_t>(ctx->d_counts), ptr<uint32_t>(ctx->d_view_ids), pa,
                                                         ctx->p2p_step, ptr<uint32_t>(ctx->d_done));
        ctx->cur_published = true;
    } else {
        Span s(ctx, K_COUNT, 1);
        popcount_rows_kernel<<<V, 256, 0, ctx->stream>>>(ptr<uint64_t>(ctx->d_rows[cur]), words, ptr<uint32_t>(ctx->d_counts));
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_rows_ready[cur], ctx->stream));
    ctx->cur = cur;
    ctx->last_cast_axis = ctx->variant == PRV_VARIANT_AXIS;
    ctx->last_cast_views = V;
    ctx->last_mode = mode;
    ctx->have_pixels = pixels;
    ctx->cast_done = true;
    ctx->greedy_done = false;
    ctx->gathered = false;
    ctx->g_rows = ptr<uint64_t>(ctx->d_rows[cur]);
    ctx->g_nrows = V;
    return PRV_OK;
}

int greedy_impl(prv_ctx* ctx, uint32_t first_view, uint32_t max_iter) {
    if (!ctx->cast_done) return fail(ctx, PRV_ERR_INVALID, "prv_greedy: no coverage bitsets resident (call prv_cast_* first)");
    const std::vector<uint32_t>& h_map = ctx->gathered ? ctx->h_row_of_id_all : ctx->h_row_of_id;
    if (first_view >= h_map.size()) return fail(ctx, PRV_ERR_INVALID, "prv_greedy: first_view %u out of range", first_view);
    const uint32_t words = ctx->map.words64;
    int rc;
    if ((rc = ensure(ctx, ctx->d_best, 8 * ((size_t)max_iter + 2)))) return rc;
    if ((rc = ensure(ctx, ctx->d_cov[0], 8 * (size_t)words))) return rc;
    if ((rc = ensure(ctx, ctx->d_cov[1], 8 * (size_t)words))) return rc;
    cudaStream_t qs = ctx->score_stream;
    if (!ctx->gathered) CU(cudaStreamWaitEvent(qs, ctx->ev_rows_ready[ctx->cur], 0));  // (a gathered table was ordered by the all-gather)
    CU(cudaMemsetAsync(ctx->d_best.p, 0, 8 * ((size_t)max_iter + 2), qs));
    const uint32_t* ids = ctx->gathered ? ctx->g_ids : ptr<uint32_t>(ctx->d_view_ids);
    const uint32_t* row_of_id = ctx->gathered ? ptr<uint32_t>(ctx->d_row_of_id_all) : ptr<uint32_t>(ctx->d_row_of_id);
    // row index of first_view in the active table
    const uint32_t first_row = h_map[first_view];
    if (first_row == kNone) return fail(ctx, PRV_ERR_INVALID, "prv_greedy: first_view %u is not a resident view id", first_view);
    const size_t cov_bytes = (size_t)words * 8;
    // 1st choice: one thread-block cluster holding the whole table, sliced by columns, in its CTAs' shared memory
    bool launched = false;
    if (!ctx->greedy_no_cluster) {
        const uint32_t T = (uint32_t)kGreedyClusterThreads;
        for (uint32_t C = (uint32_t)kGreedyClusterMax; C >= 1 && !launched; C /= 2) {
            const uint32_t half = words / 2;
            const uint32_t slice = (half + C - 1) / C;
            const uint32_t vper = (ctx->g_nrows + C - 1) / C;
            const uint32_t vp32 = (ctx->g_nrows + 31u) & ~31u;
            const uint32_t parts = vp32 >= T ? 1u : std::max(1u, std::min(slice, T / vp32));
            const size_t smem = 16 * ((size_t)slice * ctx->g_nrows + 2 * (size_t)slice) + 4 * ((size_t)C * parts * vper + (size_t)parts * ctx->g_nrows + slice);
            if (vper > T || smem > (size_t)220 * 1024) continue;
            if (cudaFuncSetAttribute(greedy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();  // this much shared memory is not available: the table does not fit one cluster
                break;
            }
            if (C > 8 && cudaFuncSetAttribute(greedy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C);
            cfg.blockDim = dim3(T);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = qs;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, greedy_cluster_kernel, &cfg) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                continue;
            }
            Span s(ctx, K_GREEDY, 1, qs);
            const cudaError_t e = cudaLaunchKernelEx(&cfg, greedy_cluster_kernel, (const uint64_t*)ctx->g_rows, words, ctx->g_nrows, ids, first_row, first_view,
                                                     max_iter, slice, vper, parts, ptr<unsigned long long>(ctx->d_best), ptr<uint64_t>(ctx->d_cov[0]));
            if (e == cudaSuccess) {
                launched = true;
                ctx->greedy_persistent = true;
                if (getenv("PRV_VERBOSE")) fprintf(stderr, "[prv] greedy: cluster of %u CTAs, slice %u x 16 B, %u views/owner, %u parts, %zu B smem\n", C, slice, vper, parts, smem);
            } else {
                cudaGetLastError();
                // only "this cluster shape cannot be launched here" leads to the next shape / the grid-barrier kernel; anything
                // else is a real failure and is reported, not masked by a slower path
                if (e != cudaErrorInvalidConfiguration && e != cudaErrorLaunchOutOfResources && e != cudaErrorInvalidValue && e != cudaErrorNotSupported &&
                    e != cudaErrorInvalidClusterSize)
                    return fail(ctx, PRV_ERR_CUDA, "prv_greedy: cluster kernel launch failed: %s", cudaGetErrorString(e));
            }
        }
    }
    ctx->greedy_path = launched ? 0 : (cov_bytes <= 160 * 1024 ? 1 : 2);
    if (!launched) {  // (the cluster kernel writes every word of the mask itself)
        CU(cudaMemsetAsync(ctx->d_cov[0].p, 0, 8 * (size_t)words, qs));
        CU(cudaMemsetAsync(ctx->d_cov[1].p, 0, 8 * (size_t)words, qs));
    }
    if (launched) {
        // done
    } else if (cov_bytes <= 160 * 1024) {
        // one persistent cooperative kernel
        // grid: 2 (else 1) blocks per SM with the block's rows resident in shared memory next to the mask when that fits
        // (<= 100 KB per block at 2 blocks/SM, <= 200 KB at 1), else rows stay in global memory
        uint32_t grid = 1;
        size_t smem_bytes = cov_bytes;
        int rows_in_smem = 0;
        for (int bps = 2; bps >= 1 && !rows_in_smem; bps--) {
            const uint32_t g = std::max(1u, std::min<uint32_t>(ctx->g_nrows, (uint32_t)(ctx->sm_count * bps)));
            const uint32_t rpb = (ctx->g_nrows + g - 1) / g;
            const size_t need = cov_bytes * (1 + (size_t)rpb);
            if (need <= (size_t)(200 * 1024) / bps) {
                grid = g;
                smem_bytes = need;
                rows_in_smem = 1;
            }
        }
        if (!rows_in_smem) grid = std::max(1u, std::min<uint32_t>(ctx->g_nrows, (uint32_t)(ctx->sm_count * ctx->greedy_blocks_per_sm)));
        CU(cudaFuncSetAttribute(greedy_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, greedy_persistent_kernel, 256, smem_bytes));
        if (occ < 1 || (uint64_t)occ * ctx->sm_count < grid) {
            if (rows_in_smem || occ < 1) return fail(ctx, PRV_ERR_CUDA, "prv_greedy: persistent kernel does not fit (smem %zu, occupancy %d)", smem_bytes, occ);
            grid = (uint32_t)(occ * ctx->sm_count);
        }
        const uint64_t* rows_p = ctx->g_rows;
        uint32_t words_a = words, nrows_a = ctx->g_nrows, first_row_a = first_row, first_id_a = first_view, max_iter_a = max_iter;
        unsigned long long* best_p = ptr<unsigned long long>(ctx->d_best);
        uint64_t* cov_p = ptr<uint64_t>(ctx->d_cov[0]);
        unsigned int* arrive_p = nullptr;  // (unused: the grid barrier is cooperative_groups::grid_group::sync)
        void* args[] = {(void*)&rows_p, (void*)&words_a, (void*)&nrows_a, (void*)&ids, (void*)&row_of_id, (void*)&first_row_a, (void*)&first_id_a,
                        (void*)&max_iter_a, (void*)&best_p, (void*)&cov_p, (void*)&arrive_p, (void*)&rows_in_smem};
        Span s(ctx, K_GREEDY, 1, qs);
        CU(cudaLaunchCooperativeKernel((const void*)greedy_persistent_kernel, dim3(grid), dim3(256), args, smem_bytes, qs));
        ctx->greedy_persistent = true;
    } else {
        Span s(ctx, K_GREEDY, max_iter + 2, qs);
        greedy_init_kernel<<<1, 256, 0, qs>>>(ctx->g_rows, words, first_row, first_view, ptr<unsigned long long>(ctx->d_best));
        for (uint32_t k = 1; k <= max_iter + 1; k++) {
            const int cover_only = k == max_iter + 1;
            greedy_iter_kernel<<<cover_only ? 1 : ctx->g_nrows, 256, 0, qs>>>(
                ctx->g_rows, words, ids, row_of_id, k, ptr<unsigned long long>(ctx->d_best), ptr<uint64_t>(ctx->d_cov[(k - 1) & 1]),
                ptr<uint64_t>(ctx->d_cov[k & 1]), cover_only);
        }
        ctx->greedy_persistent = false;
    }
    CU(cudaGetLastError());
    if (!ctx->gathered || ctx->p2p_ready) {
        // the next cast into this row buffer waits for the selection that reads it -- also with the peer-memory exchange,
        // where this is what bounds how far a rank can run ahead of the readers of its stores (four table buffers)
        CU(cudaEventRecord(ctx->ev_rows_free[ctx->cur], qs));
        ctx->rows_free_pending[ctx->cur] = true;
    }
    ctx->greedy_max_iter = max_iter;
    ctx->greedy_done = true;
    return PRV_OK;
}

int render_impl(prv_ctx* ctx, uint32_t V, int point_size, bool want_depth) {
    if (!ctx->have_cam || ctx->P == 0 || ctx->V < V || V == 0) return fail(ctx, PRV_ERR_INVALID, "prv_render: camera, cloud and views must be set");
    if (point_size < 1 || point_size > 32) return fail(ctx, PRV_ERR_INVALID, "prv_render: point_size %d out of [1,32]", point_size);
    const int W = ctx->cam.W, H = ctx->cam.H;
    const int Wc = W + point_size - 1, Hc = H + point_size - 1;
    const size_t corner_bytes = (size_t)Wc * Hc * 8;
    // keep the corner buffers of one chunk around L2 size
    uint32_t chunk = (uint32_t)std::max<size_t>(1, (size_t)(96u << 20) / corner_bytes);
    chunk = std::min<uint32_t>(std::min<uint32_t>(chunk, V), 65535u);
    int rc;
    if ((rc = ensure(ctx, ctx->d_corner, corner_bytes * chunk))) return rc;
    if ((rc = ensure(ctx, ctx->d_rgba, (size_t)V * W * H * 4))) return rc;
    if (want_depth && (rc = ensure(ctx, ctx->d_depth_img, (size_t)V * W * H * 4))) return rc;
    const float focal = prv_splat_focal(&ctx->intr);
    const size_t smem = ((size_t)(32 + point_size - 1) + 32) * (8 + point_size - 1) * 8;  // tile + row minima
    for (uint32_t vb = 0; vb < V; vb += chunk) {
        const uint32_t vn = std::min(chunk, V - vb);
        CU(cudaMemsetAsync(ctx->d_corner.p, 0xFF, corner_bytes * vn, ctx->stream));
        {
            Span s(ctx, K_SPLAT, 1);
            splat_points_kernel<<<dim3((unsigned)((ctx->P + 255) / 256), vn), 256, 0, ctx->stream>>>(
                ptr<float>(ctx->d_cloud_xyz), ctx->P, ctx->cam, ptr<ViewConst>(ctx->d_views), vb, focal, point_size,
                ptr<unsigned long long>(ctx->d_corner), Wc, Hc);
        }
        {
            Span s(ctx, K_RESOLVE, 1);
            splat_resolve_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, vn), 256, smem, ctx->stream>>>(
                ptr<unsigned long long>(ctx->d_corner), Wc, Hc, W, H, point_size, ptr<uint8_t>(ctx->d_cloud_rgb), ptr<uint8_t>(ctx->d_rgba),
                want_depth ? ptr<float>(ctx->d_depth_img) : nullptr, vb);
        }
    }
    CU(cudaGetLastError());
    ctx->rendered_views = V;
    return PRV_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* prv_last_error(const prv_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int prv_create(prv_ctx** out, int device) {
    if (!out) return fail(nullptr, PRV_ERR_INVALID, "prv_create: null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, PRV_ERR_NO_DEVICE, "prv_create: no CUDA device (%s); libprv_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, PRV_ERR_INVALID, "prv_create: device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, PRV_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, PRV_ERR_NO_DEVICE, "prv_create: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, PRV_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    prv_ctx* ctx = new prv_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->mem_bytes = prop.totalGlobalMem;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, PRV_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->score_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, PRV_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    for (auto& s : ctx->slots) cudaEventCreate(&s);
    cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_flush, cudaEventDisableTiming);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&ctx->ev_rows_ready[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_rows_free[i], cudaEventDisableTiming);
    }
    if (const char* e = getenv("PRV_GREEDY_CLUSTER")) ctx->greedy_no_cluster = atoi(e) == 0;
    *out = ctx;
    return PRV_OK;
}

void prv_destroy(prv_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->score_stream);
    prv_comm_destroy(ctx);
    DevBuf* bufs[] = {&ctx->d_tilesum, &ctx->d_check, &ctx->d_deproj, &ctx->d_arena, &ctx->d_done, &ctx->d_counts_scratch, &ctx->d_coarse, &ctx->d_bitmap, &ctx->d_bitmap_pad, &ctx->d_prefix, &ctx->d_leaf_of_raster, &ctx->d_keys, &ctx->d_rgb, &ctx->d_views, &ctx->d_view_ids,
                      &ctx->d_row_of_id, &ctx->d_rows[0], &ctx->d_rows[1], &ctx->d_counts, &ctx->d_zero, &ctx->d_queue, &ctx->d_queue2, &ctx->d_queue2b, &ctx->d_arrive, &ctx->d_ens_images, &ctx->d_ens_terms, &ctx->d_ens_scores, &ctx->d_row_of_id_all, &ctx->d_ing_xyz, &ctx->d_ing_rgb, &ctx->d_ing_k0,
                      &ctx->d_ing_k1, &ctx->d_ing_v0, &ctx->d_ing_v1, &ctx->d_ing_pos, &ctx->d_ing_tmp, &ctx->d_pix_hit, &ctx->d_pix_depth, &ctx->d_mask,
                      &ctx->d_voxel_pix, &ctx->d_voxel_hit, &ctx->d_points, &ctx->d_best, &ctx->d_cov[0], &ctx->d_cov[1], &ctx->d_all_rows,
                      &ctx->d_all_ids, &ctx->d_cloud_xyz, &ctx->d_cloud_rgb, &ctx->d_corner, &ctx->d_rgba, &ctx->d_depth_img, &ctx->d_flush};
    for (DevBuf* b : bufs) release(*b);
    for (auto& s : ctx->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    for (auto& e : ctx->free_events) cudaEventDestroy(e);
    for (auto& s : ctx->slots) cudaEventDestroy(s);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_flush) cudaEventDestroy(ctx->ev_flush);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_rows_ready[i]) cudaEventDestroy(ctx->ev_rows_ready[i]);
        if (ctx->ev_rows_free[i]) cudaEventDestroy(ctx->ev_rows_free[i]);
    }
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaStreamDestroy(ctx->score_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int prv_device_info(prv_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, uint64_t* mem_bytes) {
    if (!ctx) return PRV_ERR_INVALID;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    return PRV_OK;
}

int prv_sync(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(sync_all(ctx));
    return PRV_OK;
}

int prv_set_variant(prv_ctx* ctx, int variant) {
    if (!ctx) return PRV_ERR_INVALID;
    if (variant < PRV_VARIANT_PLAIN || variant > PRV_VARIANT_AXIS) return fail(ctx, PRV_ERR_INVALID, "prv_set_variant: unknown variant %d", variant);
    ctx->variant = variant;
    return PRV_OK;
}

int prv_set_staging(prv_ctx* ctx, int bitmap_in_shared_memory, int l2_persisting_window) {
    if (!ctx) return PRV_ERR_INVALID;
    ctx->stage_smem = bitmap_in_shared_memory != 0;
    ctx->stage_l2 = l2_persisting_window != 0;
    if (!ctx->stage_l2) {
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.num_bytes = 0;
        if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
    return PRV_OK;
}

int prv_set_brick_cull(prv_ctx* ctx, int cell, int enter_at_brick) {
    if (!ctx) return PRV_ERR_INVALID;
    if (cell != 0 && cell != 4 && cell != 8 && cell != 16) return fail(ctx, PRV_ERR_INVALID, "prv_set_brick_cull: brick edge must be 0 (automatic), 4, 8 or 16 voxels, got %d", cell);
    ctx->brick_cs = cell;
    ctx->brick_entry = enter_at_brick != 0;
    return PRV_OK;
}

// keys/rgb: host tables in leaf order, or (device_ready) d_keys / d_rgb already hold them (GPU ingest path; keys == nullptr).
static int set_map_impl(prv_ctx* ctx, const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution, bool device_ready) {
    if (!ctx) return PRV_ERR_INVALID;
    if ((!keys && !device_ready) || N == 0 || !(resolution > 0)) return fail(ctx, PRV_ERR_INVALID, "prv_set_map: null/empty map or bad resolution");
    CU(cudaSetDevice(ctx->device));
    CU(join_score(ctx));  // a selection still running on the scoring stream reads tables this call replaces
    int rc;
    if (!device_ready) {
        if ((rc = ensure(ctx, ctx->d_keys, (size_t)N * 6))) return rc;
        if ((rc = ensure(ctx, ctx->d_rgb, (size_t)N * 3))) return rc;
        // the only host->device traffic: leaf keys and colours; every table is built by kernels
        CU(h2d(ctx, ctx->d_keys.p, keys, (size_t)N * 6));
        if (rgb)
            CU(h2d(ctx, ctx->d_rgb.p, rgb, (size_t)N * 3));
        else
            CU(cudaMemsetAsync(ctx->d_rgb.p, 0, (size_t)N * 3, ctx->stream));
    }
    // leaf order (strictly ascending Morton code, no duplicates) and AABB of the keys: one kernel over the resident table
    if ((rc = ensure(ctx, ctx->d_check, 8 * 4))) return rc;
    uint32_t chk[7] = {N, 65535u, 65535u, 65535u, 0u, 0u, 0u};
    CU(cudaMemcpyAsync(ctx->d_check.p, chk, sizeof(chk), cudaMemcpyHostToDevice, ctx->stream));
    {
        Span sp(ctx, K_OTHER, 1);
        map_check_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ptr<uint16_t>(ctx->d_keys), N, ptr<uint32_t>(ctx->d_check));
    }
    CU(cudaMemcpyAsync(chk, ctx->d_check.p, sizeof(chk), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // (the caller's key / colour buffers have been consumed by now as well)
    if (chk[0] < N) return fail(ctx, PRV_ERR_INVALID, "prv_set_map: keys are not in strict leaf (Morton) order at index %u", chk[0]);
    const int lo[3] = {(int)chk[1], (int)chk[2], (int)chk[3]}, hi[3] = {(int)chk[4], (int)chk[5], (int)chk[6]};
    int n[3];
    for (int a = 0; a < 3; a++) n[a] = hi[a] - lo[a] + 1;
    const int wx = (n[0] + 31) / 32;
    const size_t nwords = (size_t)wx * n[1] * n[2];
    if (nwords > ((size_t)1 << 31)) return fail(ctx, PRV_ERR_UNSUPPORTED, "prv_set_map: occupancy AABB %dx%dx%d too large for a dense bitmap", n[0], n[1], n[2]);
    // shell-padded bitmap geometry (branch-free in-AABB march): rows padded to a power of two, 3 planes (+1 row) of
    // slack on both sides because the 4-deep speculative march may step up to 3 cells past the shell
    int row_log2 = 5;
    while ((1 << row_log2) < n[0] + 2) row_log2++;
    const size_t pad_rows = (size_t)(n[1] + 2) * (n[2] + 2);
    const size_t slack_bits = ((size_t)3 * (n[1] + 2) + 2) << row_log2;
    const size_t pad_words = ((pad_rows << row_log2) + 2 * slack_bits) / 32;
    if ((pad_rows << row_log2) + 2 * slack_bits >= ((size_t)1 << 31)) return fail(ctx, PRV_ERR_UNSUPPORTED, "prv_set_map: occupancy AABB too large for the padded bitmap");
    int nc[3];
    // automatic: 4-voxel bricks while the grid stays small (AABB up to 64 voxels: the 0.002 m maps; the walk is short and the
    // finer bricks save the march more than they cost), else 8 (measured: C1 -2.8 %, C3 -0.8 %, C2 +3.4 % with 4 instead of 8)
    const int cs = ctx->brick_cs ? ctx->brick_cs : (std::max(n[0], std::max(n[1], n[2])) <= 64 ? 4 : kCoarseDefault);
    for (int a = 0; a < 3; a++) nc[a] = (n[a] + cs - 1) / cs;
    const size_t coarse_words = ((size_t)nc[0] * nc[1] * nc[2] + 31) / 32 + 1;
    if ((rc = ensure(ctx, ctx->d_coarse, coarse_words * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_bitmap_pad, pad_words * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_bitmap, nwords * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_prefix, nwords * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_leaf_of_raster, (size_t)N * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_tilesum, ((nwords + kPrefixTile - 1) / kPrefixTile + 1) * 4))) return rc;
    CU(cudaMemsetAsync(ctx->d_coarse.p, 0, coarse_words * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_bitmap_pad.p, 0, pad_words * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_bitmap.p, 0, nwords * 4, ctx->stream));
    {
        MapBuild mb{};
        for (int a = 0; a < 3; a++) {
            mb.lo[a] = lo[a];
            mb.n[a] = n[a];
            mb.nc[a] = nc[a];
        }
        mb.wx = wx;
        mb.row_log2 = row_log2;
        mb.n_occ = N;
        mb.slack_bits = slack_bits;
        mb.nwords = nwords;
        mb.bitmap = ptr<uint32_t>(ctx->d_bitmap);
        mb.pad = ptr<uint32_t>(ctx->d_bitmap_pad);
        mb.coarse = ptr<uint32_t>(ctx->d_coarse);
        mb.prefix = ptr<uint32_t>(ctx->d_prefix);
        mb.leaf_of_raster = ptr<uint32_t>(ctx->d_leaf_of_raster);
        mb.keys = ptr<uint16_t>(ctx->d_keys);
        mb.cs = cs;
        Span sp(ctx, K_OTHER, 5);
        map_scatter_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(mb);
        map_shell_kernel<<<(uint32_t)((pad_rows + 255) / 256), 256, 0, ctx->stream>>>(mb);
        const uint32_t ntiles = (uint32_t)((nwords + kPrefixTile - 1) / kPrefixTile);
        map_tilesum_kernel<<<ntiles, kPrefixThreads, 0, ctx->stream>>>(mb, ptr<uint32_t>(ctx->d_tilesum));
        map_prefix_kernel<<<ntiles, kPrefixThreads, 0, ctx->stream>>>(mb, ptr<uint32_t>(ctx->d_tilesum));
        map_rank_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(mb);
    }
    CU(cudaGetLastError());
    // host copy of the keys (per-view origin-in-object test of prv_set_views, prv_get_map).  The table kernels keep running
    // on the ctx stream, which orders them before everything that uses the map.
    PRV_GUARD_BEGIN
    if (device_ready) {
        ctx->h_keys.resize((size_t)N * 3);
        CU(d2h(ctx, ctx->h_keys.data(), ctx->d_keys.p, (size_t)N * 6));
        CU(cudaStreamSynchronize(ctx->stream));
    } else {
        ctx->h_keys.assign(keys, keys + (size_t)N * 3);
    }
    PRV_GUARD_END(ctx, "prv_set_map")
    ctx->resolution = resolution;
    for (int a = 0; a < 3; a++) {
        ctx->lo[a] = lo[a];
        ctx->n[a] = n[a];
        ctx->map.lo[a] = lo[a];
        ctx->map.n[a] = n[a];
    }
    ctx->map.resolution = resolution;
    ctx->map.wx = wx;
    ctx->map.n_occ = N;
    ctx->map.words64 = words_for(N);
    ctx->map.bitmap = ptr<uint32_t>(ctx->d_bitmap);
    ctx->map.bitmap_pad = ptr<uint32_t>(ctx->d_bitmap_pad);
    {
        double rad2 = 0.0;
        for (int a = 0; a < 3; a++) {
            const double lo_m = (double)(lo[a] - 2 - prv::kTreeMaxVal) * resolution, hi_m = (double)(lo[a] + n[a] + 2 - prv::kTreeMaxVal) * resolution;
            ctx->map.bcen[a] = (float)(0.5 * (lo_m + hi_m));
            rad2 += 0.25 * (hi_m - lo_m) * (hi_m - lo_m);
        }
        ctx->map.brad = (float)(std::sqrt(rad2) * 1.001);
    }
    ctx->map.coarse = ptr<uint32_t>(ctx->d_coarse);
    for (int a = 0; a < 3; a++) ctx->map.nc[a] = nc[a];
    for (int a = 0; a < 3; a++) ctx->map.nhi[a] = (float)n[a] + 1.0f;
    ctx->map.cs = cs;
    ctx->map.inv_cs = 1.0f / (float)cs;
    {   // row / (n1 + 2) of the march epilogue as a multiply-high: exact while row * (n1 + 2) < 2^32
        const uint64_t n1p = (uint64_t)n[1] + 2;
        const uint64_t magic = (((uint64_t)1 << 32) + n1p - 1) / n1p;
        ctx->map.n1p_magic = (pad_rows + 8) * n1p < ((uint64_t)1 << 32) && magic < ((uint64_t)1 << 32) ? (uint32_t)magic : 0u;
    }
    ctx->map.pad_row_log2 = row_log2;
    ctx->map.pad_words = (uint32_t)pad_words;
    if (ctx->stage_l2) {  // A/B: pin the padded bitmap in L2 for the cast stream
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.base_ptr = ctx->d_bitmap_pad.p;
        av.accessPolicyWindow.num_bytes = pad_words * 4;
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)16 << 20, pad_words * 4 + ((size_t)1 << 20)));
        if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
    ctx->map.pad_bit_offset = (uint32_t)slack_bits;
    for (int a = 0; a < 3; a++) {
        ctx->map.bmin[a] = (float)((double)(lo[a] - 2 - prv::kTreeMaxVal) * resolution);
        ctx->map.bmax[a] = (float)((double)(lo[a] + n[a] + 2 - prv::kTreeMaxVal) * resolution);
    }
    ctx->map.prefix = ptr<uint32_t>(ctx->d_prefix);
    ctx->map.leaf_of_raster = ptr<uint32_t>(ctx->d_leaf_of_raster);
    ctx->map.keys = ptr<uint16_t>(ctx->d_keys);
    ctx->map.rgb = ptr<uint8_t>(ctx->d_rgb);
    ctx->have_map = true;
    ctx->V = 0;
    ctx->cast_done = false;
    ctx->greedy_done = false;
    return PRV_OK;
}

int prv_set_map(prv_ctx* ctx, const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution) {
    return set_map_impl(ctx, keys, rgb, N, resolution, false);
}

int prv_set_map_from_cloud(prv_ctx* ctx, const float* xyz, const uint8_t* rgb, uint64_t P, double resolution) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!xyz || P == 0 || P > 0x7FFFFFFFull || !(resolution > 0)) return fail(ctx, PRV_ERR_INVALID, "prv_set_map_from_cloud: null/empty cloud or bad resolution");
    CU(cudaSetDevice(ctx->device));
    CU(join_score(ctx));
    int rc;
    const uint32_t n = (uint32_t)P;
    if ((rc = ensure(ctx, ctx->d_ing_xyz, P * 12))) return rc;
    if ((rc = ensure(ctx, ctx->d_ing_rgb, P * 3))) return rc;
    if ((rc = ensure(ctx, ctx->d_ing_k0, P * 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_ing_k1, P * 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_ing_v0, (P + 1) * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_ing_v1, P * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_ing_pos, (P + 1) * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_keys, P * 6))) return rc;
    if ((rc = ensure(ctx, ctx->d_rgb, P * 3))) return rc;
    CU(h2d(ctx, ctx->d_ing_xyz.p, xyz, P * 12));
    if (rgb)
        CU(h2d(ctx, ctx->d_ing_rgb.p, rgb, P * 3));
    else
        CU(cudaMemsetAsync(ctx->d_ing_rgb.p, 0, P * 3, ctx->stream));
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, ptr<unsigned long long>(ctx->d_ing_k0), ptr<unsigned long long>(ctx->d_ing_k1),
                                    ptr<uint32_t>(ctx->d_ing_v0), ptr<uint32_t>(ctx->d_ing_v1), (int)n, 0, 49, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, ptr<uint32_t>(ctx->d_ing_v0), ptr<uint32_t>(ctx->d_ing_pos), (int)n + 1, ctx->stream);
    if ((rc = ensure(ctx, ctx->d_ing_tmp, std::max(tmp_sort, tmp_scan)))) return rc;
    size_t tmp_bytes = ctx->d_ing_tmp.cap;
    {
        Span sp(ctx, K_OTHER, 5);
        ingest_keys_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ptr<float>(ctx->d_ing_xyz), n, 1.0 / resolution, ptr<unsigned long long>(ctx->d_ing_k0),
                                                                   ptr<uint32_t>(ctx->d_ing_v0));
        // stable LSD radix sort on the 48-bit Morton code (+1 bit: invalid points sort last): points of one voxel stay in
        // cloud order, so the first one is the point whose colour the reference keeps (main.cpp:1015-1021)
        CU(cub::DeviceRadixSort::SortPairs(ctx->d_ing_tmp.p, tmp_bytes, ptr<unsigned long long>(ctx->d_ing_k0), ptr<unsigned long long>(ctx->d_ing_k1),
                                           ptr<uint32_t>(ctx->d_ing_v0), ptr<uint32_t>(ctx->d_ing_v1), (int)n, 0, 49, ctx->stream));
        ingest_heads_kernel<<<(n + 1 + 255) / 256, 256, 0, ctx->stream>>>(ptr<unsigned long long>(ctx->d_ing_k1), n, ptr<uint32_t>(ctx->d_ing_v0));
        tmp_bytes = ctx->d_ing_tmp.cap;
        CU(cub::DeviceScan::ExclusiveSum(ctx->d_ing_tmp.p, tmp_bytes, ptr<uint32_t>(ctx->d_ing_v0), ptr<uint32_t>(ctx->d_ing_pos), (int)n + 1, ctx->stream));
        ingest_compact_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ptr<unsigned long long>(ctx->d_ing_k1), ptr<uint32_t>(ctx->d_ing_v1), ptr<uint32_t>(ctx->d_ing_v0),
                                                                      ptr<uint32_t>(ctx->d_ing_pos), n, ptr<uint8_t>(ctx->d_ing_rgb), ptr<uint16_t>(ctx->d_keys),
                                                                      ptr<uint8_t>(ctx->d_rgb));
    }
    CU(cudaGetLastError());
    uint32_t N = 0;
    CU(d2h(ctx, &N, ptr<uint32_t>(ctx->d_ing_pos) + n, 4));
    CU(cudaStreamSynchronize(ctx->stream));
    if (N == 0) return fail(ctx, PRV_ERR_INVALID, "prv_set_map_from_cloud: no point has a valid key");
    return set_map_impl(ctx, nullptr, nullptr, N, resolution, true);  // keys and colours stay on the device
}

int prv_get_map(prv_ctx* ctx, uint16_t* keys_out, uint8_t* rgb_out) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!ctx->have_map) return fail(ctx, PRV_ERR_INVALID, "prv_get_map: no map set");
    CU(cudaSetDevice(ctx->device));
    const size_t N = ctx->map.n_occ;
    if (keys_out) std::memcpy(keys_out, ctx->h_keys.data(), N * 6);
    if (rgb_out) {
        CU(d2h(ctx, rgb_out, ctx->d_rgb.p, N * 3));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return PRV_OK;
}

int prv_set_camera(prv_ctx* ctx, const prv_intrinsics* intr, double max_range) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!intr || intr->width <= 0 || intr->height <= 0) return fail(ctx, PRV_ERR_INVALID, "prv_set_camera: bad intrinsics");
    if (intr->model < 0 || intr->model > 5) return fail(ctx, PRV_ERR_INVALID, "prv_set_camera: unknown distortion model %d", intr->model);
    if (intr->model == 1)
        return fail(ctx, PRV_ERR_UNSUPPORTED, "prv_set_camera: model 1 (MODIFIED_BROWN_CONRADY) cannot be deprojected (reference asserts, Share_Data.hpp:142)");
    if ((long long)(intr->width + 1) * (intr->height + 1) > (1ll << 31) || intr->width >= 65535 || intr->height >= 65535)
        return fail(ctx, PRV_ERR_UNSUPPORTED, "prv_set_camera: image too large (queues pack pixel coordinates in 16+16 bits)");
    ctx->intr = *intr;
    ctx->cam.W = intr->width;
    ctx->cam.H = intr->height;
    ctx->cam.ppx = intr->ppx;
    ctx->cam.ppy = intr->ppy;
    ctx->cam.fx = intr->fx;
    ctx->cam.fy = intr->fy;
    ctx->cam.model = intr->model;
    for (int i = 0; i < 5; i++) ctx->cam.c[i] = intr->coeffs[i];
    ctx->cam.max_range = max_range;
    ctx->cam.max_range_sq = max_range * max_range;
    ctx->cam.inv_fx = 1.0f / intr->fx;
    ctx->cam.inv_fy = 1.0f / intr->fy;
    const bool same_intr = ctx->have_cam && std::memcmp(&ctx->intr_checked, intr, sizeof(prv_intrinsics)) == 0;
    if (!same_intr) {
        const bool transcendental = intr->model == 3 || intr->model == 5;
        const bool ok = !transcendental && region_cull_valid(*intr);
        ctx->cam.region_cull_ok = ok ? 1 : 0;
        ctx->cam.deproj_table = nullptr;
        ctx->cam.deproj_exact = nullptr;
        if (transcendental) {
            // F-Theta / Kannala-Brandt: tan / atan are the host libm's (as in the reference built on this host), so the
            // deprojection of every integer pixel is tabulated here and the kernels only read it
            CU(cudaSetDevice(ctx->device));
            CU(join_score(ctx));
            PRV_GUARD_BEGIN
            const size_t GW = (size_t)intr->width + 1, GH = (size_t)intr->height + 1;
            std::vector<float> tab(GW * GH * 2);
            for (size_t y = 0; y < GH; y++)
                for (size_t x = 0; x < GW; x++) {
                    const float pix[2] = {(float)x, (float)y};
                    float pt[3];
                    prv_host_deproject_pixel_to_point(intr, pix, 1.0f, pt);
                    tab[2 * (y * GW + x)] = pt[0];
                    tab[2 * (y * GW + x) + 1] = pt[1];
                }
            int rc;
            if ((rc = ensure(ctx, ctx->d_deproj, tab.size() * 4))) return rc;
            CU(h2d(ctx, ctx->d_deproj.p, tab.data(), tab.size() * 4));
            CU(cudaStreamSynchronize(ctx->stream));
            ctx->cam.deproj_table = ptr<float2>(ctx->d_deproj);
            ctx->cam.deproj_exact = ctx->cam.deproj_table;
            PRV_GUARD_END(ctx, "prv_set_camera")
        } else {
            const char* e = getenv("PRV_DEPROJ_TABLE");  // A/B switch: 0 = the march evaluates deproject_pixel per ray
            if (!(e && e[0] == '0') && ((size_t)intr->width + 1) * ((size_t)intr->height + 1) * 8 <= ((size_t)1 << 30)) {  // (an image too large for a 1 GiB table: per ray)
                CU(cudaSetDevice(ctx->device));
                CU(join_score(ctx));
                const size_t GW = (size_t)intr->width + 1, GH = (size_t)intr->height + 1;
                int rc;
                if ((rc = ensure(ctx, ctx->d_deproj, GW * GH * 8))) return rc;
                DevCam cam = ctx->cam;
                cam.deproj_exact = nullptr;
                deproj_table_kernel<<<dim3((unsigned)((GW + 31) / 32), (unsigned)((GH + 7) / 8)), 256, 0, ctx->stream>>>(cam, ptr<float2>(ctx->d_deproj));
                CU(cudaGetLastError());
                ctx->cam.deproj_exact = ptr<float2>(ctx->d_deproj);
            }
        }
        ctx->intr_checked = *intr;
    }
    ctx->have_cam = true;
    ctx->V = 0;  // per-view fast-path proofs depend on max_range
    ctx->cast_done = false;
    return PRV_OK;
}

int prv_set_view_ids(prv_ctx* ctx, const uint32_t* ids, uint32_t V) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!ids || V == 0) {
        ctx->h_view_ids.clear();
        ctx->ids_user = false;
        return PRV_OK;
    }
    PRV_GUARD_BEGIN
    // ids index host and device tables of max_id + 1 entries and ride in the low word of the packed arg-max (~id): they must
    // be distinct and small (kMaxViewId; the north-star workload has 1024 views)
    for (uint32_t v = 0; v < V; v++)
        if (ids[v] >= kMaxViewId) return fail(ctx, PRV_ERR_INVALID, "prv_set_view_ids: id %u at index %u is not below %u", ids[v], v, kMaxViewId);
    std::vector<uint32_t> sorted(ids, ids + V);
    std::sort(sorted.begin(), sorted.end());
    for (uint32_t v = 1; v < V; v++)
        if (sorted[v] == sorted[v - 1]) return fail(ctx, PRV_ERR_INVALID, "prv_set_view_ids: id %u appears more than once", sorted[v]);
    ctx->h_view_ids.assign(ids, ids + V);
    ctx->ids_user = true;
    PRV_GUARD_END(ctx, "prv_set_view_ids")
    return PRV_OK;
}

int prv_set_views(prv_ctx* ctx, const double* pose_world, const double* init_pos, uint32_t V) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    return upload_views(ctx, pose_world, init_pos, V);
}

uint32_t prv_full_voxels(const prv_ctx* ctx) { return ctx && ctx->have_map ? ctx->map.n_occ : 0; }
uint32_t prv_bitset_words(const prv_ctx* ctx) { return ctx && ctx->have_map ? ctx->map.words64 : 0; }
uint32_t prv_num_views(const prv_ctx* ctx) { return ctx ? ctx->V : 0; }

int prv_cast_async(prv_ctx* ctx, int mode, int flags) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    return cast_impl(ctx, mode, flags & PRV_CAST_PIXELS, (flags & PRV_CAST_PUBLISH) != 0);
}

int prv_greedy_async(prv_ctx* ctx, uint32_t first_view, uint32_t max_iter) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    return greedy_impl(ctx, first_view, max_iter);
}

int prv_get_bitsets(prv_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return PRV_ERR_INVALID;
    if (!ctx->cast_done) return fail(ctx, PRV_ERR_INVALID, "prv_get_bitsets: nothing cast yet");
    CU(cudaSetDevice(ctx->device));
    CU(d2h_result(ctx, out, ctx->d_rows[ctx->cur].p, (size_t)ctx->V * ctx->map.words64 * 8));
    return PRV_OK;
}

int prv_get_coverage_counts(prv_ctx* ctx, uint32_t* out) {
    if (!ctx || !out) return PRV_ERR_INVALID;
    if (!ctx->cast_done) return fail(ctx, PRV_ERR_INVALID, "prv_get_coverage_counts: nothing cast yet");
    CU(cudaSetDevice(ctx->device));
    CU(d2h_result(ctx, out, ctx->d_counts.p, (size_t)ctx->V * 4));
    return PRV_OK;
}

int prv_get_hit_rank(prv_ctx* ctx, uint32_t view_begin, uint32_t view_count, uint32_t* out) {
    if (!ctx || !out) return PRV_ERR_INVALID;
    if (!ctx->cast_done || !ctx->have_pixels) return fail(ctx, PRV_ERR_INVALID, "prv_get_hit_rank: last cast did not keep per-pixel results");
    if ((uint64_t)view_begin + view_count > ctx->V) return fail(ctx, PRV_ERR_INVALID, "prv_get_hit_rank: view range out of bounds");
    CU(cudaSetDevice(ctx->device));
    if (ctx->last_mode == PRV_MODE_DENSE) {
        CU(d2h(ctx, out, ptr<uint32_t>(ctx->d_pix_hit) + (size_t)view_begin * ctx->pix_stride, (size_t)view_count * ctx->pix_stride * 4));
    } else {
        const uint32_t N = ctx->map.n_occ;
        int rc;
        if ((rc = ensure(ctx, ctx->d_voxel_hit, (size_t)view_count * N * 4))) return rc;
        {
            Span s(ctx, K_OTHER, 1);
            gather_voxel_hits_kernel<<<dim3((N + 255) / 256, view_count), 256, 0, ctx->stream>>>(
                N, ptr<uint32_t>(ctx->d_voxel_pix) + (size_t)view_begin * N, ptr<uint32_t>(ctx->d_pix_hit) + (size_t)view_begin * ctx->pix_stride,
                ctx->pix_stride, ptr<uint32_t>(ctx->d_voxel_hit));
        }
        CU(cudaGetLastError());
        CU(d2h(ctx, out, ctx->d_voxel_hit.p, (size_t)view_count * N * 4));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return PRV_OK;
}

int prv_get_depth(prv_ctx* ctx, uint32_t view_begin, uint32_t view_count, float* out) {
    if (!ctx || !out) return PRV_ERR_INVALID;
    if (!ctx->cast_done || !ctx->have_pixels || ctx->last_mode != PRV_MODE_DENSE)
        return fail(ctx, PRV_ERR_INVALID, "prv_get_depth: last cast was not a dense cast with per-pixel results");
    if ((uint64_t)view_begin + view_count > ctx->V) return fail(ctx, PRV_ERR_INVALID, "prv_get_depth: view range out of bounds");
    CU(cudaSetDevice(ctx->device));
    CU(d2h(ctx, out, ptr<float>(ctx->d_pix_depth) + (size_t)view_begin * ctx->pix_stride, (size_t)view_count * ctx->pix_stride * 4));
    CU(cudaStreamSynchronize(ctx->stream));
    return PRV_OK;
}

int prv_get_greedy(prv_ctx* ctx, uint32_t* seq, uint32_t* gains, uint32_t capacity, uint32_t* n_out, uint64_t* covered_out) {
    if (!ctx || !seq || !gains || !n_out) return PRV_ERR_INVALID;
    if (!ctx->greedy_done) return fail(ctx, PRV_ERR_INVALID, "prv_get_greedy: prv_greedy_async has not run");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->score_stream));
    if (ctx->p2p_ready && ctx->gathered) {
        uint32_t err = 0;
        CU(cudaMemcpy(&err, arena_error(ctx->d_arena.p), 4, cudaMemcpyDeviceToHost));
        if (err) return fail(ctx, PRV_ERR_NCCL, "prv_get_greedy: rank %u never published its rows (peer exchange timed out); the selection ran on stale rows", err - 1);
    }
    PRV_GUARD_BEGIN
    std::vector<unsigned long long> best((size_t)ctx->greedy_max_iter + 2);
    CU(d2h_result(ctx, best.data(), ctx->d_best.p, 8 * best.size()));
    uint32_t n = 0;
    for (uint32_t k = 0; k <= ctx->greedy_max_iter; k++) {
        if (k > 0 && (uint32_t)(best[k] >> 32) == 0) break;
        n++;
    }
    *n_out = n;
    // the selection ran with the max_iter of the prv_greedy_async call, which need not be what the caller sized seq / gains for
    if (n > capacity) return fail(ctx, PRV_ERR_INVALID, "prv_get_greedy: %u entries selected but seq / gains hold %u (size them max_iter + 1)", n, capacity);
    for (uint32_t k = 0; k < n; k++) {
        seq[k] = 0xFFFFFFFFu - (uint32_t)(best[k] & 0xFFFFFFFFull);
        gains[k] = (uint32_t)(best[k] >> 32);
    }
    if (covered_out) {
        CU(d2h_result(ctx, covered_out, ctx->d_cov[ctx->greedy_persistent ? 0 : (n & 1)].p, 8 * (size_t)ctx->map.words64));
    }
    PRV_GUARD_END(ctx, "prv_get_greedy")
    return PRV_OK;
}

int prv_greedy_path(const prv_ctx* ctx) { return ctx ? ctx->greedy_path : -1; }

int prv_get_cast_stats(prv_ctx* ctx, prv_cast_stats* out) {
    if (!ctx || !out) return PRV_ERR_INVALID;
    if (!ctx->cast_done) return fail(ctx, PRV_ERR_INVALID, "prv_get_cast_stats: nothing cast yet");
    CU(cudaSetDevice(ctx->device));
    PRV_GUARD_BEGIN
    // counters and queue lengths of the LATEST cast (its variant and view count, not whatever is selected now)
    const uint32_t V = ctx->last_cast_views;
    std::vector<unsigned long long> s((size_t)V * 4);
    std::vector<uint32_t> q(ctx->last_cast_axis ? V : 0);
    CU(d2h(ctx, s.data(), ctx->p_stats, s.size() * 8));
    if (ctx->last_cast_axis) CU(d2h(ctx, q.data(), ctx->p_qcount + V, q.size() * 4));
    CU(cudaStreamSynchronize(ctx->stream));
    out->rays = out->probes_in = out->hits = out->steps = out->marched = 0;
    for (uint32_t c : q) out->marched += c;
    for (uint32_t v = 0; v < V; v++) {
        out->rays += s[4 * (size_t)v + 0];
        out->probes_in += s[4 * (size_t)v + 1];
        out->hits += s[4 * (size_t)v + 2];
        out->steps += s[4 * (size_t)v + 3];
    }
    if (!ctx->last_cast_axis) out->marched = out->rays;
    PRV_GUARD_END(ctx, "prv_get_cast_stats")
    return PRV_OK;
}

int prv_cast_views(prv_ctx* ctx, const double* pose_world, const double* init_pos, uint32_t V, int mode, uint64_t* bitsets_out,
                   uint32_t* coverage_count_out, uint32_t* hit_rank_out, float* depth_out) {
    if (!ctx) return PRV_ERR_INVALID;
    int rc;
    if ((rc = prv_set_views(ctx, pose_world, init_pos, V))) return rc;
    if ((rc = cast_impl(ctx, mode, hit_rank_out || depth_out))) return rc;
    if (bitsets_out && (rc = prv_get_bitsets(ctx, bitsets_out))) return rc;
    if (coverage_count_out && (rc = prv_get_coverage_counts(ctx, coverage_count_out))) return rc;
    if (hit_rank_out && (rc = prv_get_hit_rank(ctx, 0, V, hit_rank_out))) return rc;
    if (depth_out && (rc = prv_get_depth(ctx, 0, V, depth_out))) return rc;
    return prv_sync(ctx);
}

int prv_precept(prv_ctx* ctx, const double pose_world[16], const double init_pos[3], prv_point_xyzrgb* points_out, int* view_in_map_out) {
    if (!ctx || !points_out) return PRV_ERR_INVALID;
    int rc;
    ctx->h_view_ids.clear();
    ctx->ids_user = false;
    if ((rc = prv_set_views(ctx, pose_world, init_pos, 1))) return rc;
    if (view_in_map_out) *view_in_map_out = (ctx->h_views[0].flags & kViewInMap) ? 1 : 0;
    if ((rc = cast_impl(ctx, PRV_MODE_VOXEL, 0))) return rc;
    const uint32_t N = ctx->map.n_occ;
    if ((rc = ensure(ctx, ctx->d_voxel_hit, (size_t)N * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_points, (size_t)N * sizeof(prv_point_xyzrgb)))) return rc;
    {
        Span s(ctx, K_OTHER, 2);
        gather_voxel_hits_kernel<<<dim3((N + 255) / 256, 1), 256, 0, ctx->stream>>>(N, ptr<uint32_t>(ctx->d_voxel_pix), ptr<uint32_t>(ctx->d_pix_hit),
                                                                                   ctx->pix_stride, ptr<uint32_t>(ctx->d_voxel_hit));
        precept_points_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->map, ptr<uint32_t>(ctx->d_voxel_hit), ptr<prv_point_xyzrgb>(ctx->d_points));
    }
    CU(cudaGetLastError());
    CU(d2h(ctx, points_out, ctx->d_points.p, (size_t)N * sizeof(prv_point_xyzrgb)));
    CU(cudaStreamSynchronize(ctx->stream));
    return PRV_OK;
}

int prv_greedy(prv_ctx* ctx, uint32_t first_view, uint32_t max_iter, uint32_t* seq, uint32_t* gains, uint32_t capacity, uint32_t* n_out) {
    if (!ctx) return PRV_ERR_INVALID;
    int rc;
    if ((rc = prv_greedy_async(ctx, first_view, max_iter))) return rc;
    return prv_get_greedy(ctx, seq, gains, capacity, n_out, nullptr);
}

int prv_set_cloud(prv_ctx* ctx, const float* xyz, const uint8_t* rgb, uint64_t P) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!xyz || !rgb || P == 0 || P > 0xFFFFFFFEull) return fail(ctx, PRV_ERR_INVALID, "prv_set_cloud: null/empty cloud (or more than 2^32-2 points)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->d_cloud_xyz, P * 12))) return rc;
    if ((rc = ensure(ctx, ctx->d_cloud_rgb, P * 3))) return rc;
    CU(h2d(ctx, ctx->d_cloud_xyz.p, xyz, P * 12));
    CU(h2d(ctx, ctx->d_cloud_rgb.p, rgb, P * 3));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->P = P;
    return PRV_OK;
}

int prv_render_async(prv_ctx* ctx, uint32_t V, int point_size) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    return render_impl(ctx, V, point_size, true);
}

// view constants of a render-only call (poses without a map): only the inverse pose is needed
static int upload_render_views(prv_ctx* ctx, const double* pose_world, uint32_t V) {
    CU(join_score(ctx));
    PRV_GUARD_BEGIN
    std::vector<ViewConst> vcs(V);
    for (uint32_t v = 0; v < V; v++) {
        const prv::Matrix4d pw = prv::Matrix4d::FromRowMajor(pose_world + 16 * (size_t)v);
        const prv::Matrix4d inv = pw.inverse();
        std::memset(&vcs[v], 0, sizeof(ViewConst));
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) {
                vcs[v].pose[4 * r + c] = pw(r, c);
                vcs[v].inv[4 * r + c] = inv(r, c);
            }
        vcs[v].view_id = v;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->d_views, sizeof(ViewConst) * (size_t)V))) return rc;
    CU(h2d(ctx, ctx->d_views.p, vcs.data(), sizeof(ViewConst) * (size_t)V));
    CU(cudaStreamSynchronize(ctx->stream));
    PRV_GUARD_END(ctx, "prv_render_views")
    return PRV_OK;
}

int prv_render_views(prv_ctx* ctx, const double* pose_world, uint32_t V, int point_size, uint8_t* rgba_out, float* depth_out) {
    if (!ctx || !pose_world || !rgba_out) return PRV_ERR_INVALID;
    if (!ctx->have_cam) return fail(ctx, PRV_ERR_INVALID, "prv_render_views: camera not set");
    if (V == 0) return fail(ctx, PRV_ERR_INVALID, "prv_render_views: no views");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_render_views(ctx, pose_world, V))) return rc;
    ctx->V = V;
    rc = render_impl(ctx, V, point_size, depth_out != nullptr);
    ctx->V = 0;  // resident cast views were overwritten
    ctx->cast_done = false;
    if (rc) return rc;
    const size_t px = (size_t)V * ctx->cam.W * ctx->cam.H;
    CU(d2h(ctx, rgba_out, ctx->d_rgba.p, px * 4));
    if (depth_out) CU(d2h(ctx, depth_out, ctx->d_depth_img.p, px * 4));
    CU(cudaStreamSynchronize(ctx->stream));
    return PRV_OK;
}

int prv_object_pixel_rate(prv_ctx* ctx, const double* pose_world, uint32_t V, int point_size, double* rate_out, uint32_t* counts_out) {
    if (!ctx || !pose_world || !rate_out) return PRV_ERR_INVALID;
    if (!ctx->have_cam) return fail(ctx, PRV_ERR_INVALID, "prv_object_pixel_rate: camera not set");
    if (V == 0 || V > 65535) return fail(ctx, PRV_ERR_INVALID, "prv_object_pixel_rate: 1..65535 views");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_render_views(ctx, pose_world, V))) return rc;
    ctx->V = V;
    rc = render_impl(ctx, V, point_size, false);
    ctx->V = 0;
    ctx->cast_done = false;
    if (rc) return rc;
    if ((rc = ensure(ctx, ctx->d_counts, (size_t)V * 4))) return rc;
    CU(cudaMemsetAsync(ctx->d_counts.p, 0, (size_t)V * 4, ctx->stream));
    const uint32_t npix = (uint32_t)ctx->cam.W * (uint32_t)ctx->cam.H;
    {
        Span s(ctx, K_OTHER, 1);
        count_nonwhite_kernel<<<dim3(std::min<uint32_t>((npix + 255) / 256, 64u), V), 256, 0, ctx->stream>>>(ptr<uchar4>(ctx->d_rgba), npix, ptr<uint32_t>(ctx->d_counts));
    }
    CU(cudaGetLastError());
    PRV_GUARD_BEGIN
    std::vector<uint32_t> counts(V);
    CU(d2h_result(ctx, counts.data(), ctx->d_counts.p, (size_t)V * 4));
    double rate = 0.0;  // main.cpp:925-931: the mean over the views of count / (rows * cols)
    for (uint32_t v = 0; v < V; v++) rate += (double)counts[v] / ((double)ctx->cam.W * (double)ctx->cam.H);
    *rate_out = rate / (double)V;
    if (counts_out) std::memcpy(counts_out, counts.data(), (size_t)V * 4);
    PRV_GUARD_END(ctx, "prv_object_pixel_rate")
    return PRV_OK;
}

int prv_score_ensemble(prv_ctx* ctx, const uint8_t* images, uint32_t V, uint32_t E, int W, int H, int method, const uint8_t* chosen,
                       double* scores_out, int32_t* best_view_out) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!images || V == 0 || E == 0 || W <= 0 || H <= 0 || (method != 2 && method != 3))
        return fail(ctx, PRV_ERR_INVALID, "prv_score_ensemble: bad arguments (method must be 2 or 3)");
    CU(cudaSetDevice(ctx->device));
    const uint32_t npix = (uint32_t)W * (uint32_t)H;
    const size_t img_bytes = (size_t)V * E * npix * 4;
    int rc;
    if ((rc = ensure(ctx, ctx->d_ens_images, img_bytes))) return rc;
    if ((rc = ensure(ctx, ctx->d_ens_terms, (size_t)V * npix * 3 * 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_ens_scores, (size_t)V * 8 + 256 * 8))) return rc;
    CU(h2d(ctx, ctx->d_ens_images.p, images, img_bytes));
    const double* lut = nullptr;
    double host_lut[256];
    if (method == 2 && E == 2) {
        host_lut[0] = 0.0;  // never used: variance 0 is skipped
        for (int d = 1; d < 256; d++) host_lut[d] = std::log(((double)d * (double)d) / 4.0);  // ((a-m)^2 + (b-m)^2) / 2 with m = (a+b)/2
        CU(h2d(ctx, ptr<double>(ctx->d_ens_scores) + V, host_lut, sizeof(host_lut)));
        lut = ptr<double>(ctx->d_ens_scores) + V;
    }
    {
        Span s(ctx, K_OTHER, 2);
        ensemble_terms_kernel<<<dim3((npix + 255) / 256, V), 256, 0, ctx->stream>>>(ptr<uint8_t>(ctx->d_ens_images), E, npix, method, lut,
                                                                                  ptr<double>(ctx->d_ens_terms));
        ensemble_sum_kernel<<<(V + 31) / 32, 32, 0, ctx->stream>>>(ptr<double>(ctx->d_ens_terms), V, npix, ptr<double>(ctx->d_ens_scores));
    }
    CU(cudaGetLastError());
    PRV_GUARD_BEGIN
    std::vector<double> scores(V);
    CU(d2h(ctx, scores.data(), ctx->d_ens_scores.p, (size_t)V * 8));
    CU(cudaStreamSynchronize(ctx->stream));
    double largest = -1e100;  // main.cpp:1971
    int32_t best = -1;
    for (uint32_t i = 0; i < V; i++) {
        if (chosen && chosen[i]) {
            scores[i] = 0.0;
            continue;
        }
        if (scores[i] > largest) {  // strict '>': first maximum wins (main.cpp:2088, :2152)
            largest = scores[i];
            best = (int32_t)i;
        }
    }
    if (scores_out) std::memcpy(scores_out, scores.data(), (size_t)V * 8);
    if (best_view_out) *best_view_out = best;
    PRV_GUARD_END(ctx, "prv_score_ensemble")
    return PRV_OK;
}

int prv_timing_reset(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(sync_all(ctx));
    PRV_GUARD_BEGIN
    for (auto& s : ctx->spans) {
        ctx->free_events.push_back(s.a);
        ctx->free_events.push_back(s.b);
    }
    ctx->spans.clear();
    ctx->spans_dropped = 0;
    PRV_GUARD_END(ctx, "prv_timing_reset")
    return PRV_OK;
}

int prv_get_timing(prv_ctx* ctx, prv_timing* out) {
    if (!ctx || !out) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(sync_all(ctx));
    float ms[K_NCLASS] = {0};
    uint32_t cnt[K_NCLASS] = {0};
    for (auto& s : ctx->spans) {
        float t = 0.0f;
        CU(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.cls] += t;
        cnt[s.cls] += s.launches;
    }
    out->cast_ms = ms[K_CAST] + ms[K_CULL] + ms[K_MARCH];
    out->cast_launches = cnt[K_CAST] + cnt[K_CULL] + cnt[K_MARCH];
    out->cull_ms = ms[K_CULL];       out->cull_launches = cnt[K_CULL];
    out->march_ms = ms[K_MARCH];     out->march_launches = cnt[K_MARCH];
    out->project_ms = ms[K_PROJECT]; out->project_launches = cnt[K_PROJECT];
    out->count_ms = ms[K_COUNT];     out->count_launches = cnt[K_COUNT];
    out->greedy_ms = ms[K_GREEDY];   out->greedy_launches = cnt[K_GREEDY];
    out->splat_ms = ms[K_SPLAT];     out->splat_launches = cnt[K_SPLAT];
    out->resolve_ms = ms[K_RESOLVE]; out->resolve_launches = cnt[K_RESOLVE];
    out->other_ms = ms[K_OTHER];     out->other_launches = cnt[K_OTHER];
    out->gather_ms = ms[K_GATHER];   out->gather_launches = cnt[K_GATHER];
    out->flush_ms = ms[K_FLUSH];
    out->dropped = ctx->spans_dropped;
    return PRV_OK;
}

int prv_event_record(prv_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(join_score(ctx));  // the event marks the end of everything enqueued so far, on both streams
    CU(cudaEventRecord(ctx->slots[slot], ctx->stream));
    return PRV_OK;
}

int prv_event_elapsed_ms(prv_ctx* ctx, int a, int b, float* ms_out) {
    if (!ctx || !ms_out || a < 0 || a >= 16 || b < 0 || b >= 16) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->slots[b]));
    CU(cudaEventElapsedTime(ms_out, ctx->slots[a], ctx->slots[b]));
    return PRV_OK;
}

int prv_reset_counters(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    ctx->n_launches = ctx->h2d_bytes = ctx->d2h_bytes = 0;
    return PRV_OK;
}

int prv_get_counters(prv_ctx* ctx, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!ctx) return PRV_ERR_INVALID;
    if (kernel_launches) *kernel_launches = ctx->n_launches;
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    return PRV_OK;
}

int prv_map_bytes(prv_ctx* ctx, uint64_t* bitmap_bytes) {
    if (!ctx || !ctx->have_map || !bitmap_bytes) return PRV_ERR_INVALID;
    *bitmap_bytes = (uint64_t)ctx->map.wx * ctx->map.n[1] * ctx->map.n[2] * 4;
    return PRV_OK;
}

int prv_flush_l2(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)256 << 20;  // > 126 MB L2
    int rc;
    if ((rc = ensure(ctx, ctx->d_flush, bytes))) return rc;
    {
        Span s(ctx, K_FLUSH, 0);  // timed on its own (prv_timing::flush_ms) so a caller can take it out of a bracket around several steps
        CU(cudaMemsetAsync(ctx->d_flush.p, 0x5A, bytes, ctx->stream));
    }
    // whatever is enqueued on the scoring stream from here on starts after the flush too: nothing of the path runs underneath it
    CU(cudaEventRecord(ctx->ev_flush, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->score_stream, ctx->ev_flush, 0));
    return PRV_OK;
}

// ---------------------------------------------------------------- NCCL (dlopen)
static int load_nccl(prv_ctx* ctx, NcclApi& api) {
    if (api.handle) return PRV_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nme : names) {
        api.handle = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return fail(ctx, PRV_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    api.GetUniqueId = (int (*)(void*))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(api.handle, "ncclCommInitRank");
    api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllGather");
    api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) return fail(ctx, PRV_ERR_NCCL, "libnccl is missing required symbols");
    return PRV_OK;
}

int prv_comm_unique_id(void* id_out_128) {
    if (!id_out_128) return PRV_ERR_INVALID;
    NcclApi api;
    int rc = load_nccl(nullptr, api);
    if (rc) return rc;
    const int r = api.GetUniqueId(id_out_128);
    return r == 0 ? PRV_OK : fail(nullptr, PRV_ERR_NCCL, "ncclGetUniqueId failed: %d", r);
}

int prv_comm_init(prv_ctx* ctx, const void* id_128, int rank, int nranks) {
    if (!ctx || !id_128 || nranks < 1 || rank < 0 || rank >= nranks) return PRV_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = load_nccl(ctx, ctx->nccl);
    if (rc) return rc;
    Id128 id;
    std::memcpy(id.b, id_128, 128);
    const int r = ctx->nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != 0) return fail(ctx, PRV_ERR_NCCL, "ncclCommInitRank failed: %s", ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(r) : "?");
    ctx->rank = rank;
    ctx->nranks = nranks;
    return PRV_OK;
}

// receive side of the peer-memory exchange (+ the send side when the resident cast did not publish): scoring stream
static int p2p_gather(prv_ctx* ctx) {
    const uint32_t V = ctx->V, words = ctx->map.words64, G = (uint32_t)ctx->nranks;
    cudaStream_t qs = ctx->score_stream;
    CU(cudaStreamWaitEvent(qs, ctx->ev_rows_ready[ctx->cur], 0));
    PeerArena pa;
    if (!ctx->cur_published) {
        if (!peer_arena_for(ctx, ctx->p2p_step + 1, V, words, pa))
            return fail(ctx, PRV_ERR_INVALID, "prv_allgather_bitsets: the gathered table (%u x %u rows of %u words) does not fit the peer arena", G, V, words);
        ctx->p2p_step++;
        // (counts go to a scratch table: d_counts may be being rewritten by the next cast on the other stream)
        int rc;
        if ((rc = ensure(ctx, ctx->d_counts_scratch, (size_t)V * 4))) return rc;
        Span s(ctx, K_GATHER, 1, qs);
        count_publish_kernel<<<V, 256, 0, qs>>>(ptr<uint64_t>(ctx->d_rows[ctx->cur]), words, V, ptr<uint32_t>(ctx->d_counts_scratch), ptr<uint32_t>(ctx->d_view_ids), pa,
                                                ctx->p2p_step, ptr<uint32_t>(ctx->d_done) + 1);
        ctx->cur_published = true;
    } else {
        peer_arena_for(ctx, ctx->p2p_step, V, words, pa);
    }
    {
        Span s(ctx, K_GATHER, 1, qs);
        wait_published_kernel<<<1, 32, 0, qs>>>(arena_flags(ctx->d_arena.p), G, ctx->p2p_step, arena_error(ctx->d_arena.p));
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_rows_free[ctx->cur], qs));
    ctx->rows_free_pending[ctx->cur] = true;
    ctx->g_rows = pa.rows[ctx->rank];
    ctx->g_ids = pa.ids[ctx->rank];
    return PRV_OK;
}

int prv_allgather_bitsets_async(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    if (!ctx->cast_done) return fail(ctx, PRV_ERR_INVALID, "prv_allgather_bitsets: nothing cast yet");
    if (!ctx->comm && !ctx->p2p_ready) return fail(ctx, PRV_ERR_INVALID, "prv_allgather_bitsets: neither prv_comm_init nor prv_comm_p2p_import has been called");
    CU(cudaSetDevice(ctx->device));
    const uint32_t V = ctx->V, words = ctx->map.words64, G = (uint32_t)ctx->nranks;
    int rc;
    cudaStream_t qs = ctx->score_stream;
    if (ctx->p2p_ready) {
        if ((rc = p2p_gather(ctx))) return rc;
    } else {
        if ((rc = ensure(ctx, ctx->d_all_rows, (size_t)G * V * words * 8))) return rc;
        if ((rc = ensure(ctx, ctx->d_all_ids, (size_t)G * V * 4))) return rc;
        CU(cudaStreamWaitEvent(qs, ctx->ev_rows_ready[ctx->cur], 0));
        {
            Span s(ctx, K_GATHER, 1, qs);
            const int r = ctx->nccl.AllGather(ctx->d_rows[ctx->cur].p, ctx->d_all_rows.p, (size_t)V * words, /*ncclUint64*/ 5, ctx->comm, qs);
            if (r != 0) return fail(ctx, PRV_ERR_NCCL, "ncclAllGather failed: %s", ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(r) : "?");
        }
        CU(cudaEventRecord(ctx->ev_rows_free[ctx->cur], qs));  // the local rows have been sent: the cast after next may reuse the buffer
        ctx->rows_free_pending[ctx->cur] = true;
        ctx->g_rows = ptr<uint64_t>(ctx->d_all_rows);
        ctx->g_ids = ptr<uint32_t>(ctx->d_all_ids);
    }
    if (!ctx->gather_ids_valid) {
        // once per view set: the global view ids of the gathered rows and view id -> row of the gathered table (the per-step
        // exchange of the rows above stays fully asynchronous)
        PRV_GUARD_BEGIN
        if (!ctx->p2p_ready) {
            const int r = ctx->nccl.AllGather(ctx->d_view_ids.p, ctx->d_all_ids.p, (size_t)V, /*ncclUint32*/ 3, ctx->comm, qs);
            if (r != 0) return fail(ctx, PRV_ERR_NCCL, "ncclAllGather failed: %s", ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(r) : "?");
        }
        std::vector<uint32_t> ids((size_t)G * V);
        ctx->d2h_bytes += ids.size() * 4;
        CU(cudaMemcpyAsync(ids.data(), ctx->g_ids, ids.size() * 4, cudaMemcpyDeviceToHost, qs));
        CU(cudaStreamSynchronize(qs));
        if (ctx->p2p_ready) {
            uint32_t err = 0;
            CU(cudaMemcpy(&err, arena_error(ctx->d_arena.p), 4, cudaMemcpyDeviceToHost));
            if (err) return fail(ctx, PRV_ERR_NCCL, "prv_allgather_bitsets: rank %u never published its rows (peer exchange timed out)", err - 1);
        }
        uint32_t max_id = 0;
        for (uint32_t id : ids) {
            if (id >= kMaxViewId) return fail(ctx, PRV_ERR_INVALID, "prv_allgather_bitsets: a rank sent view id %u", id);
            max_id = std::max(max_id, id);
        }
        std::vector<uint32_t>& row_of_id = ctx->h_row_of_id_all;
        row_of_id.assign((size_t)max_id + 1, kNone);
        for (size_t r2 = 0; r2 < ids.size(); r2++) {
            // (ranks pad their shard with ids >= the number of real views: every id must still be unique across ranks)
            if (row_of_id[ids[r2]] != kNone) return fail(ctx, PRV_ERR_INVALID, "prv_allgather_bitsets: view id %u is held by two ranks", ids[r2]);
            row_of_id[ids[r2]] = (uint32_t)r2;
        }
        if ((rc = ensure(ctx, ctx->d_row_of_id_all, 4 * row_of_id.size()))) return rc;
        ctx->h2d_bytes += 4 * row_of_id.size();
        CU(cudaMemcpyAsync(ctx->d_row_of_id_all.p, row_of_id.data(), 4 * row_of_id.size(), cudaMemcpyHostToDevice, qs));
        CU(cudaStreamSynchronize(qs));
        ctx->gather_ids_valid = true;
        PRV_GUARD_END(ctx, "prv_allgather_bitsets")
    }
    ctx->gathered = true;
    ctx->g_nrows = G * V;
    return PRV_OK;
}

// Peer-memory exchange: every rank allocates an arena, hands out its cudaIpc handle (prv_comm_p2p_export), receives everybody's
// (exchanged by the caller: MPI, torch.distributed, a file ...) and maps them (prv_comm_p2p_import).  From then on
// prv_allgather_bitsets_async uses stores over NVLink + flags instead of ncclAllGather, and prv_cast_async(PRV_CAST_PUBLISH)
// fuses the send side into the coverage-count kernel.
int prv_comm_p2p_export(prv_ctx* ctx, void* handle_out_64, uint64_t table_bytes_max) {
    if (!ctx || !handle_out_64) return PRV_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(cudaSetDevice(ctx->device));
    if (table_bytes_max == 0) table_bytes_max = (uint64_t)16 << 20;
    const size_t buf = ((size_t)table_bytes_max + 255) & ~(size_t)255;
    int rc;
    if ((rc = ensure(ctx, ctx->d_arena, kArenaHeader + kPeerBuffers * buf))) return rc;
    if ((rc = ensure(ctx, ctx->d_done, 16))) return rc;
    CU(cudaMemset(ctx->d_arena.p, 0, kArenaHeader));
    CU(cudaMemset(ctx->d_done.p, 0, 16));
    ctx->arena_buf_bytes = buf;
    ctx->arena_bytes = kArenaHeader + kPeerBuffers * buf;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_arena.p));
    std::memcpy(handle_out_64, &h, 64);
    return PRV_OK;
}

int prv_comm_p2p_import(prv_ctx* ctx, const void* handles, int rank, int nranks) {
    if (!ctx || !handles || nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks) return PRV_ERR_INVALID;
    if (!ctx->d_arena.p) return fail(ctx, PRV_ERR_INVALID, "prv_comm_p2p_import: call prv_comm_p2p_export first");
    CU(cudaSetDevice(ctx->device));
    for (int r = 0; r < nranks; r++) {
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, reinterpret_cast<const char*>(handles) + 64 * (size_t)r, 64);
        const cudaError_t e = cudaIpcOpenMemHandle(&ctx->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, PRV_ERR_NCCL, "prv_comm_p2p_import: cannot map rank %d's arena: %s", r, cudaGetErrorString(e));
        }
        ctx->peer_opened[r] = true;
    }
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->p2p_step = 0;
    ctx->p2p_ready = true;
    return PRV_OK;
}

int prv_get_gathered(prv_ctx* ctx, uint64_t* rows_out, uint32_t* ids_out, uint32_t* nrows_out) {
    if (!ctx || !nrows_out) return PRV_ERR_INVALID;
    if (!ctx->cast_done || !ctx->gathered) return fail(ctx, PRV_ERR_INVALID, "prv_get_gathered: prv_allgather_bitsets_async has not run for the resident cast");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->score_stream));
    *nrows_out = ctx->g_nrows;
    if (rows_out) CU(d2h_result(ctx, rows_out, ctx->g_rows, (size_t)ctx->g_nrows * ctx->map.words64 * 8));
    if (ids_out) CU(d2h_result(ctx, ids_out, ctx->g_ids, (size_t)ctx->g_nrows * 4));
    if (ctx->p2p_ready) {
        uint32_t err = 0;
        CU(cudaMemcpy(&err, arena_error(ctx->d_arena.p), 4, cudaMemcpyDeviceToHost));
        if (err) return fail(ctx, PRV_ERR_NCCL, "prv_get_gathered: rank %u never published its rows (peer exchange timed out)", err - 1);
    }
    return PRV_OK;
}

int prv_comm_p2p_close(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    for (int r = 0; r < kMaxPeers; r++)
        if (ctx->peer_opened[r]) {
            cudaIpcCloseMemHandle(ctx->peer_base[r]);
            ctx->peer_opened[r] = false;
        }
    ctx->p2p_ready = false;
    ctx->cur_published = false;
    return PRV_OK;
}

int prv_comm_destroy(prv_ctx* ctx) {
    if (!ctx) return PRV_ERR_INVALID;
    prv_comm_p2p_close(ctx);
    if (ctx->comm && ctx->nccl.CommDestroy) {
        ctx->nccl.CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    return PRV_OK;
}

}  // extern "C"
