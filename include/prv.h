/*
 * prv.h -- C ABI of libprv_b200.so: the B200 (sm_100a) implementation of NeRF-PRV's PRV_simulation
 * ray-cast / coverage / greedy-selection / splat-render hot path.
 *
 * The reference (psc0628/NeRF-PRV) has no FFI/plugin layer: its boundary is the C++ class surface
 * Share_Data / View / View_Space / Perception_3D (PRV_simulation/Share_Data.hpp, View_Space.hpp,
 * main.cpp:17-286).  This header is what those classes bind instead of OctoMap / PCL-VTK; each entry
 * point cites the reference code it replaces.  The host-side mirror of the classes lives in
 * nerf-prv_b200/host/ and INTEGRATION.md shows the binding a maintainer adds to the reference.
 *
 * Conventions
 *   - extern "C", POD only, caller owns every host buffer, ctx owns device memory / stream / events.
 *   - every function returns PRV_OK (0) or a negative prv_status; nothing throws or aborts across the
 *     ABI; prv_last_error(ctx) returns a static/ctx-owned message for the last failure.
 *   - calls are synchronous at return unless named *_async: host INPUT buffers have been consumed and host
 *     OUTPUT buffers are filled when a call returns.  Device work that only produces device-resident state
 *     (the lookup tables built by prv_set_map, the view table of prv_set_views) may still be running on the
 *     ctx stream at return; every later call is ordered behind it on that stream, and a CUDA error of such
 *     work is reported by the next call that synchronises.  One ctx per GPU; a ctx is not re-entrant.
 *   - small results (coverage rows, counts, greedy sequence) are copied with a direct device-to-host DMA
 *     when the caller's buffer is pinned / registered host memory, else through a ctx-owned pinned buffer.
 *   - there is NO CPU fallback: without a CUDA device prv_create fails with PRV_ERR_NO_DEVICE.
 *   - matrices are row-major double[16]; "pose_world" is the reference's view_pose_world
 *     (= now_camera_pose_world * view.pose.inverse(), main.cpp:72,109).
 *   - "leaf order" is ColorOcTree::begin_leafs() order (ascending Morton code, z most significant in
 *     each bit triple); it is the index i of cloud->points[i] (main.cpp:117-121) and the bit index of
 *     the coverage bitsets.
 */
#ifndef PRV_B200_H
#define PRV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRV_ABI_VERSION 2
#define PRV_NONE 0xFFFFFFFFu /* "no voxel" in hit-rank tables */

typedef enum prv_status {
    PRV_OK = 0,
    PRV_ERR_INVALID = -1,     /* bad argument / call order                       */
    PRV_ERR_NO_DEVICE = -2,   /* no CUDA device / wrong architecture             */
    PRV_ERR_CUDA = -3,        /* CUDA runtime error (see prv_last_error)         */
    PRV_ERR_OOM = -4,         /* host or device allocation failed                */
    PRV_ERR_UNSUPPORTED = -5, /* e.g. deprojecting distortion model 1, maps too large for the dense bitmap */
    PRV_ERR_NCCL = -6,        /* NCCL not loadable / collective failed           */
    PRV_ERR_IO = -7           /* file output failed                              */
} prv_status;

/* Same layout as rs2_intrinsics (Share_Data.hpp:79-89); model is the rs2_distortion enum (:67-76).
 * coeffs are used BY INDEX exactly as rs2_project/deproject do (Share_Data.hpp:101-105,150-152). */
typedef struct prv_intrinsics {
    int   width;
    int   height;
    float ppx;
    float ppy;
    float fx;
    float fy;
    int   model;
    float coeffs[5];
} prv_intrinsics;

/* Memory image of pcl::PointXYZRGB (32 bytes: xyz + 1.0f, bgra, 12 pad bytes) so prv_precept can
 * write straight into cloud->points.data() (main.cpp:105,249,283). */
typedef struct prv_point_xyzrgb {
    float   x, y, z, w;
    uint8_t b, g, r, a;
    float   pad[3];
} prv_point_xyzrgb;

/* cast modes */
#define PRV_MODE_VOXEL 0 /* precept-exact: one ray per occupied voxel through its truncated pixel (main.cpp:238-284) */
#define PRV_MODE_DENSE 1 /* one ray per integer pixel (x,y) in [0,W)x[0,H)                                        */

/* ray-march kernel variants (all must give identical results; tests enforce it) */
#define PRV_VARIANT_PLAIN 0 /* literal sequential castRay march incl. max-range test every step           */
#define PRV_VARIANT_FAST 1  /* + exact AABB pre-cull and exact early exit                                   */
#define PRV_VARIANT_AXIS 2  /* + exact per-axis approach march (default)                                    */

typedef struct prv_ctx prv_ctx;

/* deterministic per-cast counters (oracle prints the same numbers; roofline numerators) */
typedef struct prv_cast_stats {
    uint64_t rays;       /* rays cast (dense: V*W*H; voxel: set pixels)            */
    uint64_t probes_in;  /* DDA steps whose key lies inside the occupancy AABB     */
    uint64_t hits;       /* rays with a first-hit voxel                            */
    uint64_t steps;      /* DDA steps actually executed by the chosen variant      */
    uint64_t marched;    /* rays that survived the conservative culls and were marched */
} prv_cast_stats;

/* per-kernel-class device time accumulated with CUDA events on the ctx stream since prv_timing_reset */
typedef struct prv_timing {
    float    cast_ms;     uint32_t cast_launches;    /* single-kernel PLAIN/FAST cast, or cull+march together */
    float    cull_ms;     uint32_t cull_launches;    /* AXIS pipeline, kernel 1 */
    float    march_ms;    uint32_t march_launches;   /* AXIS pipeline, kernel 2 (the dominant kernel) */
    float    project_ms;  uint32_t project_launches;
    float    count_ms;    uint32_t count_launches;
    float    greedy_ms;   uint32_t greedy_launches;
    float    splat_ms;    uint32_t splat_launches;
    float    resolve_ms;  uint32_t resolve_launches;
    float    other_ms;    uint32_t other_launches;
    float    gather_ms;   uint32_t gather_launches;  /* NCCL all-gather of the coverage rows (scoring stream) */
    float    flush_ms;    /* prv_flush_l2 memsets (not work of the path: benchmarks subtract it) */
    uint32_t dropped;     /* spans not recorded because 65536 were already held: call prv_timing_reset more often */
} prv_timing;

/* ---------------------------------------------------------------- lifecycle */
int         prv_abi_version(void);
int         prv_create(prv_ctx** ctx, int device);
void        prv_destroy(prv_ctx* ctx);
const char* prv_last_error(const prv_ctx* ctx); /* ctx may be NULL: last create error */
int         prv_device_info(prv_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, uint64_t* mem_bytes);
int         prv_sync(prv_ctx* ctx);
int         prv_set_variant(prv_ctx* ctx, int variant);
/* Tuning of the conservative brick cull in front of the exact march (variant AXIS).  `cell` = edge of the bricks of the
 * coarse occupancy grid in voxels (4, 8 or 16; 0 = chosen by map size, the default: 4 while the AABB spans at most 64 voxels,
 * else 8).  With enter_at_brick != 0 (default) the exact march of a
 * surviving ray starts where the ray enters the first set brick of the cull's walk (grown by one voxel) instead of at the
 * AABB face: the voxels in between are provably empty and are crossed by the cheap per-axis additions.  Results are
 * identical for every setting; only the probes_in / marched counters of prv_cast_stats move.
 * Takes effect at the next prv_set_map / prv_set_map_from_cloud. */
int         prv_set_brick_cull(prv_ctx* ctx, int cell, int enter_at_brick);
/* Where the occupancy bitmap is staged for the exact march (results identical).  bitmap_in_shared_memory (default on): the
 * shell-padded bitmap is copied into every march block's shared memory when it is at most 44 KB (the 0.002 m maps: -1.7 % march
 * time on C3); larger maps are read through the read-only L1 path.  l2_persisting_window (default off: measured no effect, the
 * map is L2-resident anyway): an L2 persisting access-policy window over the bitmap, taking effect at the next prv_set_map.
 * Numbers: profiles/r2_staging_ab.md. */
int         prv_set_staging(prv_ctx* ctx, int bitmap_in_shared_memory, int l2_persisting_window);

/* ---------------------------------------------------------------- host-side logic (pure host, no device)
 * One implementation of the reference's pose / view-space / map-insertion arithmetic for every caller. */
/* Eigen::Matrix4d::inverse() as used at main.cpp:72,109,244 and View_Space.hpp:73-137 */
int prv_host_mat4_inverse(const double m[16], double out[16]);
/* View::get_next_camera_pos(now_camera_pose_world, object_center_world, 0): View_Space.hpp:67-140 */
int prv_host_view_pose(const double now_camera_pose_world[16], const double init_pos[3],
                       const double object_center_world[3], double pose_out[16]);
/* view_pose_world = now_camera_pose_world * pose.inverse(): main.cpp:72,109 */
int prv_host_view_pose_world(const double now_camera_pose_world[16], const double pose[16], double out[16]);
/* View_Space::get_view_space: View_Space.hpp:517-558.  init_pos_out has room for N rows; *n_views_out <= N. */
int prv_host_view_space(const float* pts_xyz, uint64_t P, const double* pt_sphere, int N, double pt_norm,
                        double view_space_radius, double center_out[3], double* predicted_size_out,
                        double* init_pos_out, int* n_views_out);
/* cloud normalisation of the NBV_Net_Labeler ctor: main.cpp:674 (toward pose 4), :768-790, :800-832, :1008-1010 */
int prv_host_normalize_cloud(float* pts_xyz, uint64_t P, double target_size, double* predicted_size_before_out);
/* ground_truth_model insertion loop + leaf enumeration: main.cpp:1005-1036, 1055-1058, 116-121.
 * keys_out/rgb_out need room for P entries; *n_out = full_voxels. */
int prv_host_build_map(const float* pts_xyz, const uint8_t* rgb, uint64_t P, double resolution,
                       uint16_t* keys_out, uint8_t* rgb_out, uint32_t* n_out);

/* rs2_project_point_to_pixel / rs2_deproject_pixel_to_point (Share_Data.hpp:92-137, 140-196), all six distortion models, float
 * arithmetic in the reference's order; tan / atan of models 3 and 5 are this host's float libm, as in the reference built on
 * this host.  The library itself uses them for those two models (per-pixel deprojection table, host-side voxel projection). */
int prv_host_project_point_to_pixel(const prv_intrinsics* intr, const float point[3], float pixel_out[2]);
int prv_host_deproject_pixel_to_point(const prv_intrinsics* intr, const float pixel[2], float depth, float point_out[3]);

/* Leaf-order check of a key table: keys must be in strictly ascending begin_leafs() (Morton) order, as prv_set_map requires
 * (main.cpp:116-121).  *first_bad_out = index of the first key that is not above its predecessor, or N when the table is in
 * order.  (prv_set_map runs this itself; BMI2 pdep when the CPU has it.) */
int prv_host_check_leaf_order(const uint16_t* keys /* N x 3 */, uint32_t N, uint32_t* first_bad_out);

/* ---------------------------------------------------------------- device inputs */
/* ground_truth_model (Share_Data.hpp:258,458): occupied leaf keys in leaf order + voxel colours (may be NULL). */
int prv_set_map(prv_ctx* ctx, const uint16_t* keys /* N x 3 */, const uint8_t* rgb /* N x 3 */, uint32_t N,
                double resolution);
/* GPU ingest: the same insertion rule as prv_host_build_map (main.cpp:1005-1036: key per point, a voxel keeps the colour
 * of its FIRST point; leaf order) executed on the device from the normalised cloud_ground_truth points, then prv_set_map.
 * rgb may be NULL.  prv_get_map returns the resulting leaf keys / colours (prv_full_voxels() entries). */
int prv_set_map_from_cloud(prv_ctx* ctx, const float* xyz /* P x 3 */, const uint8_t* rgb /* P x 3 */, uint64_t P, double resolution);
int prv_get_map(prv_ctx* ctx, uint16_t* keys_out /* N x 3 or NULL */, uint8_t* rgb_out /* N x 3 or NULL */);
/* share_data->color_intrinsics (Share_Data.hpp:388-399) and castRay's maxRange (main.cpp:258: 1.0) */
int prv_set_camera(prv_ctx* ctx, const prv_intrinsics* intr, double max_range);
/* candidate views: view_pose_world (main.cpp:109) and View::init_pos (snapped to a voxel centre on upload, main.cpp:112-114) */
int prv_set_views(prv_ctx* ctx, const double* pose_world /* V x 16 */, const double* init_pos /* V x 3 */, uint32_t V);
/* multi-GPU view sharding: global id of each local view (ties in the greedy argmax break on this id).  ids must be distinct
 * and below 2^24; the next prv_set_views must pass the same V.  NULL (or V = 0) = 0..V-1. */
int prv_set_view_ids(prv_ctx* ctx, const uint32_t* ids, uint32_t V);

uint32_t prv_full_voxels(const prv_ctx* ctx);  /* share_data->full_voxels (main.cpp:1055-1058) */
uint32_t prv_bitset_words(const prv_ctx* ctx); /* u64 words per coverage row (padded to 16 B)   */
uint32_t prv_num_views(const prv_ctx* ctx);

/* ---------------------------------------------------------------- resident path (inputs already in HBM)
 * Kernels are enqueued and the call returns without synchronising.  The ctx owns two streams: the cast runs on one, the
 * consumers of its coverage rows (prv_allgather_bitsets_async, prv_greedy_async) on the other, with the rows double-
 * buffered -- a caller that enqueues cast / [all-gather] / greedy for several steps back to back gets the scoring of step k
 * overlapped with the cull and march of step k+1.  Every prv_get_* / prv_sync / prv_event_record joins both streams. */
#define PRV_CAST_PIXELS 1  /* also write the per-pixel hit rank + depth tables */
#define PRV_CAST_PUBLISH 2 /* multi-GPU (after prv_comm_p2p_import): the coverage-count kernel also stores every row into every
                              rank's gathered table over NVLink peer memory -- the send side of the all-gather fused into the
                              kernel that reads the rows anyway; EVERY rank must then call prv_allgather_bitsets_async */
int prv_cast_async(prv_ctx* ctx, int mode, int flags /* PRV_CAST_* */);
int prv_greedy_async(prv_ctx* ctx, uint32_t first_view, uint32_t max_iter);

/* ---------------------------------------------------------------- results (synchronise, then copy D2H) */
int prv_get_bitsets(prv_ctx* ctx, uint64_t* out /* V x words */);
int prv_get_coverage_counts(prv_ctx* ctx, uint32_t* out /* V */);
/* dense: [V][H][W]; voxel mode: per-voxel [V][full_voxels] (rank seen by voxel i's ray) */
int prv_get_hit_rank(prv_ctx* ctx, uint32_t view_begin, uint32_t view_count, uint32_t* out);
int prv_get_depth(prv_ctx* ctx, uint32_t view_begin, uint32_t view_count, float* out); /* dense only */
/* seq / gains hold `capacity` entries; *n_out = entries selected (<= max_iter + 1 of the prv_greedy_async call).  If that is
 * more than `capacity` nothing is written and PRV_ERR_INVALID is returned (*n_out still tells the size needed). */
int prv_get_greedy(prv_ctx* ctx, uint32_t* seq, uint32_t* gains, uint32_t capacity, uint32_t* n_out, uint64_t* covered_out /* words, may be NULL */);
/* which implementation the last prv_greedy_async launched: 0 = one thread-block cluster (table in distributed shared memory),
 * 1 = cooperative grid-barrier kernel, 2 = one launch per iteration; -1 without a ctx */
int prv_greedy_path(const prv_ctx* ctx);
int prv_get_cast_stats(prv_ctx* ctx, prv_cast_stats* out);

/* ---------------------------------------------------------------- host-buffer one-call API (what the classes bind) */
/* Replaces the per-view loop over Perception_3D::precept (main.cpp:98-236, callers :1453,:1523,:1607).
 * Any output pointer may be NULL.  Leaves the bitsets resident for prv_greedy. */
int prv_cast_views(prv_ctx* ctx, const double* pose_world, const double* init_pos, uint32_t V, int mode,
                   uint64_t* bitsets_out, uint32_t* coverage_count_out, uint32_t* hit_rank_out, float* depth_out);
/* Perception_3D::precept for one view: fills cloud->points[0..full_voxels) (main.cpp:105,238-284).
 * Returns PRV_OK; *view_in_map_out = 0 reproduces "View out of map.check." (main.cpp:139). */
int prv_precept(prv_ctx* ctx, const double pose_world[16], const double init_pos[3], prv_point_xyzrgb* points_out,
                int* view_in_map_out);
/* greedy set-cover over the resident bitsets (frozen definition; tie rule of main.cpp:2006,2088,2152).
 * seq / gains hold `capacity` entries (max_iter + 1 always suffices; fewer is an error only if more get selected). */
int prv_greedy(prv_ctx* ctx, uint32_t first_view, uint32_t max_iter, uint32_t* seq, uint32_t* gains, uint32_t capacity, uint32_t* n_out);

/* ---------------------------------------------------------------- splat z-buffer render (replaces Perception_3D::render, main.cpp:68-96) */
/* share_data->cloud_ground_truth (main.cpp:38) */
int prv_set_cloud(prv_ctx* ctx, const float* xyz /* P x 3 */, const uint8_t* rgb /* P x 3 */, uint64_t P);
/* Renders V views; rgba_out [V][H][W][4] already has white->alpha0 (Share_Data.hpp:771-784) and the
 * 180-degree flip (main.cpp:1616) applied, i.e. it is the pixel content of rgbaClip_<i>.png.  depth_out [V][H][W] may be NULL. */
int prv_render_views(prv_ctx* ctx, const double* pose_world, uint32_t V, int point_size, uint8_t* rgba_out,
                     float* depth_out);
/* Size-augmentation probe of the NBV_Net_Labeler constructor (main.cpp:873-938): renders the V test views like
 * prv_render_views, counts on the device the pixels that are not (255,255,255) and returns the mean over the views of
 * count / (W*H) (main.cpp:913-931: the value compared with object_pixel_rate).  Nothing but V counters leaves the device.
 * counts_out [V] may be NULL. */
int prv_object_pixel_rate(prv_ctx* ctx, const double* pose_world, uint32_t V, int point_size, double* rate_out, uint32_t* counts_out);
int prv_render_async(prv_ctx* ctx, uint32_t V, int point_size); /* resident: uses prv_set_views poses, output stays on device */
float prv_splat_focal(const prv_intrinsics* intr);

/* ---------------------------------------------------------------- ensemble-uncertainty view scoring
 * NBV_Net_Labeler::nbv_loop cases 2 (EnsembleRGB) and 3 (EnsembleRGBDensity), main.cpp:2039-2161: per candidate view the
 * per-pixel variance of `E` ensemble renders is accumulated into one score and the arg-max view is returned (strict '>'
 * from -1e100, main.cpp:1971,2088,2152; `chosen[i] != 0` skips view i like chosen_nbvs_set, may be NULL).
 * images: [V][E][H][W][4] uint8 in the channel order cv::imread(IMREAD_UNCHANGED) yields (0..2 colour, 3 alpha).
 * The per-pixel terms are computed in parallel and summed per view in the reference's pixel order, so method 3 is
 * bit-exact; method 2 is bit-exact for E == 2 (the value the reference uses, Share_Data.hpp:505-507: log terms come from
 * a host-computed table), otherwise within 1e-12 relative (CUDA log). */
int prv_score_ensemble(prv_ctx* ctx, const uint8_t* images, uint32_t V, uint32_t E, int W, int H, int method,
                       const uint8_t* chosen, double* scores_out /* V */, int32_t* best_view_out);

/* ---------------------------------------------------------------- timing (CUDA events on the ctx stream) */
int prv_timing_reset(prv_ctx* ctx);
int prv_get_timing(prv_ctx* ctx, prv_timing* out); /* synchronises */
int prv_event_record(prv_ctx* ctx, int slot /* 0..15 */);
int prv_event_elapsed_ms(prv_ctx* ctx, int slot_begin, int slot_end, float* ms_out); /* synchronises on slot_end */
/* counters since prv_reset_counters: kernels of this library launched, bytes copied host->device / device->host */
int prv_reset_counters(prv_ctx* ctx);
int prv_get_counters(prv_ctx* ctx, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* size of the dense occupancy bitmap in HBM (B_occ of the roofline formula) */
int prv_map_bytes(prv_ctx* ctx, uint64_t* bitmap_bytes);
/* writes >L2-size scratch (256 MiB) on the cast stream so the next kernels start cold; its device time is reported separately
 * (prv_timing::flush_ms) */
int prv_flush_l2(prv_ctx* ctx);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU; NCCL over NVLink) */
int prv_comm_unique_id(void* id_out_128 /* ncclUniqueId bytes */);
int prv_comm_init(prv_ctx* ctx, const void* id_128, int rank, int nranks);
/* all-gathers the local coverage rows (equal V on every rank, view ids from prv_set_view_ids) so every rank
 * holds the full table; the greedy then runs replicated and deterministic on every rank.  Transport: the peer-memory
 * exchange when prv_comm_p2p_import has been called (stores over NVLink into every rank's table + release / acquire flags,
 * no NCCL; fused into the count kernel by PRV_CAST_PUBLISH), else ncclAllGather (prv_comm_init). */
int prv_allgather_bitsets_async(prv_ctx* ctx);
/* Peer-memory exchange set-up (one process per GPU, all on one NVLink / NVSwitch box): every rank allocates its arena and
 * exports a 64-byte cudaIpc handle; the caller exchanges the handles (torch.distributed, MPI, a file ...) and every rank
 * imports all of them (handles = nranks x 64 bytes in rank order; its own slot is ignored).  table_bytes_max bounds
 * nranks * V * (words * 8 + 4), the gathered table + ids of one step (0 = 16 MB; the 1024-view workload needs 1.8 MB). */
int prv_comm_p2p_export(prv_ctx* ctx, void* handle_out_64, uint64_t table_bytes_max);
int prv_comm_p2p_import(prv_ctx* ctx, const void* handles, int rank, int nranks);
int prv_comm_p2p_close(prv_ctx* ctx); /* unmap the peer arenas; prv_allgather_bitsets_async goes back to NCCL (prv_comm_destroy calls it too) */
/* the table the replicated selection runs over after the all-gather: nranks * V rows in rank order and their view ids
 * (rows_out [nrows][words], ids_out [nrows]; either may be NULL; *nrows_out = nranks * V).  Synchronises. */
int prv_get_gathered(prv_ctx* ctx, uint64_t* rows_out, uint32_t* ids_out, uint32_t* nrows_out);
int prv_comm_destroy(prv_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PRV_B200_H */
