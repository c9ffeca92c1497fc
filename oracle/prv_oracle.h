/*
 * prv_oracle.h -- CPU ORACLE for the PRV_simulation ray-cast / coverage / greedy / splat path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (nerf-prv_b200/, include/) may include, link,
 * import or execute this.  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline
 * and `--impl reference` legs.
 *
 * PARITY: pinned bit for bit against the reference's OWN code, compiled from where it lies into oracle/_ref/
 * (`make -C oracle ref`, tests/test_oracle_kat.py): the camera functions rs2_project_point_to_pixel /
 * rs2_deproject_pixel_to_point (nothing but reference code), View::get_next_camera_pos (reference logic over our Eigen
 * shim) and Perception_3D::precept_thread_process + project_pixel_to_ray_end (reference logic over Eigen / OctoMap / PCL
 * shims, castRay answered by this oracle).  PARITY UNPINNED for what lives in the absent libraries: OctoMap 1.9.6
 * castRay / key maths / leaf order and Eigen 3.3.9 rounding.  The reference (psc0628/NeRF-PRV) ships no tests, golden
 * vectors or fixtures for this path, and as a whole cannot be compiled here (needs OctoMap, PCL 1.9.1, VTK, Eigen,
 * OpenCV, Gurobi, JsonCpp and Win32 headers; main.cpp:7 is ill-formed for g++).  For those parts the oracle restates the
 * published algorithm (OctoMap 1.9.6 castRay / key maths, octomath::Vector3, Eigen 3.3 fixed-size 4x4 inverse) and pins
 * itself with analytic known-answer tests, an independent Python restatement of castRay (tests/test_oracle_kat.py) and
 * frozen golden vectors (tests/golden/).
 *
 * All entry points are extern "C" so tests can drive them through ctypes.
 */
#ifndef PRV_ORACLE_H
#define PRV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as rs2_intrinsics, reference PRV_simulation/Share_Data.hpp:79-89. */
typedef struct orc_intrinsics {
    int   width;
    int   height;
    float ppx;
    float ppy;
    float fx;
    float fy;
    int   model;      /* rs2_distortion enum, Share_Data.hpp:67-76 */
    float coeffs[5];
} orc_intrinsics;

/* Same memory image as pcl::PointXYZRGB as far as the reference uses it (main.cpp:240-283). */
typedef struct orc_point_xyzrgb {
    float   x, y, z;
    uint8_t r, g, b, pad;
} orc_point_xyzrgb;

typedef struct orc_map orc_map;

typedef struct orc_cast_stats {
    uint64_t rays;        /* rays handed to castRay                                   */
    uint64_t steps;       /* DDA loop iterations executed (plain sequential march)      */
    uint64_t probes_in;   /* steps whose key lies inside the occupancy AABB (= S_in)    */
    uint64_t hits;        /* rays that returned a first-hit voxel                       */
} orc_cast_stats;

/* ---- linear algebra (Eigen 3.3 restatement) ---- */
void orc_mat4_inverse(const double m[16], double out[16]);            /* row-major */
void orc_mat4_mul(const double a[16], const double b[16], double out[16]);

/* ---- camera maths: Share_Data.hpp:92-137, :140-196, :719-726 ---- */
void orc_project_point_to_pixel(float pixel[2], const orc_intrinsics* in, const float point[3]);
void orc_deproject_pixel_to_point(float point[3], const orc_intrinsics* in, const float pixel[2], float depth);
void orc_project_pixel_to_ray_end(int x, int y, const orc_intrinsics* in, const double pose_world[16],
                                  float max_range, float out[3]);

/* ---- View::get_next_camera_pos case 0 (View_Space.hpp:69-140) and view_pose_world (main.cpp:109) ---- */
void orc_view_pose(const double now_camera_pose_world[16], const double init_pos[3],
                   const double object_center_world[3], double pose_out[16]);
void orc_view_pose_world(const double now_camera_pose_world[16], const double pose[16], double out[16]);

/* ---- View_Space::get_view_space (View_Space.hpp:517-558) ----
 * pts: P x 3 float (cloud_ground_truth).  sphere: N x 3 double (pt_sphere).  Returns number of views written. */
int orc_view_space(const float* pts, uint64_t P, const double* sphere, int N, double pt_norm,
                   double view_space_radius, double center_out[3], double* predicted_size_out,
                   double* init_pos_out /* N x 3 */);

/* ---- cloud normalisation (main.cpp:674, 749-832, 1005-1012): rotate by toward pose 4, centre, scale ---- */
void orc_normalize_cloud(float* pts /* P x 3, in place */, uint64_t P, double target_size /* random_size */,
                         double* predicted_size_before_out);

/* ---- OctoMap 1.9.6 key maths ---- */
int    orc_coord_to_key(double coord, double resolution, uint16_t* key_out);   /* coordToKeyChecked */
double orc_key_to_coord(uint16_t key, double resolution);                      /* keyToCoord (double) */

/* ---- map build: main.cpp:1005-1036, 1041, 1055-1058 (first point's colour wins; Morton leaf order) ---- */
orc_map* orc_map_build(const float* pts, const uint8_t* rgb, uint64_t P, double resolution);
orc_map* orc_map_from_keys(const uint16_t* keys /* N x 3 */, const uint8_t* rgb /* N x 3 or NULL */,
                           uint32_t N, double resolution);
void     orc_map_free(orc_map*);
uint32_t orc_map_size(const orc_map*);                       /* full_voxels */
void     orc_map_keys(const orc_map*, uint16_t* keys_out);   /* N x 3, leaf (Morton) order */
void     orc_map_rgb(const orc_map*, uint8_t* rgb_out);      /* N x 3 */
void     orc_map_aabb(const orc_map*, int lo[3], int hi[3]); /* inclusive key bounds */
void     orc_map_set_slow_lookup(orc_map*, int on);          /* 1: hash-set search() instead of bitmap */

/* ---- OccupancyOcTreeBase::castRay (OctoMap 1.9.6), call site main.cpp:258 ----
 * Returns 1 (hit) / 0; end_out = last `end` written by the algorithm; hit_rank_out = leaf rank or 0xFFFFFFFF. */
int orc_cast_ray(const orc_map*, const float origin[3], const float direction[3], int ignore_unknown,
                 double max_range, float end_out[3], uint32_t* hit_rank_out, orc_cast_stats* stats);

/* ---- Perception_3D::precept (main.cpp:98-236) + precept_thread_process (main.cpp:238-284) ----
 * out: full_voxels points (zeros = not visible).  hit_rank_out (optional): rank of the voxel seen by voxel i's ray.
 * Returns 1 if the view origin has a key ("View out of map" otherwise -> 0). */
int orc_precept(const orc_map*, const orc_intrinsics*, const double view_pose_world[16], const double init_pos[3],
                double max_range, orc_point_xyzrgb* out, uint32_t* hit_rank_out, orc_cast_stats* stats);

/* Same as orc_precept but with the reference's own execution structure (main.cpp:124-130): one std::thread per voxel,
 * created and joined in batches of num_of_thread, each recomputing the pose inverse, with the hash-set (sparse) lookup.
 * Only for the "reference-faithful structure" CPU baseline of BASELINE.md section 3; results equal orc_precept. */
int orc_precept_threads(orc_map*, const orc_intrinsics*, const double view_pose_world[16], const double init_pos[3],
                        double max_range, int num_of_thread, orc_point_xyzrgb* out, uint32_t* hit_rank_out);

/* ---- dense per-pixel mode (north-star "one ray per pixel"): same per-ray code as precept from
 * project_pixel_to_ray_end on, for every integer pixel (x,y) in [0,W)x[0,H).
 * hit_rank_out: H*W (row-major, 0xFFFFFFFF = none); depth_out (optional): sqrt of castRay's d^2, float. */
int orc_cast_view_dense(const orc_map*, const orc_intrinsics*, const double view_pose_world[16],
                        const double init_pos[3], double max_range, uint32_t* hit_rank_out, float* depth_out,
                        orc_cast_stats* stats, int num_threads);

/* ---- coverage bitsets (frozen definition) ---- */
uint32_t orc_bitset_words(uint32_t n_occ);                  /* u64 words per row, padded to 16 B */
void orc_bitset_from_ranks(const uint32_t* ranks, uint64_t n, uint64_t* row, uint32_t words);
uint32_t orc_popcount_row(const uint64_t* row, uint32_t words);

/* ---- greedy set cover (frozen definition) ----
 * Returns length of seq (>=1).  gains[0] = popcount(vis[first_view]). */
uint32_t orc_greedy(const uint64_t* vis /* V x words */, uint32_t V, uint32_t words, uint32_t first_view,
                    uint32_t max_iter, uint32_t* seq, uint32_t* gains, uint64_t* covered_out /* words, optional */,
                    uint64_t* views_scored_out);

/* ---- splat z-buffer (frozen definition) ----
 * rgba_out: H*W*4 (R,G,B,A), depth_out: H*W float (0 = background), index_out optional H*W (0xFFFFFFFF none). */
void orc_splat(const float* pts, const uint8_t* rgb, uint64_t P, const orc_intrinsics*,
               const double view_pose_world[16], int point_size, uint8_t* rgba_out, float* depth_out,
               uint32_t* index_out);
/* focal length in pixels of the PCL/VTK-equivalent pin-hole used by the splat definition */
float orc_splat_focal(const orc_intrinsics*);

/* ---- ensemble-uncertainty view scoring: NBV_Net_Labeler::nbv_loop cases 2 and 3 (main.cpp:2039-2161) ----
 * images: [V][E][H][W][4] uint8 in the channel order cv::imread(IMREAD_UNCHANGED) yields (channels 0..2 colour, 3 alpha).
 * method 2: sum over pixels/channels of log(variance) where variance > 1e-10; method 3: mean colour variance + (1-mean alpha)^2.
 * chosen (may be NULL): views to skip.  Returns argmax with the reference's strict '>' from -1e100 (-1 if none). */
int orc_score_ensemble(const uint8_t* images, uint32_t V, uint32_t E, int W, int H, int method, const uint8_t* chosen,
                       double* scores_out);

#ifdef __cplusplus
}
#endif
#endif
