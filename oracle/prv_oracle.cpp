/*
 * prv_oracle.cpp -- CPU ORACLE (test infrastructure only; see prv_oracle.h).  Camera maths pinned to the reference's compiled code; castRay / pose maths PARITY UNPINNED.
 *
 * Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile).  -ffp-contract=off
 * is REQUIRED: every float/double expression below must round exactly as written (no FMA).
 *
 * Reference paths are relative to /root/reference/PRV_simulation/.
 */
#include "prv_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ------------------------------------------------------------------------------------------------
 * Eigen 3.3 restatement (un-vendored third party; README.md:7 pins Eigen 3.3.9).
 *  - Matrix4d * Matrix4d / Matrix4d * Vector4d: coefficient-based product for fixed sizes < 8; with
 *    SSE2 packets of two doubles the accumulation runs over the inner index in ascending order,
 *    res = a0*b0; res = a_k*b_k + res (mul then add, no FMA on SSE2)  => ((p0+p1)+p2)+p3.
 *  - Matrix4d::inverse(): scalar cofactor path (compute_inverse_size4 via cofactor_4x4 /
 *    general_det3_helper), determinant = (col(0) . row(0) of the cofactor transpose) summed as
 *    (e0+e1)+(e2+e3), then element-wise division.
 *  - Vector3d::norm(): sqrt(x*x + (y*y + z*z))  (fixed-size unrolled redux splits 3 as 1 + 2).
 *  - Vector3d::normalized(): n = squaredNorm(); n > 0 ? v / sqrt(n) : v.
 * ------------------------------------------------------------------------------------------------ */
struct M4 {
    double a[4][4];
};

M4 m4_identity() {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.a[i][j] = (i == j) ? 1.0 : 0.0;
    return r;
}

M4 m4_load(const double* p) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.a[i][j] = p[i * 4 + j];
    return r;
}

void m4_store(const M4& m, double* p) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) p[i * 4 + j] = m.a[i][j];
}

M4 m4_mul(const M4& x, const M4& y) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = x.a[i][0] * y.a[0][j];
            s = x.a[i][1] * y.a[1][j] + s;
            s = x.a[i][2] * y.a[2][j] + s;
            s = x.a[i][3] * y.a[3][j] + s;
            r.a[i][j] = s;
        }
    return r;
}

void m4_vec(const M4& x, const double v[4], double out[4]) {
    for (int i = 0; i < 4; i++) {
        double s = x.a[i][0] * v[0];
        s = x.a[i][1] * v[1] + s;
        s = x.a[i][2] * v[2] + s;
        s = x.a[i][3] * v[3] + s;
        out[i] = s;
    }
}

double det3_helper(const M4& m, int i1, int i2, int i3, int j1, int j2, int j3) {
    return m.a[i1][j1] * (m.a[i2][j2] * m.a[i3][j3] - m.a[i2][j3] * m.a[i3][j2]);
}

double cofactor4(const M4& m, int i, int j) {
    const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4;
    const int j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
    return det3_helper(m, i1, i2, i3, j1, j2, j3) + det3_helper(m, i2, i3, i1, j1, j2, j3) +
           det3_helper(m, i3, i1, i2, j1, j2, j3);
}

M4 m4_inverse(const M4& m) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            const double c = cofactor4(m, i, j);
            r.a[j][i] = ((i + j) & 1) ? -c : c;
        }
    const double det = (m.a[0][0] * r.a[0][0] + m.a[1][0] * r.a[0][1]) + (m.a[2][0] * r.a[0][2] + m.a[3][0] * r.a[0][3]);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.a[i][j] = r.a[i][j] / det;
    return r;
}

struct V3 {
    double x, y, z;
};

double v3_sqnorm(const V3& v) { return v.x * v.x + (v.y * v.y + v.z * v.z); }
double v3_norm(const V3& v) { return std::sqrt(v3_sqnorm(v)); }
V3 v3_normalized(const V3& v) {
    const double n = v3_sqnorm(v);
    if (n > 0.0) {
        const double s = std::sqrt(n);
        return V3{v.x / s, v.y / s, v.z / s};
    }
    return v;
}
V3 v3_cross(const V3& a, const V3& b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

/* ------------------------------------------------------------------------------------------------
 * OctoMap 1.9.6 restatement (un-vendored third party; README.md:7).  tree depth 16.
 * ------------------------------------------------------------------------------------------------ */
constexpr int kTreeMaxVal = 32768;

bool coord_to_key(double coord, double resolution_factor, uint16_t& key) {
    const int scaled = ((int)std::floor(resolution_factor * coord)) + kTreeMaxVal;
    if (scaled >= 0 && ((unsigned)scaled) < (unsigned)(2 * kTreeMaxVal)) {
        key = (uint16_t)scaled;
        return true;
    }
    return false;
}

double key_to_coord(uint16_t key, double resolution) {
    return (double((int)key - kTreeMaxVal) + 0.5) * resolution;
}

/* leaf order of begin_leafs(): children visited in ascending index, x = bit 0, y = bit 1, z = bit 2
 * at every level => ascending 48-bit Morton code with z most significant inside each triple. */
uint64_t morton48(uint16_t kx, uint16_t ky, uint16_t kz) {
    uint64_t m = 0;
    for (int b = 15; b >= 0; b--) {
        m = (m << 3) | (uint64_t)((((kz >> b) & 1) << 2) | (((ky >> b) & 1) << 1) | ((kx >> b) & 1));
    }
    return m;
}

uint64_t pack_key(uint16_t kx, uint16_t ky, uint16_t kz) {
    return ((uint64_t)kz << 32) | ((uint64_t)ky << 16) | (uint64_t)kx;
}

}  // namespace

struct orc_map {
    double resolution = 0.0;
    double resolution_factor = 0.0;
    std::vector<uint16_t> keys;  // N x 3, Morton order
    std::vector<uint8_t> rgb;    // N x 3
    int lo[3] = {0, 0, 0}, hi[3] = {-1, -1, -1};
    int dim[3] = {0, 0, 0};
    std::vector<uint32_t> grid;  // dense AABB table: rank+1 or 0
    std::unordered_set<uint64_t> set;
    std::vector<std::pair<uint64_t, uint32_t>> sorted;  // packed key -> rank (slow path)
    bool slow = false;

    uint32_t n() const { return (uint32_t)(keys.size() / 3); }

    bool in_aabb(const uint16_t k[3]) const {
        return (int)k[0] >= lo[0] && (int)k[0] <= hi[0] && (int)k[1] >= lo[1] && (int)k[1] <= hi[1] &&
               (int)k[2] >= lo[2] && (int)k[2] <= hi[2];
    }

    /* search(key) != NULL && isNodeOccupied(node).  Only occupied leaves are ever inserted and the
     * reference never prunes (main.cpp:1019 lazy_eval=true, :1041 updateInnerOccupancy only), so this
     * is membership in the inserted-key set.  Returns rank or 0xFFFFFFFF. */
    uint32_t lookup(const uint16_t k[3]) const {
        if (slow) {
            const uint64_t p = pack_key(k[0], k[1], k[2]);
            if (!set.count(p)) return 0xFFFFFFFFu;
            auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(p, (uint32_t)0));
            return it->second;
        }
        if (!in_aabb(k)) return 0xFFFFFFFFu;
        const size_t idx = ((size_t)(k[2] - lo[2]) * dim[1] + (size_t)(k[1] - lo[1])) * dim[0] + (size_t)(k[0] - lo[0]);
        const uint32_t v = grid[idx];
        return v ? v - 1 : 0xFFFFFFFFu;
    }

    void finalize() {
        const uint32_t N = n();
        if (N == 0) return;
        for (int a = 0; a < 3; a++) {
            lo[a] = 65536;
            hi[a] = -1;
        }
        for (uint32_t i = 0; i < N; i++)
            for (int a = 0; a < 3; a++) {
                lo[a] = std::min(lo[a], (int)keys[3 * i + a]);
                hi[a] = std::max(hi[a], (int)keys[3 * i + a]);
            }
        for (int a = 0; a < 3; a++) dim[a] = hi[a] - lo[a] + 1;
        grid.assign((size_t)dim[0] * dim[1] * dim[2], 0u);
        sorted.resize(N);
        for (uint32_t i = 0; i < N; i++) {
            const uint16_t* k = &keys[3 * i];
            const size_t idx = ((size_t)(k[2] - lo[2]) * dim[1] + (size_t)(k[1] - lo[1])) * dim[0] + (size_t)(k[0] - lo[0]);
            grid[idx] = i + 1;
            const uint64_t p = pack_key(k[0], k[1], k[2]);
            set.insert(p);
            sorted[i] = std::make_pair(p, i);
        }
        std::sort(sorted.begin(), sorted.end());
    }
};

namespace {

/* rs2_project_point_to_pixel, Share_Data.hpp:92-137 (float arithmetic, evaluation order as written). */
void project_point_to_pixel(float pixel[2], const orc_intrinsics* in, const float point[3]) {
    float x = point[0] / point[2], y = point[1] / point[2];
    if (in->model == 1 || in->model == 2) {
        float r2 = x * x + y * y;
        float f = 1 + in->coeffs[0] * r2 + in->coeffs[1] * r2 * r2 + in->coeffs[4] * r2 * r2 * r2;
        x *= f;
        y *= f;
        float dx = x + 2 * in->coeffs[2] * x * y + in->coeffs[3] * (r2 + 2 * x * x);
        float dy = y + 2 * in->coeffs[3] * x * y + in->coeffs[2] * (r2 + 2 * y * y);
        x = dx;
        y = dy;
    }
    // models 3 / 5: the reference is C++ with `using namespace std` (Share_Data.hpp:63), so tan / atan on float arguments
    // are the FLOAT overloads of <cmath>; checked against the reference's compiled code (oracle/_ref, tests/test_oracle_kat.py)
    if (in->model == 3) {
        float r = sqrtf(x * x + y * y);
        if (r < FLT_EPSILON) r = FLT_EPSILON;
        float rd = (float)(1.0f / in->coeffs[0] * std::atan(2 * r * std::tan(in->coeffs[0] / 2.0f)));
        x *= rd / r;
        y *= rd / r;
    }
    if (in->model == 5) {
        float r = sqrtf(x * x + y * y);
        if (r < FLT_EPSILON) r = FLT_EPSILON;
        float theta = std::atan(r);
        float theta2 = theta * theta;
        float series = 1 + theta2 * (in->coeffs[0] + theta2 * (in->coeffs[1] + theta2 * (in->coeffs[2] + theta2 * in->coeffs[3])));
        float rd = theta * series;
        x *= rd / r;
        y *= rd / r;
    }
    pixel[0] = x * in->fx + in->ppx;
    pixel[1] = y * in->fy + in->ppy;
}

/* rs2_deproject_pixel_to_point, Share_Data.hpp:140-196.  (Share_Data.hpp:142 asserts model != 1.) */
void deproject_pixel_to_point(float point[3], const orc_intrinsics* in, const float pixel[2], float depth) {
    float x = (pixel[0] - in->ppx) / in->fx;
    float y = (pixel[1] - in->ppy) / in->fy;
    if (in->model == 2) {
        float r2 = x * x + y * y;
        float f = 1 + in->coeffs[0] * r2 + in->coeffs[1] * r2 * r2 + in->coeffs[4] * r2 * r2 * r2;
        float ux = x * f + 2 * in->coeffs[2] * x * y + in->coeffs[3] * (r2 + 2 * x * x);
        float uy = y * f + 2 * in->coeffs[3] * x * y + in->coeffs[2] * (r2 + 2 * y * y);
        x = ux;
        y = uy;
    }
    if (in->model == 5) {
        float rd = sqrtf(x * x + y * y);
        if (rd < FLT_EPSILON) rd = FLT_EPSILON;
        float theta = rd;
        float theta2 = rd * rd;
        for (int i = 0; i < 4; i++) {
            float f = theta * (1 + theta2 * (in->coeffs[0] + theta2 * (in->coeffs[1] + theta2 * (in->coeffs[2] + theta2 * in->coeffs[3])))) - rd;
            if (std::abs(f) < FLT_EPSILON) break;
            float df = 1 + theta2 * (3 * in->coeffs[0] + theta2 * (5 * in->coeffs[1] + theta2 * (7 * in->coeffs[2] + 9 * theta2 * in->coeffs[3])));
            theta -= f / df;
            theta2 = theta * theta;
        }
        float r = std::tan(theta);
        x *= r / rd;
        y *= r / rd;
    }
    if (in->model == 3) {
        float rd = sqrtf(x * x + y * y);
        if (rd < FLT_EPSILON) rd = FLT_EPSILON;
        float r = (float)(std::tan(in->coeffs[0] * rd) / std::atan(2 * std::tan(in->coeffs[0] / 2.0f)));
        x *= r / rd;
        y *= r / rd;
    }
    point[0] = depth * x;
    point[1] = depth * y;
    point[2] = depth;
}

/* project_pixel_to_ray_end, Share_Data.hpp:719-726.  NOTE the int parameters: the caller's float
 * pixel is truncated toward zero at the call (main.cpp:253). */
void pixel_to_ray_end(int x, int y, const orc_intrinsics* in, const M4& pose, float max_range, float out[3]) {
    float pixel[2] = {float(x), float(y)};
    float point[3];
    deproject_pixel_to_point(point, in, pixel, max_range);
    const double pw[4] = {point[0], point[1], point[2], 1};
    double r[4];
    m4_vec(pose, pw, r);
    out[0] = (float)r[0];
    out[1] = (float)r[1];
    out[2] = (float)r[2];
}

/* OccupancyOcTreeBase<NODE>::castRay, OctoMap 1.9.6 (octomap/OccupancyOcTreeBase.hxx), restated from the
 * published algorithm; point3d is 3 x float, octomath::Vector3::normalized() as in octomath/Vector3.h. */
int cast_ray(const orc_map* map, const float origin[3], const float directionP[3], bool ignoreUnknown,
             double maxRange, float end[3], uint32_t* hit_rank, orc_cast_stats* st) {
    if (hit_rank) *hit_rank = 0xFFFFFFFFu;
    if (st) st->rays++;
    const double res = map->resolution;
    uint16_t key[3];
    for (int i = 0; i < 3; i++)
        if (!coord_to_key((double)origin[i], map->resolution_factor, key[i])) return 0;

    {
        const uint32_t r0 = map->lookup(key);
        if (r0 != 0xFFFFFFFFu) {  // occupied node at origin
            for (int i = 0; i < 3; i++) end[i] = (float)key_to_coord(key[i], res);
            if (hit_rank) *hit_rank = r0;
            if (st) st->hits++;
            return 1;
        } else if (!ignoreUnknown) {
            for (int i = 0; i < 3; i++) end[i] = (float)key_to_coord(key[i], res);
            return 0;
        }
    }

    // direction = directionP.normalized():  norm_sq in float, len = sqrt((double)norm_sq), v /= (float)len
    float direction[3] = {directionP[0], directionP[1], directionP[2]};
    {
        const float nsq = direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2];
        const double len = std::sqrt((double)nsq);
        if (len > 0) {
            const float fl = (float)len;
            for (int i = 0; i < 3; i++) direction[i] /= fl;
        }
    }
    const bool max_range_set = (maxRange > 0.0);

    int step[3];
    double tMax[3], tDelta[3];
    for (int i = 0; i < 3; i++) {
        if (direction[i] > 0.0)
            step[i] = 1;
        else if (direction[i] < 0.0)
            step[i] = -1;
        else
            step[i] = 0;
        if (step[i] != 0) {
            double voxelBorder = key_to_coord(key[i], res);
            voxelBorder += double(step[i] * res * 0.5);
            tMax[i] = (voxelBorder - (double)origin[i]) / (double)direction[i];
            tDelta[i] = res / std::fabs((double)direction[i]);
        } else {
            tMax[i] = DBL_MAX;
            tDelta[i] = DBL_MAX;
        }
    }
    if (step[0] == 0 && step[1] == 0 && step[2] == 0) return 0;

    const double maxrange_sq = maxRange * maxRange;

    for (;;) {
        unsigned dim;
        if (tMax[0] < tMax[1]) {
            if (tMax[0] < tMax[2])
                dim = 0;
            else
                dim = 2;
        } else {
            if (tMax[1] < tMax[2])
                dim = 1;
            else
                dim = 2;
        }
        if ((step[dim] < 0 && key[dim] == 0) || (step[dim] > 0 && key[dim] == 2 * kTreeMaxVal - 1)) {
            for (int i = 0; i < 3; i++) end[i] = (float)key_to_coord(key[i], res);
            return 0;
        }
        key[dim] = (uint16_t)((int)key[dim] + step[dim]);
        tMax[dim] += tDelta[dim];
        if (st) st->steps++;

        for (int i = 0; i < 3; i++) end[i] = (float)key_to_coord(key[i], res);

        if (max_range_set) {
            double dist_from_origin_sq = 0.0;
            for (int j = 0; j < 3; j++) dist_from_origin_sq += ((end[j] - origin[j]) * (end[j] - origin[j]));
            if (dist_from_origin_sq > maxrange_sq) return 0;
        }
        if (st && map->in_aabb(key)) st->probes_in++;
        const uint32_t r = map->lookup(key);
        if (r != 0xFFFFFFFFu) {
            if (hit_rank) *hit_rank = r;
            if (st) st->hits++;
            return 1;
        } else if (!ignoreUnknown) {
            return 0;
        }
    }
}

/* per-view set-up shared by precept and the dense mode: main.cpp:111-114. */
bool view_origin(const orc_map* map, const double init_pos[3], float origin[3]) {
    uint16_t k[3];
    for (int i = 0; i < 3; i++)
        if (!coord_to_key(init_pos[i], map->resolution_factor, k[i])) return false;
    for (int i = 0; i < 3; i++) origin[i] = (float)key_to_coord(k[i], map->resolution);
    return true;
}

/* the tail of precept_thread_process from main.cpp:253 on, for an integer pixel.  Returns hit rank. */
uint32_t cast_pixel(const orc_map* map, const orc_intrinsics* in, const M4& pose_world, const float origin[3],
                    int px, int py, double max_range, float end_point[3], orc_cast_stats* st) {
    float end[3];
    pixel_to_ray_end(px, py, in, pose_world, 1.0f, end);  // main.cpp:253 (max_range argument is the literal 1.0)
    const float direction[3] = {end[0] - origin[0], end[1] - origin[1], end[2] - origin[2]};  // main.cpp:255
    uint32_t rank = 0xFFFFFFFFu;
    const int found = cast_ray(map, origin, direction, true, max_range, end_point, &rank, st);  // main.cpp:258
    if (!found) return 0xFFFFFFFFu;                                                                // :259-262
    if (end_point[0] == origin[0] && end_point[1] == origin[1] && end_point[2] == origin[2]) {     // :263-267
        if (st && st->hits) st->hits--;
        return 0xFFFFFFFFu;
    }
    // main.cpp:269-271: coordToKeyChecked(end_point) + search(key_end)
    uint16_t ke[3];
    for (int i = 0; i < 3; i++)
        if (!coord_to_key((double)end_point[i], map->resolution_factor, ke[i])) return 0xFFFFFFFFu;
    return map->lookup(ke);
}

}  // namespace

extern "C" {

void orc_mat4_inverse(const double m[16], double out[16]) { m4_store(m4_inverse(m4_load(m)), out); }
void orc_mat4_mul(const double a[16], const double b[16], double out[16]) { m4_store(m4_mul(m4_load(a), m4_load(b)), out); }

void orc_project_point_to_pixel(float pixel[2], const orc_intrinsics* in, const float point[3]) {
    project_point_to_pixel(pixel, in, point);
}
void orc_deproject_pixel_to_point(float point[3], const orc_intrinsics* in, const float pixel[2], float depth) {
    deproject_pixel_to_point(point, in, pixel, depth);
}
void orc_project_pixel_to_ray_end(int x, int y, const orc_intrinsics* in, const double pose_world[16], float max_range,
                                  float out[3]) {
    pixel_to_ray_end(x, y, in, m4_load(pose_world), max_range, out);
}

/* View::get_next_camera_pos, type_of_pose == 0: View_Space.hpp:69-140. */
void orc_view_pose(const double now_camera_pose_world[16], const double init_pos[3], const double object_center_world[3],
                   double pose_out[16]) {
    const M4 now = m4_load(now_camera_pose_world);
    const M4 now_inv = m4_inverse(now);
    const double oc[4] = {object_center_world[0], object_center_world[1], object_center_world[2], 1};
    const double vp[4] = {init_pos[0], init_pos[1], init_pos[2], 1};
    double object_c[4], view_c[4];
    m4_vec(now_inv, oc, object_c);  // :73
    m4_vec(now_inv, vp, view_c);    // :75
    const V3 object{object_c[0], object_c[1], object_c[2]};
    const V3 view{view_c[0], view_c[1], view_c[2]};
    V3 Z{object.x - view.x, object.y - view.y, object.z - view.z};
    Z = v3_normalized(Z);                      // :79
    V3 X = v3_normalized(v3_cross(Z, view));   // :81
    V3 Y = v3_normalized(v3_cross(Z, X));      // :82
    M4 T = m4_identity();                      // :83-87
    T.a[0][3] = -view.x;
    T.a[1][3] = -view.y;
    T.a[2][3] = -view.z;
    M4 R = m4_identity();                      // :88-92
    R.a[0][0] = X.x; R.a[0][1] = Y.x; R.a[0][2] = Z.x;
    R.a[1][0] = X.y; R.a[1][1] = Y.y; R.a[1][2] = Z.y;
    R.a[2][0] = X.z; R.a[2][1] = Y.z; R.a[2][2] = Z.z;

    double Rz_min[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};  // :94
    const double xv[4] = {1, 0, 0, 1}, yv[4] = {0, 1, 0, 1};
    double x_ray[4], y_ray[4];
    {
        const M4 A = m4_mul(m4_inverse(R), T);  // R.inverse() * T * ray  == (R.inverse()*T)*ray, :99-100
        m4_vec(A, xv, x_ray);
        m4_vec(A, yv, y_ray);
    }
    double min_y = std::acos(yv[1] * y_ray[1]);  // :101
    double min_x = std::acos(xv[0] * x_ray[0]);  // :102
    for (double i = 5; i < 360; i += 5) {        // :103
        /* AngleAxisd(0,X)*AngleAxisd(0,Y)*AngleAxisd(a,Z) is a Quaternion product; the two identity
         * quaternions leave (w,x,y,z) = (cos(a/2),0,0,sin(a/2)) exactly; Quaternion::toRotationMatrix(). */
        const double ang = i * std::acos(-1.0) / 180.0;
        const double qw = std::cos(0.5 * ang), qz = std::sin(0.5 * ang);
        const double tz = 2.0 * qz;
        const double twz = tz * qw, tzz = tz * qz;
        double rot[3][3] = {{1.0 - (0.0 + tzz), 0.0 - twz, 0.0}, {0.0 + twz, 1.0 - (0.0 + tzz), 0.0}, {0.0, 0.0, 1.0 - (0.0 + 0.0)}};
        M4 Rz = m4_identity();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) Rz.a[r][c] = rot[r][c];
        const M4 A = m4_mul(m4_inverse(m4_mul(R, Rz)), T);  // :114-115
        double xr[4], yr[4];
        m4_vec(A, xv, xr);
        m4_vec(A, yv, yr);
        const double cos_y = std::acos(yv[1] * yr[1]);
        const double cos_x = std::acos(xv[0] * xr[0]);
        if (cos_y < min_y) {  // :118
            std::memcpy(Rz_min, rot, sizeof(rot));
            min_y = cos_y;
            min_x = cos_x;
        } else if (std::fabs(cos_y - min_y) < 1e-6 && cos_x < min_x) {  // :123
            std::memcpy(Rz_min, rot, sizeof(rot));
            min_y = cos_y;
            min_x = cos_x;
        }
    }
    M4 Rz = m4_identity();  // :132-136
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Rz.a[r][c] = Rz_min[r][c];
    const M4 pose = m4_mul(m4_inverse(m4_mul(R, Rz)), T);  // :137
    m4_store(pose, pose_out);
}

/* view_pose_world = now_camera_pose_world * pose.inverse(), main.cpp:72 / :109. */
void orc_view_pose_world(const double now_camera_pose_world[16], const double pose[16], double out[16]) {
    m4_store(m4_mul(m4_load(now_camera_pose_world), m4_inverse(m4_load(pose))), out);
}

/* View_Space::get_view_space, View_Space.hpp:517-558 (points are float cloud coordinates widened to double, :567-569). */
int orc_view_space(const float* pts, uint64_t P, const double* sphere, int N, double pt_norm, double view_space_radius,
                   double center_out[3], double* predicted_size_out, double* init_pos_out) {
    double c[3] = {0, 0, 0};
    for (uint64_t i = 0; i < P; i++) {
        c[0] += (double)pts[3 * i + 0];
        c[1] += (double)pts[3 * i + 1];
        c[2] += (double)pts[3 * i + 2];
    }
    c[0] /= (double)P;
    c[1] /= (double)P;
    c[2] /= (double)P;
    double predicted_size = 0.0;
    for (uint64_t i = 0; i < P; i++) {
        const V3 d{c[0] - (double)pts[3 * i + 0], c[1] - (double)pts[3 * i + 1], c[2] - (double)pts[3 * i + 2]};
        predicted_size = std::max(predicted_size, v3_norm(d));
    }
    predicted_size *= 17.0 / 16.0;
    int nv = 0;
    for (int i = 0; i < N; i++) {
        if (sphere[3 * i + 2] < 0) continue;
        const double scale = 1.0 / pt_norm * view_space_radius;
        init_pos_out[3 * nv + 0] = sphere[3 * i + 0] * scale + c[0];
        init_pos_out[3 * nv + 1] = sphere[3 * i + 1] * scale + c[1];
        init_pos_out[3 * nv + 2] = sphere[3 * i + 2] * scale + c[2];
        nv++;
    }
    center_out[0] = c[0];
    center_out[1] = c[1];
    center_out[2] = c[2];
    *predicted_size_out = predicted_size;
    return nv;
}

/* main.cpp:674 (toward pose 4: (x,y,z)->(x,z,y), exact), :768-790 (centre), :800-832 (size), :1008-1010 (scale). */
void orc_normalize_cloud(float* pts, uint64_t P, double target_size, double* predicted_size_before_out) {
    for (uint64_t i = 0; i < P; i++) std::swap(pts[3 * i + 1], pts[3 * i + 2]);
    double c[3] = {0, 0, 0};
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) c[a] += (double)pts[3 * i + a];
    for (int a = 0; a < 3; a++) c[a] /= (double)P;
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) pts[3 * i + a] = (float)((double)pts[3 * i + a] - c[a]);
    double c2[3] = {0, 0, 0};
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) c2[a] += (double)pts[3 * i + a];
    for (int a = 0; a < 3; a++) c2[a] /= (double)P;
    double predicted_size = 0.0;
    for (uint64_t i = 0; i < P; i++) {
        const V3 d{c2[0] - (double)pts[3 * i + 0], c2[1] - (double)pts[3 * i + 1], c2[2] - (double)pts[3 * i + 2]};
        predicted_size = std::max(predicted_size, v3_norm(d));
    }
    predicted_size *= 17.0 / 16.0;
    if (predicted_size_before_out) *predicted_size_before_out = predicted_size;
    const double scale = target_size / predicted_size;  // main.cpp:962
    const float unit = 1.0f;                            // main.cpp:756
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) pts[3 * i + a] = (float)((double)pts[3 * i + a] * scale * unit);  // :1008-1010
}

int orc_coord_to_key(double coord, double resolution, uint16_t* key_out) {
    uint16_t k = 0;
    const bool ok = coord_to_key(coord, 1.0 / resolution, k);
    if (ok && key_out) *key_out = k;
    return ok ? 1 : 0;
}
double orc_key_to_coord(uint16_t key, double resolution) { return key_to_coord(key, resolution); }

orc_map* orc_map_from_keys(const uint16_t* keys, const uint8_t* rgb, uint32_t N, double resolution) {
    orc_map* m = new orc_map();
    m->resolution = resolution;
    m->resolution_factor = 1.0 / resolution;
    std::vector<std::pair<uint64_t, uint32_t>> order(N);
    for (uint32_t i = 0; i < N; i++) order[i] = std::make_pair(morton48(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]), i);
    std::sort(order.begin(), order.end());
    m->keys.reserve((size_t)N * 3);
    m->rgb.reserve((size_t)N * 3);
    uint64_t prev = ~0ull;
    for (uint32_t j = 0; j < N; j++) {
        if (order[j].first == prev) continue;  // duplicates: first occurrence (stable by index) wins
        prev = order[j].first;
        const uint32_t i = order[j].second;
        for (int a = 0; a < 3; a++) m->keys.push_back(keys[3 * i + a]);
        for (int a = 0; a < 3; a++) m->rgb.push_back(rgb ? rgb[3 * i + a] : (uint8_t)0);
    }
    m->finalize();
    return m;
}

/* main.cpp:1005-1036: per point coordToKeyChecked(point3d(x,y,z)); insert only if search(key)==NULL;
 * the first point's colour is the voxel colour (integrateNodeColor on a fresh node). */
orc_map* orc_map_build(const float* pts, const uint8_t* rgb, uint64_t P, double resolution) {
    const double rf = 1.0 / resolution;
    std::vector<uint16_t> keys;
    std::vector<uint8_t> col;
    std::unordered_set<uint64_t> seen;
    keys.reserve(P);
    for (uint64_t i = 0; i < P; i++) {
        uint16_t k[3];
        bool ok = true;
        for (int a = 0; a < 3; a++) ok = coord_to_key((double)pts[3 * i + a], rf, k[a]) && ok;
        if (!ok) continue;
        if (!seen.insert(pack_key(k[0], k[1], k[2])).second) continue;
        for (int a = 0; a < 3; a++) keys.push_back(k[a]);
        for (int a = 0; a < 3; a++) col.push_back(rgb ? rgb[3 * i + a] : (uint8_t)0);
    }
    return orc_map_from_keys(keys.data(), col.data(), (uint32_t)(keys.size() / 3), resolution);
}

void orc_map_free(orc_map* m) { delete m; }
uint32_t orc_map_size(const orc_map* m) { return m->n(); }
void orc_map_keys(const orc_map* m, uint16_t* out) { std::memcpy(out, m->keys.data(), m->keys.size() * sizeof(uint16_t)); }
void orc_map_rgb(const orc_map* m, uint8_t* out) { std::memcpy(out, m->rgb.data(), m->rgb.size()); }
void orc_map_aabb(const orc_map* m, int lo[3], int hi[3]) {
    for (int a = 0; a < 3; a++) {
        lo[a] = m->lo[a];
        hi[a] = m->hi[a];
    }
}
void orc_map_set_slow_lookup(orc_map* m, int on) { m->slow = on != 0; }

int orc_cast_ray(const orc_map* m, const float origin[3], const float direction[3], int ignore_unknown, double max_range,
                 float end_out[3], uint32_t* hit_rank_out, orc_cast_stats* stats) {
    float end[3] = {0, 0, 0};
    const int r = cast_ray(m, origin, direction, ignore_unknown != 0, max_range, end, hit_rank_out, stats);
    if (end_out) std::memcpy(end_out, end, sizeof(end));
    return r;
}

int orc_precept(const orc_map* map, const orc_intrinsics* in, const double view_pose_world[16], const double init_pos[3],
                double max_range, orc_point_xyzrgb* out, uint32_t* hit_rank_out, orc_cast_stats* stats) {
    const uint32_t N = map->n();
    std::memset(out, 0, (size_t)N * sizeof(orc_point_xyzrgb));
    if (hit_rank_out)
        for (uint32_t i = 0; i < N; i++) hit_rank_out[i] = 0xFFFFFFFFu;
    float origin[3];
    if (!view_origin(map, init_pos, origin)) return 0;  // "View out of map.check." main.cpp:139
    const M4 pose = m4_load(view_pose_world);
    for (uint32_t i = 0; i < N; i++) {  // main.cpp:124-130 fan-out; each i is independent
        // end[i] = leaf_iterator.getCoordinate(): main.cpp:116-121
        const float e[3] = {(float)key_to_coord(map->keys[3 * i + 0], map->resolution),
                            (float)key_to_coord(map->keys[3 * i + 1], map->resolution),
                            (float)key_to_coord(map->keys[3 * i + 2], map->resolution)};
        const double end_3d[4] = {e[0], e[1], e[2], 1};
        const M4 inv = m4_inverse(pose);  // main.cpp:244 (recomputed per voxel in the reference)
        double vertex[4];
        m4_vec(inv, end_3d, vertex);
        const float point_3d[3] = {(float)vertex[0], (float)vertex[1], (float)vertex[2]};  // :245
        float pixel[2];
        project_point_to_pixel(pixel, in, point_3d);  // :247
        if (pixel[0] < 0 || pixel[0] > in->width || pixel[1] < 0 || pixel[1] > in->height) continue;  // :248
        if (pixel[0] != pixel[0] || pixel[1] != pixel[1]) continue;  // NaN -> int is UB in the reference; frozen: reject
        float end_point[3];
        const uint32_t rank = cast_pixel(map, in, pose, origin, (int)pixel[0], (int)pixel[1], max_range, end_point, stats);
        if (rank == 0xFFFFFFFFu) continue;
        out[i].x = end_point[0];  // :274-279
        out[i].y = end_point[1];
        out[i].z = end_point[2];
        out[i].r = map->rgb[3 * rank + 0];
        out[i].g = map->rgb[3 * rank + 1];
        out[i].b = map->rgb[3 * rank + 2];
        if (hit_rank_out) hit_rank_out[i] = rank;
    }
    return 1;
}

namespace {
void precept_one_voxel(const orc_map* map, const orc_intrinsics* in, const M4* pose, const float* origin, double max_range, uint32_t i,
                       orc_point_xyzrgb* out, uint32_t* hit_rank_out) {
    const float e[3] = {(float)key_to_coord(map->keys[3 * i + 0], map->resolution), (float)key_to_coord(map->keys[3 * i + 1], map->resolution),
                        (float)key_to_coord(map->keys[3 * i + 2], map->resolution)};
    const double end_3d[4] = {e[0], e[1], e[2], 1};
    const M4 inv = m4_inverse(*pose);  // recomputed per voxel, main.cpp:244
    double vertex[4];
    m4_vec(inv, end_3d, vertex);
    const float point_3d[3] = {(float)vertex[0], (float)vertex[1], (float)vertex[2]};
    float pixel[2];
    project_point_to_pixel(pixel, in, point_3d);
    if (pixel[0] < 0 || pixel[0] > in->width || pixel[1] < 0 || pixel[1] > in->height) return;
    if (pixel[0] != pixel[0] || pixel[1] != pixel[1]) return;
    float end_point[3];
    const uint32_t rank = cast_pixel(map, in, *pose, origin, (int)pixel[0], (int)pixel[1], max_range, end_point, nullptr);
    if (rank == 0xFFFFFFFFu) return;
    out[i].x = end_point[0];
    out[i].y = end_point[1];
    out[i].z = end_point[2];
    out[i].r = map->rgb[3 * rank + 0];
    out[i].g = map->rgb[3 * rank + 1];
    out[i].b = map->rgb[3 * rank + 2];
    if (hit_rank_out) hit_rank_out[i] = rank;
}
}  // namespace

int orc_precept_threads(orc_map* map, const orc_intrinsics* in, const double view_pose_world[16], const double init_pos[3], double max_range,
                        int num_of_thread, orc_point_xyzrgb* out, uint32_t* hit_rank_out) {
    const uint32_t N = map->n();
    std::memset(out, 0, (size_t)N * sizeof(orc_point_xyzrgb));
    if (hit_rank_out)
        for (uint32_t i = 0; i < N; i++) hit_rank_out[i] = 0xFFFFFFFFu;
    float origin[3];
    if (!view_origin(map, init_pos, origin)) return 0;
    const M4 pose = m4_load(view_pose_world);
    const bool was_slow = map->slow;
    map->slow = true;  // sparse (tree-like) lookup
    std::vector<std::thread> precept_process;
    precept_process.reserve(N);
    for (uint32_t i = 0; i < N; i += (uint32_t)num_of_thread) {  // main.cpp:125-130
        for (uint32_t j = 0; j < (uint32_t)num_of_thread && i + j < N; j++)
            precept_process.emplace_back(precept_one_voxel, map, in, &pose, origin, max_range, i + j, out, hit_rank_out);
        for (uint32_t j = 0; j < (uint32_t)num_of_thread && i + j < N; j++) precept_process[i + j].join();
    }
    map->slow = was_slow;
    return 1;
}

int orc_cast_view_dense(const orc_map* map, const orc_intrinsics* in, const double view_pose_world[16],
                        const double init_pos[3], double max_range, uint32_t* hit_rank_out, float* depth_out,
                        orc_cast_stats* stats, int num_threads) {
    const int W = in->width, H = in->height;
    for (size_t i = 0; i < (size_t)W * H; i++) hit_rank_out[i] = 0xFFFFFFFFu;
    if (depth_out) std::memset(depth_out, 0, (size_t)W * H * sizeof(float));
    float origin[3];
    if (!view_origin(map, init_pos, origin)) return 0;
    const M4 pose = m4_load(view_pose_world);
    orc_cast_stats total = {0, 0, 0, 0};
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads)
#endif
    for (int y = 0; y < H; y++) {
        orc_cast_stats st = {0, 0, 0, 0};
        for (int x = 0; x < W; x++) {
            float end_point[3];
            const uint32_t rank = cast_pixel(map, in, pose, origin, x, y, max_range, end_point, stats ? &st : nullptr);
            hit_rank_out[(size_t)y * W + x] = rank;
            if (rank != 0xFFFFFFFFu && depth_out) {
                double d2 = 0.0;
                for (int j = 0; j < 3; j++) d2 += ((end_point[j] - origin[j]) * (end_point[j] - origin[j]));
                depth_out[(size_t)y * W + x] = (float)std::sqrt(d2);
            }
        }
        if (stats) {
#ifdef _OPENMP
#pragma omp critical
#endif
            {
                total.rays += st.rays;
                total.steps += st.steps;
                total.probes_in += st.probes_in;
                total.hits += st.hits;
            }
        }
    }
    if (stats) {
        stats->rays += total.rays;
        stats->steps += total.steps;
        stats->probes_in += total.probes_in;
        stats->hits += total.hits;
    }
    return 1;
}

uint32_t orc_bitset_words(uint32_t n_occ) {
    uint32_t w = (n_occ + 63) / 64;
    if (w == 0) w = 1;
    return (w + 1) & ~1u;  // rows padded to 16 B
}

void orc_bitset_from_ranks(const uint32_t* ranks, uint64_t n, uint64_t* row, uint32_t words) {
    std::memset(row, 0, (size_t)words * 8);
    for (uint64_t i = 0; i < n; i++)
        if (ranks[i] != 0xFFFFFFFFu) row[ranks[i] >> 6] |= 1ull << (ranks[i] & 63);
}

uint32_t orc_popcount_row(const uint64_t* row, uint32_t words) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < words; i++) c += (uint32_t)__builtin_popcountll(row[i]);
    return c;
}

/* Frozen greedy definition (tie rule from the reference's strict '>' argmax, main.cpp:2006,2088,2152). */
uint32_t orc_greedy(const uint64_t* vis, uint32_t V, uint32_t words, uint32_t first_view, uint32_t max_iter, uint32_t* seq,
                    uint32_t* gains, uint64_t* covered_out, uint64_t* views_scored_out) {
    std::vector<uint64_t> covered(words, 0);
    std::vector<uint8_t> chosen(V, 0);
    uint64_t scored = 0;
    uint32_t n = 0;
    seq[n] = first_view;
    gains[n] = orc_popcount_row(vis + (size_t)first_view * words, words);
    n++;
    chosen[first_view] = 1;
    for (uint32_t w = 0; w < words; w++) covered[w] = vis[(size_t)first_view * words + w];
    for (uint32_t it = 0; it < max_iter; it++) {
        uint32_t best = 0xFFFFFFFFu, best_gain = 0;
        for (uint32_t v = 0; v < V; v++) {
            if (chosen[v]) continue;
            scored++;
            uint32_t g = 0;
            const uint64_t* row = vis + (size_t)v * words;
            for (uint32_t w = 0; w < words; w++) g += (uint32_t)__builtin_popcountll(row[w] & ~covered[w]);
            if (best == 0xFFFFFFFFu || g > best_gain) {  // strict '>' : first maximum (lowest id) wins
                best = v;
                best_gain = g;
            }
        }
        if (best == 0xFFFFFFFFu || best_gain == 0) break;
        seq[n] = best;
        gains[n] = best_gain;
        n++;
        chosen[best] = 1;
        for (uint32_t w = 0; w < words; w++) covered[w] |= vis[(size_t)best * words + w];
    }
    if (covered_out) std::memcpy(covered_out, covered.data(), (size_t)words * 8);
    if (views_scored_out) *views_scored_out = scored;
    return n;
}

/* Frozen splat definition: the PCL 1.9.1 setCameraParameters(intrinsics, extrinsics)
 * camera (call main.cpp:79) forced to a W x H window (:80-84): vertical FOV from fy and 2*(int)ppy,
 * principal point at the window centre, square pixels, no distortion; then the 180-degree flip of
 * main.cpp:1616 turns the VTK image into the ordinary CV orientation (u right along +x_cam, v down along +y_cam). */
float orc_splat_focal(const orc_intrinsics* in) {
    return (float)((double)in->height * (double)in->fy / (2.0 * (double)(int)in->ppy));
}

void orc_splat(const float* pts, const uint8_t* rgb, uint64_t P, const orc_intrinsics* in, const double view_pose_world[16],
               int point_size, uint8_t* rgba_out, float* depth_out, uint32_t* index_out) {
    const int W = in->width, H = in->height;
    std::vector<uint64_t> zbuf((size_t)W * H, ~0ull);
    const M4 inv = m4_inverse(m4_load(view_pose_world));
    const float f = orc_splat_focal(in);
    const float cx = (float)W * 0.5f, cy = (float)H * 0.5f;
    const float off = 0.5f - 0.5f * (float)point_size;
    for (uint64_t i = 0; i < P; i++) {
        const double pw[4] = {pts[3 * i + 0], pts[3 * i + 1], pts[3 * i + 2], 1};
        double pc[4];
        m4_vec(inv, pw, pc);
        const float xc = (float)pc[0], yc = (float)pc[1], zc = (float)pc[2];
        if (!(zc > 0.01f && zc < 1000.01f)) continue;  // VTK clipping range 0.01 .. 1000.01
        const float u = (xc / zc) * f + cx;
        const float v = (yc / zc) * f + cy;
        if (!(u > -64.0f && u < (float)W + 64.0f && v > -64.0f && v < (float)H + 64.0f)) continue;
        const int lx = (int)std::floor(u + off), ly = (int)std::floor(v + off);
        uint32_t zb;
        std::memcpy(&zb, &zc, 4);
        const uint64_t packed = ((uint64_t)zb << 32) | (uint64_t)(uint32_t)i;
        for (int dy = 0; dy < point_size; dy++) {
            const int yy = ly + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = 0; dx < point_size; dx++) {
                const int xx = lx + dx;
                if (xx < 0 || xx >= W) continue;
                uint64_t& z = zbuf[(size_t)yy * W + xx];
                if (packed < z) z = packed;
            }
        }
    }
    for (size_t p = 0; p < (size_t)W * H; p++) {
        const uint64_t z = zbuf[p];
        uint8_t* o = rgba_out + 4 * p;
        if (z == ~0ull) {
            o[0] = o[1] = o[2] = 255;
            o[3] = 0;
            if (depth_out) depth_out[p] = 0.0f;
            if (index_out) index_out[p] = 0xFFFFFFFFu;
            continue;
        }
        const uint32_t idx = (uint32_t)(z & 0xFFFFFFFFu);
        const uint32_t zb = (uint32_t)(z >> 32);
        o[0] = rgb[3 * (size_t)idx + 0];
        o[1] = rgb[3 * (size_t)idx + 1];
        o[2] = rgb[3 * (size_t)idx + 2];
        o[3] = (o[0] == 255 && o[1] == 255 && o[2] == 255) ? 0 : 255;  // convertToAlpha, Share_Data.hpp:771-784
        if (depth_out) std::memcpy(&depth_out[p], &zb, 4);
        if (index_out) index_out[p] = idx;
    }
}

/* nbv_loop case 2 (main.cpp:2039-2097) and case 3 (:2099-2161), per view; argmax :2088 / :2152 with largest = -1e100 (:1971). */
int orc_score_ensemble(const uint8_t* images, uint32_t V, uint32_t E, int W, int H, int method, const uint8_t* chosen, double* scores_out) {
    double largest_view_uncertainty = -1e100;
    int best_view_id = -1;
    const size_t img = (size_t)W * H * 4;
    for (uint32_t i = 0; i < V; i++) {
        if (scores_out) scores_out[i] = 0.0;
        if (chosen && chosen[i]) continue;
        const uint8_t* base = images + (size_t)i * E * img;
        double view_uncertainty = 0.0;
        for (int j = 0; j < H; j++) {
            for (int k = 0; k < W; k++) {
                double mean[3] = {0.0, 0.0, 0.0};
                double mean_density = 0.0;
                for (uint32_t e = 0; e < E; e++) {
                    const uint8_t* px = base + e * img + ((size_t)j * W + k) * 4;
                    mean[0] += px[0];
                    mean[1] += px[1];
                    mean[2] += px[2];
                    mean_density += px[3] / 255.0;
                }
                for (int c = 0; c < 3; c++) mean[c] /= E;
                mean_density /= E;
                double variance[3] = {0.0, 0.0, 0.0};
                for (uint32_t e = 0; e < E; e++) {
                    const uint8_t* px = base + e * img + ((size_t)j * W + k) * 4;
                    for (int c = 0; c < 3; c++) variance[c] += (px[c] - mean[c]) * (px[c] - mean[c]);
                }
                for (int c = 0; c < 3; c++) variance[c] /= E;
                if (method == 2) {
                    for (int c = 0; c < 3; c++)
                        if (variance[c] > 1e-10) view_uncertainty += std::log(variance[c]);
                } else {
                    view_uncertainty += (variance[0] + variance[1] + variance[2]) / 3.0;
                    view_uncertainty += (1.0 - mean_density) * (1.0 - mean_density);
                }
            }
        }
        if (scores_out) scores_out[i] = view_uncertainty;
        if (view_uncertainty > largest_view_uncertainty) {
            largest_view_uncertainty = view_uncertainty;
            best_view_id = (int)i;
        }
    }
    return best_view_id;
}

}  // extern "C"
