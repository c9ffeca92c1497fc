// prv_host.cpp -- the pure-host entry points of include/prv.h (prv_host_*): pose, view-space,
// cloud normalisation and ground-truth map insertion arithmetic of the reference, implemented once
// on top of the host mirror classes so C++ and ctypes callers share one implementation.
// Compile with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/prv.h"
#include "../host/View_Space.hpp"
#include "prv_keys.hpp"

using prv::Matrix4d;
using prv::Vector3d;

namespace {

// index of the first key whose Morton code is not above its predecessor's, or N.  The loop runs on the host at the start of
// every prv_set_map while the GPU waits, so the codes are built with pdep where the CPU has BMI2 (45 k keys: 175 -> 45 us).
uint32_t first_order_violation_portable(const uint16_t* keys, uint32_t N) {
    uint64_t prev = 0;
    for (uint32_t i = 0; i < N; i++) {
        const uint64_t c = prv::morton_code(keys[3 * (size_t)i], keys[3 * (size_t)i + 1], keys[3 * (size_t)i + 2]);
        if (i && c <= prev) return i;
        prev = c;
    }
    return N;
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2"))) uint32_t first_order_violation_bmi2(const uint16_t* keys, uint32_t N) {
    const unsigned long long mx = 0x249249249249ull;  // bit b of x lands at bit 3b, y at 3b+1, z at 3b+2 (prv::morton_code)
    uint64_t prev = 0;
    for (uint32_t i = 0; i < N; i++) {
        const uint64_t c = __builtin_ia32_pdep_di(keys[3 * (size_t)i], mx) | __builtin_ia32_pdep_di(keys[3 * (size_t)i + 1], mx << 1) |
                           __builtin_ia32_pdep_di(keys[3 * (size_t)i + 2], mx << 2);
        if (i && c <= prev) return i;
        prev = c;
    }
    return N;
}
#endif
uint32_t first_order_violation(const uint16_t* keys, uint32_t N) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool have_bmi2 = __builtin_cpu_supports("bmi2");
    if (have_bmi2) return first_order_violation_bmi2(keys, N);
#endif
    return first_order_violation_portable(keys, N);
}

}  // namespace

extern "C" {

int prv_abi_version(void) { return PRV_ABI_VERSION; }

int prv_host_mat4_inverse(const double m[16], double out[16]) {
    if (!m || !out) return PRV_ERR_INVALID;
    Matrix4d::FromRowMajor(m).inverse().toRowMajor(out);
    return PRV_OK;
}

int prv_host_view_pose(const double now_camera_pose_world[16], const double init_pos[3], const double object_center_world[3],
                       double pose_out[16]) {
    if (!now_camera_pose_world || !init_pos || !object_center_world || !pose_out) return PRV_ERR_INVALID;
    View v(Vector3d(init_pos[0], init_pos[1], init_pos[2]));
    v.get_next_camera_pos(Matrix4d::FromRowMajor(now_camera_pose_world),
                          Vector3d(object_center_world[0], object_center_world[1], object_center_world[2]), 0);
    v.pose.toRowMajor(pose_out);
    return PRV_OK;
}

int prv_host_view_pose_world(const double now_camera_pose_world[16], const double pose[16], double out[16]) {
    if (!now_camera_pose_world || !pose || !out) return PRV_ERR_INVALID;
    (Matrix4d::FromRowMajor(now_camera_pose_world) * Matrix4d::FromRowMajor(pose).inverse()).toRowMajor(out);
    return PRV_OK;
}

int prv_host_view_space(const float* pts, uint64_t P, const double* pt_sphere, int N, double pt_norm, double view_space_radius,
                        double center_out[3], double* predicted_size_out, double* init_pos_out, int* n_views_out) {
    if (!pts || !pt_sphere || P == 0 || N < 0 || !center_out || !predicted_size_out || !init_pos_out || !n_views_out)
        return PRV_ERR_INVALID;
    // View_Space.hpp:533-547 of the reference; float cloud coordinates widened to double (:567-569)
    double c[3] = {0.0, 0.0, 0.0};
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) c[a] += (double)pts[3 * i + a];
    for (int a = 0; a < 3; a++) c[a] /= (double)P;
    const Vector3d center(c[0], c[1], c[2]);
    double size = 0.0;
    for (uint64_t i = 0; i < P; i++)
        size = std::max(size, (center - Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])).norm());
    size *= 17.0 / 16.0;
    int nv = 0;
    for (int i = 0; i < N; i++) {
        const double* s = pt_sphere + 3 * (size_t)i;
        if (s[2] < 0) continue;  // :549
        const double scale = 1.0 / pt_norm * view_space_radius;
        for (int a = 0; a < 3; a++) init_pos_out[3 * nv + a] = s[a] * scale + c[a];
        nv++;
    }
    for (int a = 0; a < 3; a++) center_out[a] = c[a];
    *predicted_size_out = size;
    *n_views_out = nv;
    return PRV_OK;
}

int prv_host_normalize_cloud(float* pts, uint64_t P, double target_size, double* predicted_size_before_out) {
    if (!pts || P == 0 || !(target_size > 0)) return PRV_ERR_INVALID;
    // main.cpp:674: get_toward_pose(4) maps (x,y,z) -> (x,z,y); the 0/1 products are exact, so it is a swap
    for (uint64_t i = 0; i < P; i++) std::swap(pts[3 * i + 1], pts[3 * i + 2]);
    auto centroid = [&](double c[3]) {
        c[0] = c[1] = c[2] = 0.0;
        for (uint64_t i = 0; i < P; i++)
            for (int a = 0; a < 3; a++) c[a] += (double)pts[3 * i + a];
        for (int a = 0; a < 3; a++) c[a] /= (double)P;
    };
    double c[3];
    centroid(c);  // main.cpp:768-783
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) pts[3 * i + a] = (float)((double)pts[3 * i + a] - c[a]);  // :786-790
    centroid(c);  // :811-821
    const Vector3d center(c[0], c[1], c[2]);
    double size = 0.0;  // :828-832
    for (uint64_t i = 0; i < P; i++)
        size = std::max(size, (center - Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])).norm());
    size *= 17.0 / 16.0;
    if (predicted_size_before_out) *predicted_size_before_out = size;
    const double scale = target_size / size;  // :962 (random_size / predicted_size)
    const float unit = 1.0f;                  // :756
    for (uint64_t i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) pts[3 * i + a] = (float)((double)pts[3 * i + a] * scale * unit);  // :1008-1010
    return PRV_OK;
}

int prv_host_build_map(const float* pts, const uint8_t* rgb, uint64_t P, double resolution, uint16_t* keys_out,
                       uint8_t* rgb_out, uint32_t* n_out) {
    if (!pts || !keys_out || !n_out || !(resolution > 0)) return PRV_ERR_INVALID;
    try {
    const double rf = 1.0 / resolution;
    // (morton, point index): sorting groups points of one voxel with the first-inserted point first,
    // which is the voxel whose colour the reference keeps (main.cpp:1015-1021).
    std::vector<std::pair<uint64_t, uint64_t>> order;
    order.reserve(P);
    for (uint64_t i = 0; i < P; i++) {
        uint16_t k[3];
        if (!prv::coord_to_key_checked(pts[3 * i], rf, k[0]) || !prv::coord_to_key_checked(pts[3 * i + 1], rf, k[1]) ||
            !prv::coord_to_key_checked(pts[3 * i + 2], rf, k[2]))
            continue;
        order.emplace_back(prv::morton_code(k[0], k[1], k[2]), i);
    }
    std::sort(order.begin(), order.end());
    uint32_t n = 0;
    uint64_t prev = ~0ull;
    for (const auto& e : order) {
        if (e.first == prev) continue;
        prev = e.first;
        uint16_t k[3];
        prv::morton_decode(e.first, k);
        for (int a = 0; a < 3; a++) keys_out[3 * (size_t)n + a] = k[a];
        if (rgb_out)
            for (int a = 0; a < 3; a++) rgb_out[3 * (size_t)n + a] = rgb ? rgb[3 * e.second + a] : (uint8_t)0;
        n++;
    }
    *n_out = n;
    } catch (...) {  // std::bad_alloc of the sort buffer: nothing throws across the C ABI
        return PRV_ERR_OOM;
    }
    return PRV_OK;
}

int prv_host_check_leaf_order(const uint16_t* keys, uint32_t N, uint32_t* first_bad_out) {
    if ((!keys && N) || !first_bad_out) return PRV_ERR_INVALID;
    *first_bad_out = first_order_violation(keys, N);
    return PRV_OK;
}

// rs2_project_point_to_pixel / rs2_deproject_pixel_to_point (Share_Data.hpp:92-137, :140-196) for every distortion model, float
// arithmetic in the reference's evaluation order.  The reference is C++ with `using namespace std`, so tan / atan / abs on
// float arguments are the float overloads: on models 3 (F-Theta) and 5 (Kannala-Brandt) the result is whatever THIS host's
// libm returns, exactly like the reference compiled on this host.  The device never evaluates a transcendental: for those two
// models prv_set_camera tabulates the deprojection of every integer pixel with this function and prv_cast_* (voxel mode)
// projects the voxel centres with the other one on the host.
int prv_host_project_point_to_pixel(const prv_intrinsics* in, const float point[3], float pixel_out[2]) {
    if (!in || !point || !pixel_out) return PRV_ERR_INVALID;
    float x = point[0] / point[2], y = point[1] / point[2];
    if (in->model == 1 || in->model == 2) {
        const float r2 = x * x + y * y;
        const float f = 1 + in->coeffs[0] * r2 + in->coeffs[1] * r2 * r2 + in->coeffs[4] * r2 * r2 * r2;
        x *= f;
        y *= f;
        const float dx = x + 2 * in->coeffs[2] * x * y + in->coeffs[3] * (r2 + 2 * x * x);
        const float dy = y + 2 * in->coeffs[3] * x * y + in->coeffs[2] * (r2 + 2 * y * y);
        x = dx;
        y = dy;
    } else if (in->model == 3 || in->model == 5) {
        float r = std::sqrt(x * x + y * y);
        if (r < 1.1920928955078125e-7f) r = 1.1920928955078125e-7f;  // FLT_EPSILON
        float rd;
        if (in->model == 3) {
            rd = (float)(1.0f / in->coeffs[0] * std::atan(2 * r * std::tan(in->coeffs[0] / 2.0f)));
        } else {
            const float theta = std::atan(r);
            const float theta2 = theta * theta;
            const float series = 1 + theta2 * (in->coeffs[0] + theta2 * (in->coeffs[1] + theta2 * (in->coeffs[2] + theta2 * in->coeffs[3])));
            rd = theta * series;
        }
        x *= rd / r;
        y *= rd / r;
    }
    pixel_out[0] = x * in->fx + in->ppx;
    pixel_out[1] = y * in->fy + in->ppy;
    return PRV_OK;
}

int prv_host_deproject_pixel_to_point(const prv_intrinsics* in, const float pixel[2], float depth, float point_out[3]) {
    if (!in || !pixel || !point_out) return PRV_ERR_INVALID;
    if (in->model == 1) return PRV_ERR_UNSUPPORTED;  // the reference asserts (Share_Data.hpp:142)
    float x = (pixel[0] - in->ppx) / in->fx;
    float y = (pixel[1] - in->ppy) / in->fy;
    if (in->model == 2) {
        const float r2 = x * x + y * y;
        const float f = 1 + in->coeffs[0] * r2 + in->coeffs[1] * r2 * r2 + in->coeffs[4] * r2 * r2 * r2;
        const float ux = x * f + 2 * in->coeffs[2] * x * y + in->coeffs[3] * (r2 + 2 * x * x);
        const float uy = y * f + 2 * in->coeffs[3] * x * y + in->coeffs[2] * (r2 + 2 * y * y);
        x = ux;
        y = uy;
    } else if (in->model == 3 || in->model == 5) {
        float rd = std::sqrt(x * x + y * y);
        if (rd < 1.1920928955078125e-7f) rd = 1.1920928955078125e-7f;
        float r;
        if (in->model == 5) {
            float theta = rd, theta2 = rd * rd;
            for (int i = 0; i < 4; i++) {  // Newton on theta * series(theta^2) = rd
                const float f = theta * (1 + theta2 * (in->coeffs[0] + theta2 * (in->coeffs[1] + theta2 * (in->coeffs[2] + theta2 * in->coeffs[3])))) - rd;
                if (std::abs(f) < 1.1920928955078125e-7f) break;
                const float df = 1 + theta2 * (3 * in->coeffs[0] + theta2 * (5 * in->coeffs[1] + theta2 * (7 * in->coeffs[2] + 9 * theta2 * in->coeffs[3])));
                theta -= f / df;
                theta2 = theta * theta;
            }
            r = std::tan(theta);
        } else {
            r = (float)(std::tan(in->coeffs[0] * rd) / std::atan(2 * std::tan(in->coeffs[0] / 2.0f)));
        }
        x *= r / rd;
        y *= r / rd;
    }
    point_out[0] = depth * x;
    point_out[1] = depth * y;
    point_out[2] = depth;
    return PRV_OK;
}

float prv_splat_focal(const prv_intrinsics* intr) {
    if (!intr) return 0.0f;
    // PCL 1.9.1 PCLVisualizer::setCameraParameters(intrinsics, extrinsics) (call: reference main.cpp:79):
    // fovy = 2*atan(window_h / (2 fy)) with window_h = 2*(int)cy; the window is then forced to W x H (:80-84).
    return (float)((double)intr->height * (double)intr->fy / (2.0 * (double)(int)intr->ppy));
}

}  // extern "C"
