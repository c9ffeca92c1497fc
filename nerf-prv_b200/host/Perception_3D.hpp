// Perception_3D.hpp -- host-side mirror of the reference's `Perception_3D` (PRV_simulation/main.cpp:17-286): same
// constructor, `precept(View&)` and `render(View&, id, path)` with the same outputs, but OctoMap castRay and the
// PCL/VTK viewer are replaced by libprv_b200.so (include/prv.h).  Batched entry points (`precept_views`,
// `render_views`) are additions: they hand every candidate view to the GPU in one call.
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/prv.h"
#include "Share_Data.hpp"
#include "View_Space.hpp"
#include "prv_io.hpp"

class Perception_3D {
public:
    std::shared_ptr<Share_Data> share_data;
    std::shared_ptr<prv::GroundTruthModel> ground_truth_model;
    int full_voxels;
    prv::PointCloud::Ptr cloud;
    prv::Matrix4d view_pose_world;
    prv_ctx* ctx = nullptr;  // stands in for the octree and the PCLVisualizer
    bool ok = false;

    Perception_3D(std::shared_ptr<Share_Data>& _share_data, int device = 0) {
        share_data = _share_data;
        ground_truth_model = share_data->ground_truth_model;
        full_voxels = share_data->full_voxels;
        view_pose_world = prv::Matrix4d::Identity();
        cloud.reset(new prv::PointCloud);
        if (prv_create(&ctx, device) != PRV_OK) {
            std::cout << "Perception_3D: " << prv_last_error(nullptr) << std::endl;
            return;
        }
        ok = check(prv_set_map(ctx, ground_truth_model->keys.data(), ground_truth_model->rgb.data(), ground_truth_model->size(),
                               ground_truth_model->resolution)) &&
             check(prv_set_camera(ctx, &share_data->color_intrinsics, 1.0));  // castRay maxRange literal, main.cpp:258
        if (ok && share_data->is_shape_net && !share_data->cloud_ground_truth->points.empty()) {
            // viewer->addPointCloud(cloud_ground_truth) + point size (main.cpp:38-39)
            const auto& pts = share_data->cloud_ground_truth->points;
            std::vector<float> xyz(pts.size() * 3);
            std::vector<uint8_t> rgb(pts.size() * 3);
            for (size_t i = 0; i < pts.size(); i++) {
                xyz[3 * i] = pts[i].x; xyz[3 * i + 1] = pts[i].y; xyz[3 * i + 2] = pts[i].z;
                rgb[3 * i] = pts[i].r; rgb[3 * i + 1] = pts[i].g; rgb[3 * i + 2] = pts[i].b;
            }
            ok = check(prv_set_cloud(ctx, xyz.data(), rgb.data(), pts.size()));
        }
    }

    ~Perception_3D() {
        if (ctx) prv_destroy(ctx);
    }

    // view_pose_world of a view (main.cpp:71-72 / 108-109)
    prv::Matrix4d pose_of(View& v) {
        v.get_next_camera_pos(share_data->now_camera_pose_world, share_data->object_center_world);
        return share_data->now_camera_pose_world * v.pose.inverse();
    }

    // Off-screen render of the coloured cloud -> gt_path + path + "/rgb_<id>.png" (main.cpp:68-96).  The file has the
    // reference's orientation (the VTK image, i.e. before the 180-degree flip) and an opaque white background.
    bool render(View& now_best_view, int id, std::string path = "") {
        if (!ok) return false;
        view_pose_world = pose_of(now_best_view);
        const int W = share_data->color_intrinsics.width, H = share_data->color_intrinsics.height;
        double pw[16];
        view_pose_world.toRowMajor(pw);
        std::vector<uint8_t> rgba((size_t)W * H * 4);
        if (!check(prv_render_views(ctx, pw, 1, share_data->points_size_cloud, rgba.data(), nullptr))) return false;
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const uint8_t* s = &rgba[((size_t)(H - 1 - y) * W + (W - 1 - x)) * 4];
                uint8_t* d = &rgb[((size_t)y * W + x) * 3];
                d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
            }
        share_data->access_directory(share_data->gt_path + path);
        return prv::write_png(share_data->gt_path + path + "/rgb_" + std::to_string(id) + ".png", rgb.data(), W, H, 3);
    }

    // Virtual scan of one view: cloud->points[i] = first voxel seen along voxel i's ray, or zeros (main.cpp:98-236).
    bool precept(View& now_best_view) {
        if (!ok) return false;
        cloud.reset(new prv::PointCloud);
        cloud->is_dense = false;
        cloud->points.resize(full_voxels);
        view_pose_world = pose_of(now_best_view);
        double pw[16];
        view_pose_world.toRowMajor(pw);
        const double ip[3] = {now_best_view.init_pos(0), now_best_view.init_pos(1), now_best_view.init_pos(2)};
        int in_map = 0;
        if (!check(prv_precept(ctx, pw, ip, cloud->points.data(), &in_map))) return false;
        if (!in_map) std::cout << "View out of map.check." << std::endl;  // main.cpp:139
        share_data->vaild_clouds++;
        return true;
    }

    // All candidate views at once: coverage bitsets (bit i = leaf i visible) and counts.  mode: PRV_MODE_VOXEL is the
    // literal precept ray set, PRV_MODE_DENSE one ray per pixel.  The bitsets stay resident for greedy().
    bool precept_views(std::vector<View>& views, int mode, std::vector<uint64_t>& bitsets, std::vector<uint32_t>& counts) {
        if (!ok || views.empty()) return false;
        std::vector<double> pw(views.size() * 16), ip(views.size() * 3);
        for (size_t v = 0; v < views.size(); v++) {
            pose_of(views[v]).toRowMajor(&pw[16 * v]);
            for (int a = 0; a < 3; a++) ip[3 * v + a] = views[v].init_pos(a);
        }
        bitsets.assign(views.size() * (size_t)prv_bitset_words(ctx), 0);
        counts.assign(views.size(), 0);
        return check(prv_cast_views(ctx, pw.data(), ip.data(), (uint32_t)views.size(), mode, bitsets.data(), counts.data(), nullptr, nullptr));
    }

    // Greedy set-cover over the resident bitsets: start view, at most num_of_max_iteration further picks.
    bool greedy(uint32_t first_view, std::vector<uint32_t>& seq, std::vector<uint32_t>& gains) {
        const uint32_t max_iter = (uint32_t)std::max(share_data->num_of_max_iteration, 0);
        seq.assign(max_iter + 1, 0);
        gains.assign(max_iter + 1, 0);
        uint32_t n = 0;
        if (!ok || !check(prv_greedy(ctx, first_view, max_iter, seq.data(), gains.data(), (uint32_t)seq.size(), &n))) return false;
        seq.resize(n);
        gains.resize(n);
        return true;
    }

    // RGBA pixel content of rgbaClip_<i>.png for every view (white -> alpha 0, rotated 180 degrees: main.cpp:1611-1617)
    bool render_views(std::vector<View>& views, std::vector<uint8_t>& rgba) {
        if (!ok || views.empty()) return false;
        std::vector<double> pw(views.size() * 16);
        for (size_t v = 0; v < views.size(); v++) pose_of(views[v]).toRowMajor(&pw[16 * v]);
        const size_t px = (size_t)share_data->color_intrinsics.width * share_data->color_intrinsics.height;
        rgba.assign(views.size() * px * 4, 0);
        return check(prv_render_views(ctx, pw.data(), (uint32_t)views.size(), share_data->points_size_cloud, rgba.data(), nullptr));
    }

private:
    bool check(int rc) {
        if (rc == PRV_OK) return true;
        std::cout << "prv error " << rc << ": " << prv_last_error(ctx) << std::endl;
        return false;
    }
};
