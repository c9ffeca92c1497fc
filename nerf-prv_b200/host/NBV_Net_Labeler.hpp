// NBV_Net_Labeler.hpp -- host-side mirror of the part of the reference's `NBV_Net_Labeler`
// (PRV_simulation/main.cpp:597-2279) that feeds and drives the hot path: the constructor's cloud normalisation, size
// selection and ground-truth map insertion (main.cpp:630-1115) and `get_coverage()` (main.cpp:1581-1656).  The iterative
// NBV loop, TSP planner and the Instant-NGP / PRVNet bridges are out of scope.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Perception_3D.hpp"
#include "Share_Data.hpp"
#include "View_Space.hpp"
#include "prv_io.hpp"

inline double get_random_coordinate(double from, double to) {  // View_Space.hpp:32-38 of the reference
    const double len = to - from;
    const long long x = (long long)rand() * ((long long)RAND_MAX + 1) + (long long)rand();
    const long long field = (long long)RAND_MAX * (long long)RAND_MAX + 2 * (long long)RAND_MAX;
    return (double)x / (double)field * len + from;
}

class NBV_Net_Labeler {
public:
    std::shared_ptr<Share_Data> share_data;
    std::shared_ptr<View_Space> view_space;
    std::shared_ptr<Perception_3D> percept;
    int toward_state, rotate_state;
    bool object_is_ok_size = true;
    int device;

    NBV_Net_Labeler(std::shared_ptr<Share_Data>& _share_data, int _toward_state = 0, int _rotate_state = 0, int _device = 0) {
        share_data = _share_data;
        toward_state = _toward_state;
        rotate_state = _rotate_state;
        device = _device;
        std::vector<float> xyz;
        std::vector<uint8_t> rgb;
        const std::string ply = share_data->model_path + "ShapeNet/" + share_data->name_of_pcd + ".ply";  // main.cpp:647
        if (!share_data->is_shape_net || !prv::read_ply_xyzrgb(ply, xyz, rgb) || xyz.empty()) {
            std::cout << "Can not read 3d model file. Check. (" << ply << ")" << std::endl;
            object_is_ok_size = false;
            return;
        }
        const uint64_t P = xyz.size() / 3;
        std::cout << "points size is " << P << std::endl;

        // size: reuse size.txt, else draw 0.075..0.115 until the object fills enough pixels (main.cpp:851-964)
        double random_size = -1;
        share_data->access_directory(share_data->gt_path);
        std::ifstream size_reader(share_data->gt_path + "/size.txt");
        if (size_reader.is_open()) {
            size_reader >> random_size;
            if (random_size < 0) {
                std::cout << "no size. Skip." << std::endl;
                object_is_ok_size = false;
                return;
            }
        } else {
            random_size = 0.075;
            double object_rate = -1;
            int test_times = 0;
            do {
                random_size = get_random_coordinate(random_size, 0.115);
                std::cout << "random size is " << random_size << std::endl;
                object_rate = object_pixel_rate_at(xyz, rgb, random_size);
                std::cout << "now object rate is " << object_rate << std::endl;
                test_times++;
            } while (object_rate <= share_data->object_pixel_rate && test_times <= 5);
            std::ofstream fout(share_data->gt_path + "/size.txt");
            if (test_times <= 5) {
                fout << random_size;
            } else {
                object_is_ok_size = false;
                fout << -1;
                return;
            }
        }
        std::cout << "shapenet object " << share_data->name_of_pcd << " random size is " << random_size << " m." << std::endl;

        // rotate (toward pose 4), centre, scale: main.cpp:674, 768-832, 1008-1010
        std::vector<float> cloud_xyz = xyz;
        double before = 0;
        prv_host_normalize_cloud(cloud_xyz.data(), P, random_size, &before);
        share_data->octomap_resolution = random_size * 2.0 / 32.0;  // main.cpp:967-969
        share_data->cloud_ground_truth->points.resize(P);
        for (uint64_t i = 0; i < P; i++)
            share_data->cloud_ground_truth->points[i] = prv::make_point(cloud_xyz[3 * i], cloud_xyz[3 * i + 1], cloud_xyz[3 * i + 2],
                                                                       rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        share_data->cloud_ground_truth->width = (uint32_t)P;
        double min_z = 0;
        for (uint64_t i = 0; i < P; i++) min_z = std::min(min_z, (double)cloud_xyz[3 * i + 2]);
        share_data->min_z_table = min_z - share_data->ground_truth_resolution;  // main.cpp:1038

        // ground-truth map: main.cpp:1005-1036, leaf count :1055-1058
        auto& gt = share_data->ground_truth_model;
        gt->resolution = share_data->ground_truth_resolution;
        gt->keys.resize(P * 3);
        gt->rgb.resize(P * 3);
        uint32_t n = 0;
        prv_host_build_map(cloud_xyz.data(), rgb.data(), P, gt->resolution, gt->keys.data(), gt->rgb.data(), &n);
        gt->keys.resize((size_t)n * 3);
        gt->rgb.resize((size_t)n * 3);
        share_data->full_voxels = (int)n;

        view_space = std::make_shared<View_Space>(share_data);                 // main.cpp:1061
        percept = std::make_shared<Perception_3D>(share_data, device);         // main.cpp:1111
        object_is_ok_size = percept->ok;
    }

    // reload Hemisphere/<N>.txt and rebuild the view space (main.cpp:2353-2369)
    void set_view_count(int n) {
        share_data->load_view_space(n);
        view_space.reset();
        view_space = std::make_shared<View_Space>(share_data);
    }

    // images + instant-ngp transforms for the current view set (main.cpp:1581-1656); additionally (B200 path) the
    // per-view coverage counts and the greedy cover sequence go to <gt_path>/<N>_coverage.txt
    int get_coverage(bool write_coverage = true) {
        const rs2_intrinsics& in = share_data->color_intrinsics;
        const int N = share_data->num_of_views;
        const std::string dir = share_data->gt_path + "/" + std::to_string(N);
        share_data->access_directory(dir);
        std::vector<uint8_t> rgba;
        if (!percept->render_views(view_space->views, rgba)) return -1;
        const size_t px = (size_t)in.width * in.height * 4;
        std::vector<prv::JsonFrame> frames(view_space->views.size());
        for (size_t i = 0; i < view_space->views.size(); i++) {
            if (!prv::write_png(dir + "/rgbaClip_" + std::to_string(i) + ".png", rgba.data() + i * px, in.width, in.height, 4)) return -1;
            frames[i].file_path = std::to_string(N) + "/rgbaClip_" + std::to_string(i) + ".png";
            // transform_matrix = P * view_pose_world * P1, P: (x,y,z)->(y,z,x) rows, P1 = diag(1,-1,-1,1) (main.cpp:1629-1640)
            const prv::Matrix4d vpw = percept->pose_of(view_space->views[i]);
            prv::Matrix4d Pm, P1 = prv::Matrix4d::Identity();
            Pm(0, 2) = 1; Pm(1, 0) = 1; Pm(2, 1) = 1; Pm(3, 3) = 1;
            P1(1, 1) = -1; P1(2, 2) = -1;
            const prv::Matrix4d t = (Pm * vpw) * P1;
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) frames[i].transform[r][c] = t(r, c);
        }
        std::map<std::string, double> reals;
        std::map<std::string, long long> ints;
        reals["camera_angle_x"] = 2.0 * std::atan(0.5 * in.width / in.fx);
        reals["camera_angle_y"] = 2.0 * std::atan(0.5 * in.height / in.fy);
        reals["fl_x"] = in.fx; reals["fl_y"] = in.fy;
        reals["k1"] = in.coeffs[0]; reals["k2"] = in.coeffs[1]; reals["k3"] = in.coeffs[2]; reals["p1"] = in.coeffs[3]; reals["p2"] = in.coeffs[4];
        reals["cx"] = in.ppx; reals["cy"] = in.ppy;
        ints["w"] = in.width; ints["h"] = in.height;
        ints["aabb_scale"] = share_data->ray_casting_aabb_scale;
        reals["scale"] = 0.5 / share_data->predicted_size;
        const double offset[3] = {0.5 + share_data->object_center_world(2), 0.5 + share_data->object_center_world(0),
                                  0.5 + share_data->object_center_world(1)};
        if (write_coverage) {
            std::vector<uint64_t> bits;
            std::vector<uint32_t> counts, seq, gains;
            if (percept->precept_views(view_space->views, PRV_MODE_DENSE, bits, counts) && percept->greedy(0, seq, gains)) {
                std::ofstream fc(share_data->gt_path + "/" + std::to_string(N) + "_coverage.txt");
                fc << "full_voxels " << share_data->full_voxels << "\n" << "coverage_count";
                for (uint32_t c : counts) fc << ' ' << c;
                fc << "\ngreedy_seq";
                for (uint32_t s : seq) fc << ' ' << s;
                fc << "\ngreedy_gain";
                for (uint32_t g : gains) fc << ' ' << g;
                fc << "\n";
            }
        }
        // the JSON is written last: its existence is the reference's "this view set is done" marker (main.cpp:2351-2352)
        return prv::write_transforms_json(share_data->gt_path + "/" + std::to_string(N) + ".json", reals, ints, offset, frames) ? 0 : -1;
    }

private:
    // 5 test renders from Hemisphere/5.txt, fraction of non-white pixels (main.cpp:873-938)
    double object_pixel_rate_at(const std::vector<float>& xyz, const std::vector<uint8_t>& rgb, double size) {
        const uint64_t P = xyz.size() / 3;
        std::vector<float> pts = xyz;
        prv_host_normalize_cloud(pts.data(), P, size, nullptr);
        prv_ctx* ctx = nullptr;
        if (prv_create(&ctx, device) != PRV_OK) return -1;
        double rate = -1;
        std::ifstream fin(share_data->viewspace_path + "5.txt");
        std::vector<double> pw(5 * 16);
        double c[3] = {0, 0, 0};
        for (uint64_t i = 0; i < P; i++)
            for (int a = 0; a < 3; a++) c[a] += pts[3 * i + a];
        for (int a = 0; a < 3; a++) c[a] /= (double)P;
        const prv::Vector3d center(c[0], c[1], c[2]);
        bool good = fin.is_open();
        for (int i = 0; i < 5 && good; i++) {
            prv::Vector3d p;
            fin >> p(0) >> p(1) >> p(2);
            p = p / p.norm() * share_data->view_space_radius + center;
            View view(p);
            view.get_next_camera_pos(share_data->now_camera_pose_world, center);
            (share_data->now_camera_pose_world * view.pose.inverse()).toRowMajor(&pw[16 * i]);
        }
        // 5 renders and the non-white count stay on the device (main.cpp:913-931); only the counters come back
        double r = -1;
        if (good && prv_set_camera(ctx, &share_data->color_intrinsics, 1.0) == PRV_OK && prv_set_cloud(ctx, pts.data(), rgb.data(), P) == PRV_OK &&
            prv_object_pixel_rate(ctx, pw.data(), 5, share_data->points_size_cloud, &r, nullptr) == PRV_OK)
            rate = r;
        prv_destroy(ctx);
        return rate;
    }
};
