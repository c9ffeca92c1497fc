// Small host-only harness for tests/test_host_cpp.py: exercises the C++ mirror classes (no GPU calls).
//   host_check yaml <config.yaml> [name]         -> dumps the parsed Share_Data fields
//   host_check views <config.yaml> <cloud.txt>   -> View_Space + poses for the configured view set
//   host_check png <out.png> <w> <h>             -> writes a deterministic RGBA test pattern
//   host_check json <out.json>                   -> writes a 2-frame transforms file
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../nerf-prv_b200/host/NBV_Net_Labeler.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    const std::string cmd = argv[1];
    std::cout.precision(17);
    if (cmd == "yaml") {
        Share_Data sd(argv[2], argc > 3 ? argv[3] : "", -1, argc > 4 ? atoi(argv[4]) : -1);
        std::cout << "RESULT pre_path=" << sd.pre_path << "\nRESULT model_path=" << sd.model_path << "\nRESULT viewspace_path=" << sd.viewspace_path
                  << "\nRESULT name_of_pcd=" << sd.name_of_pcd << "\nRESULT gt_path=" << sd.gt_path << "\nRESULT save_path=" << sd.save_path
                  << "\nRESULT num_of_thread=" << sd.num_of_thread << "\nRESULT ground_truth_resolution=" << sd.ground_truth_resolution
                  << "\nRESULT octomap_resolution=" << sd.octomap_resolution << "\nRESULT coverage_view_num_max=" << sd.coverage_view_num_max
                  << "\nRESULT coverage_view_num_add=" << sd.coverage_view_num_add << "\nRESULT points_size_cloud=" << sd.points_size_cloud
                  << "\nRESULT num_of_max_iteration=" << sd.num_of_max_iteration << "\nRESULT num_of_views=" << sd.num_of_views
                  << "\nRESULT view_space_radius=" << sd.view_space_radius << "\nRESULT width=" << sd.color_intrinsics.width
                  << "\nRESULT height=" << sd.color_intrinsics.height << "\nRESULT fx=" << sd.color_intrinsics.fx << "\nRESULT fy=" << sd.color_intrinsics.fy
                  << "\nRESULT ppx=" << sd.color_intrinsics.ppx << "\nRESULT ppy=" << sd.color_intrinsics.ppy << "\nRESULT model=" << sd.color_intrinsics.model
                  << "\nRESULT c0=" << sd.color_intrinsics.coeffs[0] << "\nRESULT c1=" << sd.color_intrinsics.coeffs[1] << "\nRESULT c2=" << sd.color_intrinsics.coeffs[2]
                  << "\nRESULT c3=" << sd.color_intrinsics.coeffs[3] << "\nRESULT c4=" << sd.color_intrinsics.coeffs[4] << "\nRESULT depth_scale=" << sd.depth_scale
                  << "\nRESULT object_pixel_rate=" << sd.object_pixel_rate << "\nRESULT is_shape_net=" << sd.is_shape_net << "\nRESULT show=" << sd.show
                  << "\nRESULT ensemble_num=" << sd.ensemble_num << "\nRESULT pt_sphere=" << sd.pt_sphere.size() << "\nRESULT pt_norm=" << sd.pt_norm
                  << "\nRESULT ray_casting_aabb_scale=" << sd.ray_casting_aabb_scale << "\nRESULT n_steps=" << sd.n_steps << std::endl;
        return 0;
    }
    if (cmd == "views") {
        std::shared_ptr<Share_Data> sd = std::make_shared<Share_Data>(argv[2], "obj", -1);
        std::ifstream fin(argv[3]);
        float x, y, z;
        while (fin >> x >> y >> z) sd->cloud_ground_truth->points.push_back(prv::make_point(x, y, z, 1, 2, 3));
        View_Space vs(sd);
        std::cout << "RESULT center " << vs.object_center_world(0) << ' ' << vs.object_center_world(1) << ' ' << vs.object_center_world(2) << "\n";
        std::cout << "RESULT size " << vs.predicted_size << "\n";
        for (size_t i = 0; i < vs.views.size(); i++) {
            vs.views[i].get_next_camera_pos(sd->now_camera_pose_world, sd->object_center_world);
            prv::Matrix4d pw = sd->now_camera_pose_world * vs.views[i].pose.inverse();
            std::cout << "RESULT view " << i;
            for (int a = 0; a < 3; a++) std::cout << ' ' << vs.views[i].init_pos(a);
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) std::cout << ' ' << pw(r, c);
            std::cout << "\n";
        }
        // toward poses
        for (int s = 0; s < 6; s++) {
            prv::Matrix4d t = sd->get_toward_pose(s);
            std::cout << "RESULT toward " << s;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) std::cout << ' ' << t(r, c);
            std::cout << "\n";
        }
        return 0;
    }
    if (cmd == "png") {
        const int w = atoi(argv[3]), h = atoi(argv[4]);
        std::vector<uint8_t> px((size_t)w * h * 4);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                uint8_t* p = &px[((size_t)y * w + x) * 4];
                p[0] = (uint8_t)(x * 7 + y);
                p[1] = (uint8_t)(y * 3);
                p[2] = (uint8_t)(x ^ y);
                p[3] = (uint8_t)((x + y) & 1 ? 255 : 0);
            }
        return prv::write_png(argv[2], px.data(), w, h, 4) ? 0 : 1;
    }
    if (cmd == "json") {
        std::map<std::string, double> reals = {{"camera_angle_x", 1.25}, {"fl_x", 915.60668945312500}, {"scale", 5.0}, {"k1", 0.12042199820280075}};
        std::map<std::string, long long> ints = {{"w", 1280}, {"h", 720}, {"aabb_scale", 1}};
        const double offset[3] = {0.5, 0.5000001, 0.25};
        std::vector<prv::JsonFrame> frames(2);
        for (int f = 0; f < 2; f++) {
            frames[f].file_path = "3/rgbaClip_" + std::to_string(f) + ".png";
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) frames[f].transform[r][c] = (r == c) ? 1.0 : 0.1 * (r + 1) + 0.01 * c + f;
        }
        return prv::write_transforms_json(argv[2], reals, ints, offset, frames) ? 0 : 1;
    }
    return 1;
}
