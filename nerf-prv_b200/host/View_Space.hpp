// View_Space.hpp -- host-side mirror of the reference's `View` and `View_Space`
// (PRV_simulation/View_Space.hpp:40-199 and :492-728), same public names and argument meaning.
// All arithmetic is double and follows the reference's evaluation order; the visualisation branch
// (`show`) and the path-planning helpers (:206-490) are outside the hot path and not mirrored.
#pragma once
#include <cmath>
#include <iostream>
#include <memory>
#include <vector>

#include "Share_Data.hpp"
#include "prv_linalg.hpp"

class View {
public:
    prv::Vector3d init_pos;  // view position in the world
    prv::Matrix4d pose;      // world -> this view's camera frame (view_i to view_i+1 in the reference's words)

    explicit View(prv::Vector3d _init_pos) : init_pos(_init_pos), pose(prv::Matrix4d::Identity()) {}

    // type_of_pose 0: look-at with the 5-degree roll search (View_Space.hpp:69-140).
    // type_of_pose 1: look-at with "camera y highest" roll (View_Space.hpp:142-193).
    void get_next_camera_pos(prv::Matrix4d now_camera_pose_world, prv::Vector3d object_center_world, int type_of_pose = 0) {
        const prv::Matrix4d to_camera = now_camera_pose_world.inverse();
        const prv::Vector4d oc = to_camera * prv::Vector4d(object_center_world(0), object_center_world(1), object_center_world(2), 1);
        const prv::Vector4d vc = to_camera * prv::Vector4d(init_pos(0), init_pos(1), init_pos(2), 1);
        const prv::Vector3d object(oc(0), oc(1), oc(2));
        const prv::Vector3d view(vc(0), vc(1), vc(2));
        // camera frame: Z looks at the object, X = Z x view, Y = Z x X
        const prv::Vector3d Z = (object - view).normalized();
        const prv::Vector3d X = Z.cross(view).normalized();
        const prv::Vector3d Y = Z.cross(X).normalized();
        prv::Matrix4d T = prv::Matrix4d::Identity();
        prv::Matrix4d R = prv::Matrix4d::Identity();
        for (int r = 0; r < 3; r++) {
            T(r, 3) = -view(r);
            R(r, 0) = X(r);
            R(r, 1) = Y(r);
            R(r, 2) = Z(r);
        }
        double best[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        if (type_of_pose == 0) {
            // The reference pushes the POINTS (1,0,0,1) and (0,1,0,1) through (R*Rz)^-1 * T (translation included)
            // and minimises acos(y_ray.y), ties (within 1e-6) on acos(x_ray.x); NaN compares false.
            double min_y, min_x;
            roll_score(R, T, nullptr, min_y, min_x);
            for (double deg = 5; deg < 360; deg += 5) {
                double rot[3][3];
                prv::rotation_about_z(deg * std::acos(-1.0) / 180.0, rot);
                double cy, cx;
                roll_score(R, T, rot, cy, cx);
                const bool better = (cy < min_y) || (std::fabs(cy - min_y) < 1e-6 && cx < min_x);
                if (better) {
                    copy3(rot, best);
                    min_y = cy;
                    min_x = cx;
                }
            }
        } else {
            prv::Vector4d y_highest = ((now_camera_pose_world * R) * T) * prv::Vector4d(0, 1, 0, 1);
            for (double deg = 5; deg < 360; deg += 5) {
                double rot[3][3];
                prv::rotation_about_z(deg * std::acos(-1.0) / 180.0, rot);
                const prv::Vector4d y_now = (((now_camera_pose_world * R) * embed(rot)) * T) * prv::Vector4d(0, 1, 0, 1);
                if (y_now(2) > y_highest(2)) {
                    copy3(rot, best);
                    y_highest = y_now;
                }
            }
        }
        pose = (R * embed(best)).inverse() * T;
    }

private:
    static prv::Matrix4d embed(const double rot[3][3]) {
        prv::Matrix4d m = prv::Matrix4d::Identity();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) m(r, c) = rot[r][c];
        return m;
    }
    static void copy3(const double src[3][3], double dst[3][3]) {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) dst[r][c] = src[r][c];
    }
    static void roll_score(const prv::Matrix4d& R, const prv::Matrix4d& T, const double rot[3][3], double& score_y, double& score_x) {
        const prv::Matrix4d A = (rot ? (R * embed(rot)).inverse() : R.inverse()) * T;
        const prv::Vector4d x_ray = A * prv::Vector4d(1, 0, 0, 1);
        const prv::Vector4d y_ray = A * prv::Vector4d(0, 1, 0, 1);
        score_y = std::acos(1.0 * y_ray(1));
        score_x = std::acos(1.0 * x_ray(0));
    }
};

class View_Space {
public:
    int num_of_views = 0;
    std::vector<View> views;
    prv::Vector3d object_center_world;
    double predicted_size = 0;
    prv::Matrix4d now_camera_pose_world;
    int occupied_voxels = 0;
    double map_entropy = 0;
    double octomap_resolution = 0;
    std::shared_ptr<Share_Data> share_data;

    // fraction of points inside the axis-aligned cube of half-size `predicted_size` (View_Space.hpp:506-515)
    double check_size(double predicted_size, std::vector<prv::Vector3d>& points) {
        int vaild_points = 0;
        for (auto& p : points) {
            bool inside = true;
            for (int a = 0; a < 3; a++)
                if (p(a) < object_center_world(a) - predicted_size || p(a) > object_center_world(a) + predicted_size) inside = false;
            if (inside) vaild_points++;
        }
        return (double)vaild_points / (double)points.size();
    }

    // centroid, bounding radius * 17/16, and the scaled hemisphere (View_Space.hpp:517-558)
    void get_view_space(std::vector<prv::Vector3d>& points) {
        object_center_world = prv::Vector3d(0, 0, 0);
        for (auto& p : points)
            for (int a = 0; a < 3; a++) object_center_world(a) += p(a);
        for (int a = 0; a < 3; a++) object_center_world(a) /= points.size();
        predicted_size = 0.0;
        for (auto& p : points) predicted_size = std::max(predicted_size, (object_center_world - p).norm());
        predicted_size *= 17.0 / 16.0;
        std::cout << "object's pos is (" << object_center_world(0) << "," << object_center_world(1) << "," << object_center_world(2)
                  << ") and size is " << predicted_size << std::endl;
        for (size_t i = 0; i < share_data->pt_sphere.size(); i++) {
            const std::vector<double>& s = share_data->pt_sphere[i];
            if (s[2] < 0) continue;
            const double scale = 1.0 / share_data->pt_norm * share_data->view_space_radius;
            views.push_back(View(prv::Vector3d(s[0] * scale + object_center_world(0), s[1] * scale + object_center_world(1),
                                               s[2] * scale + object_center_world(2))));
        }
        std::cout << "view_space " << views.size() << " getted." << std::endl;
    }

    explicit View_Space(std::shared_ptr<Share_Data>& _share_data) {
        share_data = _share_data;
        num_of_views = share_data->num_of_views;
        now_camera_pose_world = share_data->now_camera_pose_world;
        octomap_resolution = share_data->octomap_resolution;
        std::vector<prv::Vector3d> points;
        points.reserve(share_data->cloud_ground_truth->points.size());
        for (auto& p : share_data->cloud_ground_truth->points) points.push_back(prv::Vector3d(p.x, p.y, p.z));
        get_view_space(points);
        share_data->object_center_world = object_center_world;
        share_data->predicted_size = predicted_size;
    }
};
