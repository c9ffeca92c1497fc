// Share_Data.hpp -- host-side mirror of the reference's `Share_Data` (PRV_simulation/Share_Data.hpp:204-713)
// restricted to what the ray-cast / coverage / render path touches.  Same constructor signature, same
// public field names, same DefaultConfiguration.yaml keys (all 49 parse unchanged), same gt_path /
// save_path composition, same Hemisphere/<N>.txt loader -- but no OpenCV / PCL / OctoMap: the YAML
// subset is parsed here, clouds are plain vectors with pcl::PointXYZRGB's memory layout and the
// ground-truth OctoMap is a leaf-ordered key list handed to libprv_b200.so.
#pragma once
#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/prv.h"
#include "prv_linalg.hpp"

// rs2_distortion values (Share_Data.hpp:67-76 of the reference); prv_intrinsics::model holds one of these.
enum rs2_distortion {
    RS2_DISTORTION_NONE = 0,
    RS2_DISTORTION_MODIFIED_BROWN_CONRADY = 1,
    RS2_DISTORTION_INVERSE_BROWN_CONRADY = 2,
    RS2_DISTORTION_FTHETA = 3,
    RS2_DISTORTION_BROWN_CONRADY = 4,
    RS2_DISTORTION_KANNALA_BRANDT4 = 5,
    RS2_DISTORTION_COUNT = 6
};
typedef prv_intrinsics rs2_intrinsics;

namespace prv {

typedef prv_point_xyzrgb PointXYZRGB;  // 32-byte pcl::PointXYZRGB image

inline PointXYZRGB make_point(float x, float y, float z, uint8_t r, uint8_t g, uint8_t b) {
    PointXYZRGB p;
    p.x = x; p.y = y; p.z = z; p.w = 1.0f;
    p.r = r; p.g = g; p.b = b; p.a = 255;
    p.pad[0] = p.pad[1] = p.pad[2] = 0.0f;
    return p;
}

struct PointCloud {
    typedef std::shared_ptr<PointCloud> Ptr;
    std::vector<PointXYZRGB> points;
    uint32_t width = 0, height = 1;
    bool is_dense = false;
};

// The slice of octomap::ColorOcTree the path needs: occupied leaf keys in begin_leafs() order + colours.
struct GroundTruthModel {
    double resolution = 0.0;
    std::vector<uint16_t> keys;  // N x 3
    std::vector<uint8_t> rgb;    // N x 3
    explicit GroundTruthModel(double res) : resolution(res) {}
    uint32_t size() const { return (uint32_t)(keys.size() / 3); }
    double getResolution() const { return resolution; }
};

// Minimal reader for OpenCV FileStorage YAML 1.0 files made of flat `key: value` pairs.
class YamlLite {
public:
    bool open(const std::string& path) {
        std::ifstream in(path);
        if (!in.is_open()) return false;
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            const size_t first = line.find_first_not_of(" \t");
            if (first == std::string::npos || line[first] == '#' || line[first] == '%' || line.compare(first, 3, "---") == 0) continue;
            const size_t colon = line.find(':', first);
            if (colon == std::string::npos) continue;
            std::string key = trim(line.substr(first, colon - first));
            std::string val = trim(line.substr(colon + 1));
            if (!val.empty() && val[0] == '"') {
                const size_t q = val.find('"', 1);
                val = val.substr(1, q == std::string::npos ? std::string::npos : q - 1);
            } else {
                const size_t hash = val.find(" #");
                if (hash != std::string::npos) val = trim(val.substr(0, hash));
            }
            kv_[key] = val;
        }
        return true;
    }
    bool has(const std::string& k) const { return kv_.count(k) != 0; }
    // FileStorage semantics: a missing node leaves the destination at its default (0 / empty).
    void get(const std::string& k, std::string& out) const { out = has(k) ? kv_.at(k) : std::string(); }
    void get(const std::string& k, int& out) const { out = has(k) ? (int)std::strtod(kv_.at(k).c_str(), nullptr) : 0; }
    void get(const std::string& k, bool& out) const { out = has(k) ? std::strtod(kv_.at(k).c_str(), nullptr) != 0.0 : false; }
    void get(const std::string& k, double& out) const { out = has(k) ? std::strtod(kv_.at(k).c_str(), nullptr) : 0.0; }
    void get(const std::string& k, float& out) const { out = has(k) ? (float)std::strtod(kv_.at(k).c_str(), nullptr) : 0.0f; }
    size_t size() const { return kv_.size(); }

private:
    static std::string trim(const std::string& s) {
        const size_t a = s.find_first_not_of(" \t");
        if (a == std::string::npos) return std::string();
        const size_t b = s.find_last_not_of(" \t");
        return s.substr(a, b - a + 1);
    }
    std::map<std::string, std::string> kv_;
};

}  // namespace prv

#define RandomIterative 0
#define RandomOneshot 1
#define EnsembleRGB 2
#define EnsembleRGBDensity 3
#define PVBCoverage 4

class Share_Data {
public:
    // configurable paths
    std::string model_path, pcd_file_path, ply_file_path, yaml_file_path, name_of_pcd, nbv_net_path;
    std::string viewspace_path, instant_ngp_path, orginalviews_path, shape_net, pvb_path;

    int num_of_views = 0;
    double cost_weight = 0;
    rs2_intrinsics color_intrinsics;
    double depth_scale = 0;
    double view_space_radius = 0;
    int num_of_thread = 0;

    int process_cnt = -1;
    bool show = false;
    int num_of_max_iteration = 0;

    int vaild_clouds = 0;
    prv::PointCloud::Ptr cloud_pcd;
    prv::PointCloud::Ptr cloud_ground_truth;
    prv::PointCloud::Ptr cloud_final;
    bool move_wait = false;

    std::shared_ptr<prv::GroundTruthModel> ground_truth_model;
    double octomap_resolution = 0;
    double ground_truth_resolution = 0;
    double p_unknown_upper_bound = 0, p_unknown_lower_bound = 0;

    prv::Matrix4d now_camera_pose_world;
    prv::Vector3d object_center_world;
    double predicted_size = 0;

    int method_of_IG = 0;
    double skip_coefficient = 0;
    bool robot_cost_negtive = false;
    int num_of_max_flow_node = 0;
    double interesting_threshold = 0, see_threshold = 0, need_threshold = 0;

    int init_voxels = 0;
    int full_voxels = 0;

    std::string pre_path, gt_path, save_path;

    std::vector<std::vector<double>> pt_sphere;
    double pt_norm = 0;
    double min_z_table = 0;
    std::vector<unsigned long long> view_cases;

    int ray_casting_aabb_scale = 0, num_of_novel_test_views = 0, num_of_neighbors_with_self = 0, num_of_choose = 0;
    int num_of_random_test = 0, num_of_most_cover = 0, cost_on = 0, visit_weight_type = 0;
    double cost_rate = 0, trunc_threshold = 0, approaching_threshold = 0;
    int is_shape_net = 0, coverage_view_num_max = 0, coverage_view_num_add = 0, points_size_cloud = 0, n_steps = 0;
    double object_pixel_rate = 0;
    int id_of_batch = 0, ensemble_num = 0, evaluate = 0;

    bool config_loaded = false;

    Share_Data(std::string _config_file_path, std::string test_name = "", int _num_of_views = -1, int _id_of_batch = -1,
               int test_method = -1) {
        yaml_file_path = _config_file_path;
        color_intrinsics = rs2_intrinsics();
        prv::YamlLite fs;
        config_loaded = fs.open(yaml_file_path);
        if (!config_loaded) std::cout << "can not open config " << yaml_file_path << std::endl;
        load_config(fs);
        if (test_name != "") name_of_pcd = test_name;
        if (test_method != -1) method_of_IG = test_method;
        if (_num_of_views != -1) num_of_views = _num_of_views;
        if (_id_of_batch != -1) id_of_batch = _id_of_batch;
        if (!is_shape_net) {  // reference Share_Data.hpp:406-409
            coverage_view_num_max = 90;
            coverage_view_num_add = 1;
        }
        pcd_file_path = model_path + "PCD/";
        ply_file_path = model_path + "PLY/";
        cloud_pcd.reset(new prv::PointCloud);
        ground_truth_model = std::make_shared<prv::GroundTruthModel>(ground_truth_resolution);
        if (num_of_max_flow_node == -1) num_of_max_flow_node = num_of_views;
        now_camera_pose_world = prv::Matrix4d::Identity();
        cloud_final.reset(new prv::PointCloud);
        cloud_ground_truth.reset(new prv::PointCloud);
        compose_paths(test_method);
        if (method_of_IG == 2) ensemble_num = 2;       // reference Share_Data.hpp:505-510
        else if (method_of_IG == 3) ensemble_num = 5;
        std::cout << "gt_path is: " << gt_path << std::endl;
        std::cout << "save_path is: " << save_path << std::endl;
        load_view_space(num_of_views);
        std::ifstream fin_view_cases(pre_path + "/view_cases.txt");  // vestigial in the reference (:530-536)
        unsigned long long cas;
        while (fin_view_cases >> cas) view_cases.push_back(cas);
    }

    // Hemisphere/<N>.txt loader, reference Share_Data.hpp:517-528 and main.cpp:2355-2367.
    // pt_norm is the norm of ROW 0 and scales every row (kept).
    bool load_view_space(int n) {
        num_of_views = n;
        std::ifstream fin_sphere(viewspace_path + std::to_string(num_of_views) + ".txt");
        pt_sphere.assign(std::max(num_of_views, 0), std::vector<double>(3, 0.0));
        for (int i = 0; i < num_of_views; i++)
            for (int j = 0; j < 3; j++) fin_sphere >> pt_sphere[i][j];
        std::cout << "view space size is: " << pt_sphere.size() << std::endl;
        if (pt_sphere.empty()) {
            pt_norm = 0;
            return false;
        }
        pt_norm = prv::Vector3d(pt_sphere[0][0], pt_sphere[0][1], pt_sphere[0][2]).norm();
        return fin_sphere.is_open();
    }

    prv::Matrix4d get_toward_pose(int toward_state) const {  // reference Share_Data.hpp:591-629
        // each state is a signed axis permutation: rows give (source axis, sign) for x', y', z'
        static const int perm[6][3][2] = {
            {{0, 1}, {1, 1}, {2, 1}},  {{0, 1}, {1, 1}, {2, -1}}, {{2, 1}, {1, 1}, {0, 1}},
            {{2, 1}, {1, 1}, {0, -1}}, {{0, 1}, {2, 1}, {1, 1}},  {{0, 1}, {2, 1}, {1, -1}}};
        prv::Matrix4d pose = prv::Matrix4d::Identity();
        if (toward_state < 0 || toward_state > 5) return pose;
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) pose(r, c) = 0.0;
        }
        // state 3 and 5 in the reference put the -1 in row 2; state 1 likewise
        for (int r = 0; r < 3; r++) pose(r, perm[toward_state][r][0]) = (double)perm[toward_state][r][1];
        return pose;
    }

    void access_directory(std::string cd) {  // reference Share_Data.hpp:639-649 (mkdir -p)
        std::string temp;
        for (size_t i = 0; i < cd.length(); i++) {
            if (cd[i] == '/' && !temp.empty()) make_dir(temp);
            temp += cd[i];
        }
        if (!temp.empty()) make_dir(temp);
    }

private:
    static void make_dir(const std::string& d) {
        struct stat st;
        if (stat(d.c_str(), &st) != 0) mkdir(d.c_str(), 0755);
    }

    void load_config(const prv::YamlLite& fs) {
        fs.get("pre_path", pre_path);
        fs.get("model_path", model_path);
        fs.get("viewspace_path", viewspace_path);
        fs.get("instant_ngp_path", instant_ngp_path);
        fs.get("orginalviews_path", orginalviews_path);
        fs.get("pvb_path", pvb_path);
        fs.get("shape_net", shape_net);
        fs.get("name_of_pcd", name_of_pcd);
        fs.get("method_of_IG", method_of_IG);
        fs.get("num_of_thread", num_of_thread);
        fs.get("octomap_resolution", octomap_resolution);
        fs.get("ground_truth_resolution", ground_truth_resolution);
        fs.get("num_of_neighbors_with_self", num_of_neighbors_with_self);
        fs.get("num_of_choose", num_of_choose);
        fs.get("num_of_random_test", num_of_random_test);
        fs.get("num_of_most_cover", num_of_most_cover);
        fs.get("is_shape_net", is_shape_net);
        fs.get("approaching_threshold", approaching_threshold);
        fs.get("points_size_cloud", points_size_cloud);
        fs.get("n_steps", n_steps);
        fs.get("object_pixel_rate", object_pixel_rate);
        fs.get("id_of_batch", id_of_batch);
        fs.get("evaluate", evaluate);
        fs.get("ensemble_num", ensemble_num);
        fs.get("cost_on", cost_on);
        fs.get("cost_rate", cost_rate);
        fs.get("num_of_max_iteration", num_of_max_iteration);
        fs.get("coverage_view_num_max", coverage_view_num_max);
        fs.get("coverage_view_num_add", coverage_view_num_add);
        fs.get("show", show);
        fs.get("move_wait", move_wait);
        fs.get("nbv_net_path", nbv_net_path);
        fs.get("p_unknown_upper_bound", p_unknown_upper_bound);
        fs.get("p_unknown_lower_bound", p_unknown_lower_bound);
        fs.get("num_of_views", num_of_views);
        fs.get("num_of_novel_test_views", num_of_novel_test_views);
        fs.get("ray_casting_aabb_scale", ray_casting_aabb_scale);
        fs.get("view_space_radius", view_space_radius);
        fs.get("visit_weight_type", visit_weight_type);
        fs.get("trunc_threshold", trunc_threshold);
        fs.get("cost_weight", cost_weight);
        fs.get("robot_cost_negtive", robot_cost_negtive);
        fs.get("skip_coefficient", skip_coefficient);
        fs.get("num_of_max_flow_node", num_of_max_flow_node);
        fs.get("interesting_threshold", interesting_threshold);
        fs.get("see_threshold", see_threshold);
        fs.get("need_threshold", need_threshold);
        fs.get("color_width", color_intrinsics.width);
        fs.get("color_height", color_intrinsics.height);
        fs.get("color_fx", color_intrinsics.fx);
        fs.get("color_fy", color_intrinsics.fy);
        fs.get("color_ppx", color_intrinsics.ppx);
        fs.get("color_ppy", color_intrinsics.ppy);
        fs.get("color_model", color_intrinsics.model);
        // NOTE the reference's name/index quirk (Share_Data.hpp:395-399): k3,p1,p2 land in coeffs[2],[3],[4]
        fs.get("color_k1", color_intrinsics.coeffs[0]);
        fs.get("color_k2", color_intrinsics.coeffs[1]);
        fs.get("color_k3", color_intrinsics.coeffs[2]);
        fs.get("color_p1", color_intrinsics.coeffs[3]);
        fs.get("color_p2", color_intrinsics.coeffs[4]);
        fs.get("depth_scale", depth_scale);
    }

    void compose_paths(int test_method) {  // reference Share_Data.hpp:482-503
        gt_path = pre_path + "Coverage_images/";
        save_path = pre_path + "Compare/";
        if (is_shape_net) {
            gt_path += "ShapeNet";
            save_path += "ShapeNet";
            if (id_of_batch >= 0) {
                gt_path += "_" + std::to_string(id_of_batch);
                save_path += "_" + std::to_string(id_of_batch);
            }
            gt_path += "/";
            save_path += "/";
        }
        gt_path += name_of_pcd;
        save_path += name_of_pcd;
        if (test_method != -1) save_path += "_m" + std::to_string(method_of_IG);
    }
};

inline double pow2(double x) { return x * x; }
