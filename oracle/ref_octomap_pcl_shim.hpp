// ref_octomap_pcl_shim.hpp -- TEST INFRASTRUCTURE.  The slice of the OctoMap / PCL API that the reference's
// Perception_3D::precept_thread_process (PRV_simulation/main.cpp:238-284) and project_pixel_to_ray_end
// (Share_Data.hpp:719-726) use, so that those two functions can be compiled FROM WHERE THEY LIE without OctoMap / PCL
// (oracle/Makefile target `ref`).  What is pinned: the reference's per-voxel logic (projection, the '>' bounds test, the
// float->int truncation at the call, ray end / direction arithmetic, result and colour handling).  What is NOT: castRay,
// coordToKeyChecked and search are answered by the oracle's own restatement of OctoMap (prv_oracle.h), which stays
// unpinned; octomath::Vector3 is restated here (float components, component-wise float subtraction, exact ==).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <vector>

#include "prv_oracle.h"

namespace octomap {

class point3d {  // octomath::Vector3
  public:
    point3d(float x = 0.0f, float y = 0.0f, float z = 0.0f) : d{x, y, z} {}
    float x() const { return d[0]; }
    float y() const { return d[1]; }
    float z() const { return d[2]; }
    float operator()(int i) const { return d[i]; }
    point3d operator-(const point3d& o) const {
        point3d r(*this);
        r.d[0] -= o.d[0];
        r.d[1] -= o.d[1];
        r.d[2] -= o.d[2];
        return r;
    }
    bool operator==(const point3d& o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2]; }
    const float* data() const { return d; }
    float d[3];
};

struct OcTreeKey {
    uint16_t k[3];
    OcTreeKey() : k{0, 0, 0} {}
    bool operator<(const OcTreeKey& o) const { return k[0] != o.k[0] ? k[0] < o.k[0] : (k[1] != o.k[1] ? k[1] < o.k[1] : k[2] < o.k[2]); }
};

class ColorOcTreeNode {
  public:
    struct Color {
        uint8_t r, g, b;
    };
    Color getColor() const { return c; }
    Color c;
};

class ColorOcTree {
  public:
    ColorOcTree(const orc_map* m, double resolution) : map(m), res(resolution) {
        const uint32_t n = orc_map_size(m);
        keys.resize(3 * (size_t)n);
        std::vector<uint8_t> rgb(3 * (size_t)n);
        orc_map_keys(m, keys.data());
        orc_map_rgb(m, rgb.data());
        for (uint32_t i = 0; i < n; i++) {
            OcTreeKey k;
            for (int a = 0; a < 3; a++) k.k[a] = keys[3 * (size_t)i + a];
            ColorOcTreeNode node;
            node.c = ColorOcTreeNode::Color{rgb[3 * (size_t)i], rgb[3 * (size_t)i + 1], rgb[3 * (size_t)i + 2]};
            nodes[k] = node;
        }
    }
    bool coordToKeyChecked(double x, double y, double z, OcTreeKey& key) const {
        return orc_coord_to_key(x, res, &key.k[0]) && orc_coord_to_key(y, res, &key.k[1]) && orc_coord_to_key(z, res, &key.k[2]);
    }
    bool coordToKeyChecked(const point3d& p, OcTreeKey& key) const { return coordToKeyChecked(p(0), p(1), p(2), key); }
    point3d keyToCoord(const OcTreeKey& key) const {
        return point3d((float)orc_key_to_coord(key.k[0], res), (float)orc_key_to_coord(key.k[1], res), (float)orc_key_to_coord(key.k[2], res));
    }
    bool castRay(const point3d& origin, const point3d& direction, point3d& end, bool ignoreUnknownCells, double maxRange) const {
        uint32_t rank = 0;
        float e[3];
        const int found = orc_cast_ray(map, origin.data(), direction.data(), ignoreUnknownCells ? 1 : 0, maxRange, e, &rank, nullptr);
        end = point3d(e[0], e[1], e[2]);
        return found != 0;
    }
    ColorOcTreeNode* search(const OcTreeKey& key) {
        auto it = nodes.find(key);
        return it == nodes.end() ? nullptr : &it->second;
    }
    const orc_map* map;
    double res;
    std::vector<uint16_t> keys;  // leaf order
    std::map<OcTreeKey, ColorOcTreeNode> nodes;
};

}  // namespace octomap

namespace pcl {
struct PointXYZRGB {
    float x = 0, y = 0, z = 0;
    uint8_t b = 0, g = 0, r = 0, a = 255;
};
template <typename T>
struct PointCloud {
    std::vector<T> points;
    bool is_dense = true;
};
}  // namespace pcl
