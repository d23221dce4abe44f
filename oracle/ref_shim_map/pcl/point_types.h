// Stand-in for <pcl/point_types.h> (TEST INFRASTRUCTURE, see oracle/ref_mapping_wrap.cpp): the point records and the
// container src/SurfelMapping.cpp fills in SurfelMapping::Stop -- containers only, no algorithm.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZRGB {
    float x, y, z;
    uint8_t r, g, b;
};
struct PointSurfel {
    float x, y, z, normal_x, normal_y, normal_z;
    uint8_t r, g, b, a;
    float radius, confidence, curvature;
};
template <typename T> class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    std::vector<T> points;
    void push_back(const T &p) { points.push_back(p); }
};
}  // namespace pcl
