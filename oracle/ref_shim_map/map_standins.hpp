// map_standins.hpp -- stand-ins for what src/SurfelMapping.cpp needs beyond OpenCV / Eigen / SurfelFusion (TEST
// INFRASTRUCTURE, see oracle/ref_mapping_wrap.cpp).  include/SurfelMapping.h includes System.h and Map.h, which pull in the
// whole of ManhattanSLAM; SurfelMapping.cpp uses of them: Map::mvLocalSurfels / mvInactiveSurfels (the two surfel vectors),
// Map::GetAllMapPlanes + MapPlane (only in Stop(), which the checks never call) and the names System.h brings into scope.
// oracle/Makefile force-includes this header and pre-defines SYSTEM_H and MAP_H, so that the reference's own
// include/SurfelMapping.h, include/SurfelFusion.h, include/Surfel.h and src/SurfelMapping.cpp compile UNMODIFIED.
// cv::FileStorage (the constructor reads the camera from the settings file) answers from a table the wrapper fills.
#pragma once
#include <algorithm>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include <opencv2/opencv.hpp>
#include <pcl/point_types.h>

#include "Surfel.h"

using namespace std;  // as System.h's includes do for the reference

namespace cv {
struct FileNode {
    double v;
    operator float() const { return (float)v; }
    operator int() const { return (int)v; }
    operator double() const { return v; }
};
class FileStorage {
public:
    enum { READ = 0 };
    static std::map<std::string, double> &table() {
        static std::map<std::string, double> t;
        return t;
    }
    FileStorage(const std::string &, int) {}
    FileNode operator[](const char *key) const { return FileNode{table().at(key)}; }
};
}  // namespace cv

namespace ORB_SLAM2 {
class MapPlane {
public:
    std::shared_ptr<pcl::PointCloud<pcl::PointXYZRGB>> mvPlanePoints;
    cv::Mat GetWorldPos() { return cv::Mat(); }
};
class Map {
public:
    std::vector<Surfel> mvLocalSurfels;     // include/Map.h
    std::vector<Surfel> mvInactiveSurfels;
    std::vector<MapPlane *> GetAllMapPlanes() { return std::vector<MapPlane *>(); }
};
}  // namespace ORB_SLAM2
