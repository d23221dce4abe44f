// mp_standins.hpp -- stand-ins for KeyFrame / Frame / Map as far as src/MapPoint.cc uses them (TEST INFRASTRUCTURE, see
// oracle/ref_mappoint_wrap.cpp).  oracle/Makefile force-includes this header and pre-defines KEYFRAME_H, FRAME_H and MAP_H, so
// that the reference's own include/MapPoint.h, include/ORBmatcher.h and src/MapPoint.cc compile UNMODIFIED: MapPoint itself
// -- ComputeDistinctiveDescriptors, PredictScale, the distance-invariance getters, AddObservation -- is the reference's code.
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <Eigen/Core>
#include <opencv2/core/core.hpp>

namespace cv {
namespace line_descriptor {}  // src/MapPoint.cc names the namespace (line features are out of scope)
}  // namespace cv

using namespace std;

namespace ORB_SLAM2 {
class MapPoint;
class Map {
public:
    std::mutex mMutexPointCreation;
    void EraseMapPoint(MapPoint *) {}
};
class KeyFrame {
public:
    long unsigned int mnId = 0;
    std::vector<float> mvScaleFactors, mvuRight;
    int mnScaleLevels = 8;
    float mfLogScaleFactor = 1;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors, Ow;
    bool bad = false;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    bool isBad() { return bad; }
    void EraseMapPointMatch(const size_t &) {}
    void EraseMapPointMatch(MapPoint *) {}
    void ReplaceMapPointMatch(const size_t &, MapPoint *) {}
};
class Frame {
public:
    long unsigned int mnId = 0;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels = 8;
    float mfLogScaleFactor = 1;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors, Ow;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
};
}  // namespace ORB_SLAM2
