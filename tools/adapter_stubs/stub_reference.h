// Stand-ins for the reference's include/{Frame,KeyFrame,MapPoint,ORBmatcher}.h, reduced to the members the bindings use
// (names, types, constness and access levels copied from those headers).  Only for tools/check_adapters.sh.
#pragma once
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>
#include <opencv2/core/core.hpp>
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2
namespace ORB_SLAM2 {
class MapPoint;
class KeyFrame;
class Frame {  // include/Frame.h
public:
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float fx, fy, cx, cy;
    float mb, mbf;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    cv::Mat mTcw;
    int mnScaleLevels;
    float mfScaleFactor, mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
class KeyFrame {  // include/KeyFrame.h
public:
    cv::Mat GetPose();
    cv::Mat GetCameraCenter();
    void AddMapPoint(MapPoint *pMP, const size_t &idx);
    std::vector<MapPoint *> GetMapPointMatches();
    MapPoint *GetMapPoint(const size_t &idx);
    bool isBad();
    const float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    const float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
    const int N = 0;
    const std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    const std::vector<float> mvuRight, mvDepth;
    const cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    const int mnScaleLevels = 0;
    const float mfScaleFactor = 0, mfLogScaleFactor = 0;
    const std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    const int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
};
class MapPoint {  // include/MapPoint.h
public:
    cv::Mat GetWorldPos();
    cv::Mat GetNormal();
    int Observations();
    void AddObservation(KeyFrame *pKF, size_t idx);
    bool IsInKeyFrame(KeyFrame *pKF);
    bool isBad();
    void Replace(MapPoint *pMP);
    cv::Mat GetDescriptor();
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
protected:
    std::map<KeyFrame *, size_t> mObservations;
    cv::Mat mDescriptor;
    bool mbBad;
    float mfMinDistance, mfMaxDistance;
    std::mutex mMutexFeatures;
};
class ORBmatcher {  // include/ORBmatcher.h
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th);
    int SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th,
                           const int ORBdist);
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches);
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
                               std::vector<std::pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo);
    int Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th = 3.0);
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
protected:
    float mfNNratio;
    bool mbCheckOrientation;
};
}  // namespace ORB_SLAM2
