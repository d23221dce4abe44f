// slam_standins.hpp -- stand-ins for the reference's MapPoint / KeyFrame / Frame (TEST INFRASTRUCTURE, see
// oracle/ref_match_wrap.cpp).  include/MapPoint.h, KeyFrame.h and Frame.h pull in the whole of ManhattanSLAM (PCL, g2o,
// line features); src/ORBmatcher.cc uses a thin slice of them: data members and trivial getters.  oracle/Makefile
// force-includes this header and pre-defines the three include guards, so that the reference's own include/ORBmatcher.h
// and src/ORBmatcher.cc compile UNMODIFIED on top of it.
//
// What is data here: every member ORBmatcher.cc reads (names, types and constness as in the reference headers).  What is
// code here, i.e. NOT the reference's: the grid query GetFeaturesInArea (the oracle's orc_features_in_area), IsInImage, the
// two five-line PredictScale and the 0.8 / 1.2 distance getters -- each of which is pinned to the reference's own source on
// its own: the grid and IsInImage against the functions of src/Frame.cc / src/KeyFrame.cc (oracle/_ref/libframe_ref.so),
// PredictScale and the getters against src/MapPoint.cc (oracle/_ref/libmappoint_ref.so), tests/test_oracle_ref.py.
// Everything else ORBmatcher does -- projections, windows, level rules, ratio tests, slot blocking, rotation histograms,
// epipolar test, chi-square gates -- is the reference's own code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;  // the reference's headers rely on it (include/ORBmatcher.h names `pair` and `vector` unqualified)

namespace ORB_SLAM2 {

class KeyFrame;
class Frame;

struct GridGeom {  // what orc_features_in_area needs of a Frame / KeyFrame
    orc_frame_geom g;
    std::vector<float> xy;        // undistorted keypoints, 2 per keypoint
    std::vector<int32_t> octave;
    std::vector<size_t> query(float x, float y, float r, int minLevel, int maxLevel) const {
        std::vector<int32_t> out(octave.size() + 1);
        const int n = orc_features_in_area(&g, xy.data(), octave.data(), (int)octave.size(), x, y, r, minLevel, maxLevel, out.data(),
                                           (int)out.size());
        return std::vector<size_t>(out.begin(), out.begin() + n);
    }
};

class MapPoint {
public:
    // public tracking members of the reference (include/MapPoint.h)
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;

    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    bool IsInKeyFrame(KeyFrame *) { return inKeyFrame; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float &currentDist, KeyFrame *pKF);
    int PredictScale(const float &currentDist, Frame *pF);
    void Replace(MapPoint *pMP) { replacedBy = pMP; }
    void AddObservation(KeyFrame *, size_t idx) { addedAt = (int)idx; }

    // members adapters/MapPoint_msl.cc reaches through a derived struct (protected in include/MapPoint.h:109-141)
    std::map<KeyFrame *, size_t> mObservations;
    std::mutex mMutexFeatures;
    // stand-in state
    int id = -1, nObs = 0, addedAt = -1;
    bool mbBad = false, inKeyFrame = false;
    float mfMinDistance = 0, mfMaxDistance = 0;
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    MapPoint *replacedBy = nullptr;
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0), mTrackViewCos(0) {}
};

class Frame {
public:
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                          const int maxLevel = -1) const {
        return grid.query(x, y, r, minLevel, maxLevel);
    }
    DBoW2::FeatureVector mFeatVec;
    static float fx, fy, cx, cy;  // static in the reference as well (include/Frame.h:117-122); defined in slam_standins.cpp
    float mbf, mb;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw, mK, mDistCoef;
    void UndistortKeyPoints();                         // defined by adapters/FrameGlue_msl.cc (src/Frame.cc:437-463 in the reference)
    void ComputeStereoFromRGBD(const cv::Mat &imDepth);  // (src/Frame.cc:495-513)
    int mnScaleLevels;
    float mfScaleFactor, mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;                          // include/Frame.h:228-231
    static float mfGridElementWidthInv, mfGridElementHeightInv;          // include/Frame.h:202-203
    GridGeom grid;
};

class KeyFrame {
public:
    bool isBad() { return mbBadKF; }
    bool mbBadKF = false;
    cv::Mat GetPose() { return Tcw.clone(); }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint *, const size_t &) {}  // bookkeeping outside the search (the binding's job)
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const { return grid.query(x, y, r, -1, -1); }
    bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }

    const float fx, fy, cx, cy, mbf, mb;
    const float mfGridElementWidthInv, mfGridElementHeightInv;
    const int N;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    const int mnScaleLevels;
    const float mfLogScaleFactor;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    const int mnMinX, mnMinY, mnMaxX, mnMaxY;
    // stand-in state
    cv::Mat Tcw, Ow;
    std::vector<MapPoint *> mvpMapPoints;
    GridGeom grid;
    KeyFrame(const orc_frame_geom &g, int n, int nlevels, float logScaleFactor)
        : fx(g.fx), fy(g.fy), cx(g.cx), cy(g.cy), mbf(g.mbf), mb(g.mb), mfGridElementWidthInv(g.gridWInv),
          mfGridElementHeightInv(g.gridHInv), N(n), mnScaleLevels(nlevels), mfLogScaleFactor(logScaleFactor),
          mnMinX((int)g.mnMinX), mnMinY((int)g.mnMinY), mnMaxX((int)g.mnMaxX), mnMaxY((int)g.mnMaxY) {}
};

// src/MapPoint.cc:334-364 (both overloads): ceil(log(ratio) / mfLogScaleFactor) with float operands, clamped
inline int MapPoint::PredictScale(const float &currentDist, KeyFrame *pKF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0)
        nScale = 0;
    else if (nScale >= pKF->mnScaleLevels)
        nScale = pKF->mnScaleLevels - 1;
    return nScale;
}
inline int MapPoint::PredictScale(const float &currentDist, Frame *pF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0)
        nScale = 0;
    else if (nScale >= pF->mnScaleLevels)
        nScale = pF->mnScaleLevels - 1;
    return nScale;
}

}  // namespace ORB_SLAM2
