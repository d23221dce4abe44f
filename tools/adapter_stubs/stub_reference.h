#include <memory>
// Stand-ins for the reference's include/{Frame,KeyFrame,MapPoint,ORBmatcher}.h, reduced to the members the bindings use
// (names, types, constness and access levels copied from those headers).  Only for tools/check_adapters.sh.
#pragma once
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>
#include <list>
#include <opencv2/core/core.hpp>
#include <Eigen/Eigen>
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2
namespace ORB_SLAM2 {
class MapPoint;
class KeyFrame;
class Frame {  // include/Frame.h
public:
    void ComputeStereoFromRGBD(const cv::Mat &imDepth);
    cv::Mat mDistCoef;
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
private:
    void UndistortKeyPoints();
};
class ORBextractor {  // include/ORBextractor.h
public:
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor() {}
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors);
    std::vector<cv::Mat> mvImagePyramid;
protected:
    std::vector<cv::Point> pattern;
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
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
class Map;
}  // namespace ORB_SLAM2

// ---- global-namespace classes of the reference
struct Surfel {  // include/Surfel.h
    float px, py, pz;
    float nx, ny, nz;
    float size;
    float color;
    int r, g, b;
    float weight;
    int updateTimes;
    int lastUpdate;
};
#define SP_SIZE 8
class SurfelFusion {  // include/SurfelFusion.h (member order as declared there)
private:
    float fx, fy, cx, cy;
    int imageWidth, imageHeight;
    int spWidth, spHeight;
    float fuseFar, fuseNear;
public:
    SurfelFusion(int width, int height, float _fx, float _fy, float _cx, float _cy, float _fuseFar, float _fuseNear);
    void fuseInitializeMap(const int referenceFrameIndex, const cv::Mat &inputImage, const cv::Mat &inputDepth,
                           const cv::Mat &inputPlaneMembershipImg, const Eigen::Matrix4f &pose,
                           std::vector<Surfel> &localSurfels, std::vector<Surfel> &newSurfels);
};
typedef Eigen::Vector3d VertexType;  // include/PlaneExtractor.h
typedef cv::Vec3d VertexColour;
struct ImagePointCloud {
    std::vector<VertexType> vertices;
    std::vector<VertexColour> verticesColour;
    int w, h;
};
namespace ahc {  // include/peac/AHCTypes.hpp, AHCParamSet.hpp, AHCPlaneSeg.hpp, AHCPlaneFitter.hpp (members the binding uses)
using std::shared_ptr;
struct ParamSet {};
struct NullImage3D {
    int width() { return 0; }
    int height() { return 0; }
    bool get(const int, const int, double &, double &, double &) const { return false; }
};
struct PlaneSeg {
    typedef ahc::shared_ptr<PlaneSeg> shared_ptr;
    int rid;
    double mse;
    double center[3];
    double normal[3];
    int N;
    double curvature;
    bool nouse;
    template <class Image3D>
    PlaneSeg(const Image3D &points, const int root_block_id, const int seed_row, const int seed_col, const int imgWidth,
             const int imgHeight, const int winWidth, const int winHeight, const ParamSet &params);
};
template <class Image3D> struct PlaneFitter {
    int windowWidth, windowHeight;
    ParamSet params;
    std::vector<PlaneSeg::shared_ptr> extractedPlanes;
    cv::Mat membershipImg;
};
}  // namespace ahc
class PlaneDetection {
public:
    ImagePointCloud cloud;
    ahc::PlaneFitter<ImagePointCloud> plane_filter;
    std::vector<std::vector<int> > plane_vertices_;
    cv::Mat seg_img_;
    cv::Mat color_img_;
    int plane_num_;
    bool readColorImage(cv::Mat RGBImg);
    bool readDepthImage(const cv::Mat depthImg, const cv::Mat &K, const float &depthMapFactor);
    void runPlaneDetection();
};
namespace ORB_SLAM2 {
class Map {  // include/Map.h
public:
    std::vector<Surfel> mvLocalSurfels;
    std::vector<Surfel> mvInactiveSurfels;
};
class SurfelMapping {  // include/SurfelMapping.h
protected:
    void moveAddSurfels(int referenceIndex);
    void getAddRemovePoses(int rootIndex, std::vector<int> &poseToAdd, std::vector<int> &poseToRemove);
    void fuseMap(cv::Mat image, cv::Mat depth, cv::Mat planeMembershipImg, Eigen::Matrix4f poseInput, int referenceIndex);
    Map *mMap;
    SurfelFusion *mSurfelFusion;
    std::set<int> localSurfelsIndexs;
};
}  // namespace ORB_SLAM2
