// adapter_wrap.cpp -- TEST HARNESS (tests/test_adapters_on_mock_abi.py): C entry points that drive the reference's CLASS
// SURFACES as implemented by adapters/ (ORBextractor::operator(), PlaneDetection::readDepthImage / runPlaneDetection,
// SurfelFusion::fuseInitializeMap), compiled against the reference's own headers on the stand-in OpenCV / Eigen of
// oracle/ref_shim_cv/ and linked with tests/host_emul/mock_abi.cpp.  Same call sequences as Frame::ExtractORB
// (src/Frame.cc:175-177), Frame::ExtractPlanes (:607-632) and SurfelMapping::fuseMap (src/SurfelMapping.cpp:356-364).
#include <cstdint>
#include <cstring>
#include <vector>

#include <Eigen/Eigen>
#include <opencv2/opencv.hpp>

#include "ORBextractor.h"
#include "PlaneExtractor.h"
#include "SurfelFusion.h"

struct kp_rec {
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

extern "C" {

int adp_orb_extract(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, const uint8_t *gray, int w, int h,
                    int stride, kp_rec *kps, uint8_t *desc, int cap, float *scale_factors) {
    ORB_SLAM2::ORBextractor ext(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    cv::Mat image(h, w, CV_8UC1, (void *)gray, (size_t)stride), descriptors;
    std::vector<cv::KeyPoint> keypoints;
    ext(image, cv::Mat(), keypoints, descriptors);
    const int n = (int)keypoints.size();
    for (int i = 0; i < n && i < cap; i++) {
        const cv::KeyPoint &k = keypoints[i];
        kps[i] = kp_rec{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
        memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
    }
    const std::vector<float> sf = ext.GetScaleFactors();  // a getter Frame's constructor reads (src/Frame.cc:80-86)
    for (int i = 0; i < nlevels; i++) scale_factors[i] = sf[i];
    if (n == 0 && !descriptors.empty()) return -1;
    return n;
}

// Frame::ExtractPlanes' three calls, then what it reads: plane_num_, plane_vertices_[i] (flattened, with offsets),
// extractedPlanes[i]->normal / center, cloud.vertices, and plane_filter.membershipImg (Tracking.cc:228)
int adp_plane_run(const uint16_t *depth, int w, int h, int stride_px, float fx, float fy, float cx, float cy, float factor,
                  int32_t *membership, double *cloud_xyz, double *plane_normal, double *plane_center, int32_t *vertex_off,
                  int32_t *vertices, int cap) {
    PlaneDetection pd;
    cv::Mat color(h, w, CV_8UC3), K(3, 3, CV_32FC1);
    color.setTo(cv::Vec3b(7, 8, 9));
    K.setTo(0.0f);
    K.at<float>(0, 0) = fx, K.at<float>(1, 1) = fy, K.at<float>(0, 2) = cx, K.at<float>(1, 2) = cy, K.at<float>(2, 2) = 1.0f;
    cv::Mat dm(h, w, CV_16UC1, (void *)depth, sizeof(uint16_t) * (size_t)stride_px);
    if (!pd.readColorImage(color) || !pd.readDepthImage(dm, K, factor)) return -1;
    pd.runPlaneDetection();
    const cv::Mat &m = pd.plane_filter.membershipImg;
    for (int y = 0; y < m.rows; y++) memcpy(membership + (size_t)y * m.cols, m.ptr(y), sizeof(int32_t) * (size_t)m.cols);
    for (size_t i = 0; i < pd.cloud.vertices.size(); i++)
        for (int k = 0; k < 3; k++) cloud_xyz[3 * i + k] = pd.cloud.vertices[i][k];
    const int n = pd.plane_num_;
    vertex_off[0] = 0;
    for (int i = 0; i < n && i < cap; i++) {
        const auto &p = pd.plane_filter.extractedPlanes[i];
        for (int k = 0; k < 3; k++) plane_normal[3 * i + k] = p->normal[k], plane_center[3 * i + k] = p->center[k];
        const std::vector<int> &v = pd.plane_vertices_[i];
        memcpy(vertices + vertex_off[i], v.data(), sizeof(int) * v.size());
        vertex_off[i + 1] = vertex_off[i] + (int32_t)v.size();
    }
    if (pd.cloud.verticesColour.empty() || pd.cloud.verticesColour[0][0] != 7) return -2;  // colours read back (src/Frame.cc:619-621)
    return n;
}

int adp_surfel_fuse(int w, int h, float fx, float fy, float cx, float cy, float fuseFar, float fuseNear, int ref, uint8_t *gray,
                    int gray_stride, float *depth, int32_t *membership, const float *Twc, Surfel *local, int64_t n_local,
                    Surfel *new_out, int cap_new) {
    SurfelFusion f(w, h, fx, fy, cx, cy, fuseFar, fuseNear);
    cv::Mat image(h, w, CV_8UC1, gray, (size_t)gray_stride), dep(h, w, CV_32F, depth);
    cv::Mat mem((h + 1) / 2, (w + 1) / 2, CV_32SC1, membership);
    Eigen::Matrix4f pose;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) pose(i, j) = Twc[4 * i + j];
    std::vector<Surfel> loc(local, local + n_local), nw;
    f.fuseInitializeMap(ref, image, dep, mem, pose, loc, nw);
    if ((int64_t)loc.size() != n_local) return -1;
    memcpy(local, loc.data(), sizeof(Surfel) * (size_t)n_local);
    const int n = (int)nw.size();
    for (int i = 0; i < n && i < cap_new; i++) new_out[i] = nw[i];
    return n;
}

}  // extern "C"
