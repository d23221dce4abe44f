// glue_wrap.cpp -- TEST HARNESS (tests/test_adapters_on_mock_abi.py): Frame::UndistortKeyPoints and
// Frame::ComputeStereoFromRGBD of adapters/FrameGlue_msl.cc on the stand-in Frame (oracle/ref_shim_match/slam_standins.hpp,
// force-included), with the call order of the RGB-D Frame constructor (src/Frame.cc:105-112); the C ABI underneath is the
// oracle-backed mock below (msl_glue_* -> orc_undistort_keypoints / orc_stereo_from_rgbd).
#include <cstring>
#include <string>
#include <vector>

#include "msl_frontend.h"

using namespace ORB_SLAM2;

struct msl_glue {
    int w, h;
};
static thread_local std::string g_gerr;

extern "C" {
const char *msl_last_error(void) { return g_gerr.c_str(); }
int msl_glue_create(int w, int h, int, int, msl_glue **out) {
    *out = new msl_glue{w, h};
    return MSL_OK;
}
void msl_glue_destroy(msl_glue *g) { delete g; }
int msl_glue_keypoints(msl_glue *g, const msl_keypoint *kps, int n, const float K4[4], const float D5[5], const float *depth, float mbf,
                       float *xy_un, float *uright, float *kdepth) {
    std::vector<float> xy(2 * (size_t)n), un(2 * (size_t)n);
    for (int i = 0; i < n; i++) xy[2 * i] = kps[i].x, xy[2 * i + 1] = kps[i].y;
    const float zero[5] = {0, 0, 0, 0, 0};
    orc_undistort_keypoints(n, xy.data(), K4, D5 ? D5 : zero, un.data());
    if (xy_un) memcpy(xy_un, un.data(), sizeof(float) * un.size());
    if (depth) orc_stereo_from_rgbd(n, xy.data(), un.data(), depth, g->w, mbf, uright, kdepth);
    return MSL_OK;
}

// kps in: n keypoints (x, y, size, angle, response, octave, class_id); out: undistorted keypoints (all seven fields),
// mvuRight, mvDepth
int adp_frame_glue(int n, const msl_keypoint *kps, const float K4[4], const float *D, int nd, const float *depth, int w, int h, float mbf,
                   msl_keypoint *kps_un, float *uright, float *kdepth) {
    Frame F;
    F.N = n;
    F.fx = K4[0], F.fy = K4[1], F.cx = K4[2], F.cy = K4[3], F.mbf = mbf;
    F.mDistCoef = cv::Mat(nd, 1, CV_32FC1);
    for (int i = 0; i < nd; i++) F.mDistCoef.at<float>(i) = D[i];
    for (int i = 0; i < n; i++) F.mvKeys.push_back(cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id));
    F.UndistortKeyPoints();
    cv::Mat dm(h, w, CV_32FC1, (void *)depth);
    F.ComputeStereoFromRGBD(dm);
    if ((int)F.mvKeysUn.size() != n || (int)F.mvuRight.size() != n || (int)F.mvDepth.size() != n) return -1;
    for (int i = 0; i < n; i++) {
        const cv::KeyPoint &k = F.mvKeysUn[i];
        kps_un[i] = msl_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
        uright[i] = F.mvuRight[i], kdepth[i] = F.mvDepth[i];
    }
    return 0;
}
}
