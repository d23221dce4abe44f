// FrameGlue_msl.cc -- Frame::UndistortKeyPoints (src/Frame.cc:437-463) and Frame::ComputeStereoFromRGBD
// (src/Frame.cc:495-513) on the B200 front-end; wrap the two originals in `#ifndef MSL_FRONTEND`.
// Callers are unchanged (the RGB-D Frame constructor, src/Frame.cc:105-112).  The two colour / depth conversions of
// Tracking::GrabImage (src/Tracking.cc:189-207) map to msl_glue_cvt_gray / msl_glue_depth_to_float the same way.
#include <stdexcept>
#include <vector>

#include "Frame.h"
#include "msl_frontend.h"

namespace ORB_SLAM2 {

namespace {
msl_glue *glue(int w, int h) {  // w == 0: any size will do (no depth image involved)
    static msl_glue *g = nullptr;  // Tracking thread only
    static int gw = 0, gh = 0;
    if (g && w > 0 && (gw != w || gh != h)) {
        msl_glue_destroy(g);
        g = nullptr;
    }
    if (!g) {
        gw = w > 0 ? w : 1, gh = h > 0 ? h : 1;
        if (msl_glue_create(gw, gh, 1, 0, &g) != MSL_OK) throw std::runtime_error(msl_last_error());
    }
    return g;
}
std::vector<msl_keypoint> flatten(const std::vector<cv::KeyPoint> &k) {
    std::vector<msl_keypoint> o(k.size());
    for (size_t i = 0; i < k.size(); i++)
        o[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id};
    return o;
}
}  // namespace

void Frame::UndistortKeyPoints() {
    mvKeysUn = mvKeys;  // size, angle, response, octave are copied as in the reference (:455-461)
    if (mDistCoef.at<float>(0) == 0.0 || N == 0) return;
    const float K4[4] = {fx, fy, cx, cy};
    float D5[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < (int)mDistCoef.total() && i < 5; i++) D5[i] = mDistCoef.at<float>(i);
    std::vector<float> xy(2 * (size_t)N);
    const std::vector<msl_keypoint> k = flatten(mvKeys);
    if (msl_glue_keypoints(glue(0, 0), k.data(), N, K4, D5, nullptr, mbf, xy.data(), nullptr, nullptr) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int i = 0; i < N; i++) mvKeysUn[i].pt = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
}

void Frame::ComputeStereoFromRGBD(const cv::Mat &imDepth) {
    // N depth samples at the (distorted) keypoint positions (:495-513).  Done on the host: going through msl_glue_keypoints
    // would upload the whole CV_32F image (1.2 MB at 640x480) and block the tracking thread on a PCIe round trip for N
    // scalar reads.  A frame that is already on the device (msl_glue_upload_frames) uses msl_glue_keypoints_dev instead,
    // chained on the extractor's stream ("frame sets").
    mvuRight.assign(N, -1);
    mvDepth.assign(N, -1);
    CV_Assert(imDepth.type() == CV_32F);
    for (int i = 0; i < N; i++) {
        const cv::Point2f &p = mvKeys[i].pt;
        const float d = imDepth.ptr<float>((int)p.y)[(int)p.x];  // Mat::at<float>(float v, float u): both truncate to int
        if (d > 0) {
            mvDepth[i] = d;
            mvuRight[i] = mvKeysUn[i].pt.x - mbf / d;
        }
    }
}

}  // namespace ORB_SLAM2
