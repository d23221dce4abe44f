// ORBextractor_msl.cc -- ORB_SLAM2::ORBextractor on the B200 front-end (drop-in for src/ORBextractor.cc).
// Compiles against the reference's unmodified include/ORBextractor.h; Frame.cc / Tracking.cc are untouched:
// Frame::ExtractORB still calls (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors) (src/Frame.cc:175-177).
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include <opencv2/core/core.hpp>

#include "ORBextractor.h"
#include "msl_frontend.h"

namespace ORB_SLAM2 {

namespace {
// The reference class has no spare member and an inline destructor (include/ORBextractor.h:49), so the CUDA
// handle lives in a side table keyed by the object; one extractor exists per process (src/Tracking.cc:121).
struct Slot {
    msl_orb *h = nullptr;
    int w = 0, h_px = 0;
    std::vector<msl_keypoint> kps;
    std::vector<uint8_t> desc;
};
std::mutex g_mu;
std::unordered_map<const ORBextractor *, Slot> g_slots;
}  // namespace

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    // the getters (include/ORBextractor.h:58-82) are served from these vectors exactly as in the reference
    mvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvScaleFactor[0] = 1.0f;
    mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor;
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    mvImagePyramid.resize(nlevels);  // stays empty: the pyramid is device-resident and has no external readers
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask: ignored as in the reference*/,
                              std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors) {
    if (_image.empty()) return;  // src/ORBextractor.cc:815-816
    cv::Mat image = _image.getMat();
    CV_Assert(image.type() == CV_8UC1);
    std::lock_guard<std::mutex> lk(g_mu);
    Slot &s = g_slots[this];
    if (!s.h || s.w != image.cols || s.h_px != image.rows) {
        if (s.h) msl_orb_destroy(s.h);
        msl_orb_params p = {nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST};
        if (msl_orb_create(&p, image.cols, image.rows, 1, 0, &s.h) != MSL_OK) throw std::runtime_error(msl_last_error());
        s.w = image.cols, s.h_px = image.rows;
        const int cap = msl_orb_capacity(s.h);
        s.kps.resize(cap);
        s.desc.resize((size_t)cap * 32);
    }
    int32_t n = 0;
    if (msl_orb_extract(s.h, image.data, (int)image.step, image.step * image.rows, 1, s.kps.data(), s.desc.data(), &n) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    _keypoints.clear();
    _keypoints.reserve(n);
    for (int i = 0; i < n; i++) {
        const msl_keypoint &k = s.kps[i];
        _keypoints.emplace_back(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id);
    }
    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; i++) memcpy(d.ptr(i), s.desc.data() + (size_t)i * 32, 32);
    }
}

// ComputePyramid / ComputeKeyPointsOctTree / DistributeOctTree (protected, include/ORBextractor.h:86-92) are only
// called from operator() in the reference; they are intentionally not defined here.

}  // namespace ORB_SLAM2
