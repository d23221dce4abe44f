// MapPoint_msl.cc -- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263) for MANY map points at once.
//
// The reference calls the method once per map point inside loops over a KeyFrame's map points
// (src/LocalMapping.cc:125-141 ProcessNewKeyFrame, :573-581 SearchInNeighbors).  One point is a few dozen 256-bit
// distances -- far too little for a kernel launch -- so the per-point method of src/MapPoint.cc stays as it is, and the
// two loops collect their points and call this helper once after the loop instead:
//
//     std::vector<MapPoint *> touched;                       // src/LocalMapping.cc:573-581
//     for (auto pMP : vpMapPointMatches)
//         if (pMP && !pMP->isBad()) { touched.push_back(pMP); pMP->UpdateNormalAndDepth(); }
//     ORB_SLAM2::ComputeDistinctiveDescriptorsBatch(touched);
//
// The result per point is the one the reference's method computes (same observations, same order, same BestIdx).
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "KeyFrame.h"
#include "MapPoint.h"
#include "msl_frontend.h"

namespace ORB_SLAM2 {

namespace {
// mObservations, mDescriptor, mbBad and mMutexFeatures are protected (include/MapPoint.h:109-141).  A derived struct may
// name them, which yields plain pointers-to-member of MapPoint -- no change to the reference header.
struct MapPointAccess : MapPoint {
    static std::map<KeyFrame *, size_t> MapPoint::*obs() { return &MapPointAccess::mObservations; }
    static cv::Mat MapPoint::*desc() { return &MapPointAccess::mDescriptor; }
    static bool MapPoint::*bad() { return &MapPointAccess::mbBad; }
    static std::mutex MapPoint::*mutex() { return &MapPointAccess::mMutexFeatures; }
};
msl_matcher *batch_matcher() {
    struct Holder {
        msl_matcher *m = nullptr;
        ~Holder() {
            if (m) msl_matcher_destroy(m);
        }
    };
    static thread_local Holder h;  // one handle per calling thread (LocalMapping)
    if (!h.m && msl_matcher_create(4096, 4096, 1, 0, &h.m) != MSL_OK) throw std::runtime_error(msl_last_error());
    return h.m;
}
}  // namespace

void ComputeDistinctiveDescriptorsBatch(const std::vector<MapPoint *> &vpMPs) {
    std::vector<int32_t> off(1, 0);
    std::vector<uint8_t> desc;
    std::vector<MapPoint *> pts;
    std::vector<cv::Mat> rows;  // the observed descriptor rows, point after point (headers only: no pixel copy)
    for (size_t k = 0; k < vpMPs.size(); k++) {
        MapPoint *pMP = vpMPs[k];
        if (!pMP) continue;
        std::map<KeyFrame *, size_t> observations;
        {
            std::unique_lock<std::mutex> lock(pMP->*MapPointAccess::mutex());  // :219-224
            if (pMP->*MapPointAccess::bad()) continue;
            observations = pMP->*MapPointAccess::obs();
        }
        if (observations.empty()) continue;
        const size_t first = rows.size();
        for (std::map<KeyFrame *, size_t>::iterator mit = observations.begin(); mit != observations.end(); ++mit)
            if (!mit->first->isBad()) rows.push_back(mit->first->mDescriptors.row((int)mit->second));  // :228-234
        if (rows.size() == first) continue;  // :236-237
        pts.push_back(pMP);
        off.push_back((int32_t)rows.size());
    }
    if (pts.empty()) return;
    desc.resize(rows.size() * 32);
    for (size_t r = 0; r < rows.size(); r++) memcpy(&desc[r * 32], rows[r].ptr(), 32);
    std::vector<int32_t> best(pts.size());
    if (msl_distinctive_descriptors(batch_matcher(), (int)pts.size(), off.data(), desc.data(), best.data(), nullptr) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (size_t k = 0; k < pts.size(); k++) {
        std::unique_lock<std::mutex> lock(pts[k]->*MapPointAccess::mutex());  // :255-258
        pts[k]->*MapPointAccess::desc() = rows[off[k] + best[k]].clone();
    }
}

}  // namespace ORB_SLAM2
