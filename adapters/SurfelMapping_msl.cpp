// SurfelMapping_msl.cpp -- device-resident mode of the surfel map: Map::mvLocalSurfels and Map::mvInactiveSurfels
// live on the GPU between keyframes.  Replaces two methods of src/SurfelMapping.cpp (wrap the originals in
// `#ifndef MSL_SURFEL_RESIDENT`):
//   SurfelMapping::moveAddSurfels (:194-304)  -> msl_surfel_move_add   (the pose-graph walk getAddRemovePoses,
//                                                :306-326, and localSurfelsIndexs stay on the host, unchanged)
//   SurfelMapping::fuseMap        (:353-392)  -> msl_surfel_fuse(compact = 1): fuse + initialise + the refill /
//                                                swap-remove tail, no upload or download of the map
// Readers of the host vectors (MapDrawer::DrawSurfels src/MapDrawer.cc:142-143, SurfelMapping::Stop :59-106) see a
// mirror that is refreshed every MSL_SURFEL_MIRROR_EVERY keyframes (default 1; 0 = only in Stop()).
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "Map.h"
#include "SurfelMapping.h"
#include "msl_frontend.h"

msl_surfel_fusion *msl_handle_of(const SurfelFusion *);  // adapters/SurfelFusion_msl.cpp

namespace ORB_SLAM2 {

namespace {
void check(int rc) {
    if (rc != MSL_OK) throw std::runtime_error(msl_last_error());
}
// D2H of both maps into the host vectors the viewer and Stop() read
void mirror_to_host(msl_surfel_fusion *h, Map *map) {
    int64_t n = 0;
    check(msl_surfel_download_map(h, nullptr, 0, &n));
    map->mvLocalSurfels.resize((size_t)n);
    check(msl_surfel_download_map(h, reinterpret_cast<msl_surfel *>(map->mvLocalSurfels.data()), n, &n));
    check(msl_surfel_download_inactive(h, nullptr, 0, &n));
    map->mvInactiveSurfels.resize((size_t)n);
    check(msl_surfel_download_inactive(h, reinterpret_cast<msl_surfel *>(map->mvInactiveSurfels.data()), n, &n));
}
int mirror_every() {
    static const int v = [] {
        const char *e = std::getenv("MSL_SURFEL_MIRROR_EVERY");
        return e ? std::atoi(e) : 1;
    }();
    return v;
}
}  // namespace

void SurfelMapping::moveAddSurfels(int referenceIndex) {
    std::vector<int> posesToAdd, posesToRemove;
    getAddRemovePoses(referenceIndex, posesToAdd, posesToRemove);  // unchanged host logic
    for (int inactiveIndex : posesToRemove) localSurfelsIndexs.erase(inactiveIndex);        // :226
    localSurfelsIndexs.insert(posesToAdd.begin(), posesToAdd.end());                        // :232
    if (posesToAdd.empty() && posesToRemove.empty()) return;
    check(msl_surfel_move_add(msl_handle_of(mSurfelFusion), posesToRemove.data(), (int)posesToRemove.size(),
                              posesToAdd.data(), (int)posesToAdd.size(), nullptr));
}

void SurfelMapping::fuseMap(cv::Mat image, cv::Mat depth, cv::Mat planeMembershipImg, Eigen::Matrix4f poseInput,
                            int referenceIndex) {
    msl_surfel_fusion *h = msl_handle_of(mSurfelFusion);
    CV_Assert(image.type() == CV_8UC1 && depth.type() == CV_32F && depth.isContinuous() &&
              planeMembershipImg.type() == CV_32SC1 && planeMembershipImg.isContinuous());
    const Eigen::Matrix<float, 4, 4, Eigen::RowMajor> Twc = poseInput;
    int64_t stats[4];
    check(msl_surfel_fuse(h, referenceIndex, image.data, (int)image.step, depth.ptr<float>(),
                          planeMembershipImg.ptr<int32_t>(), Twc.data(), nullptr, 0, /*compact=*/1, stats));
    static int since = 0;
    const int every = mirror_every();
    if (every > 0 && ++since >= every) {
        since = 0;
        mirror_to_host(h, mMap);
    }
}

}  // namespace ORB_SLAM2
