// mapping_adapter_wrap.cpp -- TEST HARNESS (tests/test_adapters_on_mock_abi.py): SurfelMapping::moveAddSurfels and ::fuseMap
// as implemented by adapters/SurfelMapping_msl.cpp (device-resident mode), inside the reference's own SurfelMapping class.
// The reference's src/SurfelMapping.cpp is compiled with its two definitions renamed out of the way (what INTEGRATION.md's
// `#ifndef MSL_SURFEL_RESIDENT` does); its constructor, getAddRemovePoses and getDriftfreePoses are the reference's code,
// SurfelFusion is adapters/SurfelFusion_msl.cpp, the C ABI underneath is tests/host_emul/mock_abi.cpp.  A keyframe is
// processed as SurfelMapping::ProcessNewKeyFrame does (:148-192): pose-graph bookkeeping, moveAddSurfels, fuseMap.
#include <cstdint>
#include <cstring>
#include <vector>

#define protected public
#include "SurfelMapping.h"
#undef protected

using namespace ORB_SLAM2;

struct AdpMapping {
    Map map;
    SurfelMapping *sm;
};

extern "C" {

void *adp_mapping_create(int w, int h, float fx, float fy, float cx, float cy, float far, float near) {
    std::map<std::string, double> &t = cv::FileStorage::table();
    t["Camera.fx"] = fx, t["Camera.fy"] = fy, t["Camera.cx"] = cx, t["Camera.cy"] = cy;
    t["Camera.width"] = w, t["Camera.height"] = h, t["Surfel.distanceFar"] = far, t["Surfel.distanceNear"] = near;
    AdpMapping *r = new AdpMapping();
    r->sm = new SurfelMapping(&r->map, "settings.yaml");
    return r;
}

int64_t adp_mapping_keyframe(void *p, uint8_t *gray, int w, int h, float *depth, int32_t *membership, const float *Twc, int relativeIndex) {
    AdpMapping *r = (AdpMapping *)p;
    SurfelMapping *sm = r->sm;
    cv::Mat image(h, w, CV_8UC1, gray, (size_t)w), dep(h, w, CV_32FC1, depth);
    cv::Mat mem((h + 1) / 2, (w + 1) / 2, CV_32SC1, membership);
    PoseElement poseElement;  // src/SurfelMapping.cpp:161-168
    const int index = (int)sm->posesDatabase.size();
    if (!sm->posesDatabase.empty()) {
        poseElement.linkedPoseIndex.push_back(relativeIndex);
        sm->posesDatabase[relativeIndex].linkedPoseIndex.push_back(index);
    }
    sm->posesDatabase.push_back(poseElement);
    sm->localSurfelsIndexs.insert(index);
    sm->moveAddSurfels(relativeIndex);
    Eigen::Matrix4f pose;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) pose(i, j) = Twc[4 * i + j];
    sm->fuseMap(image, dep, mem, pose, relativeIndex);
    return (int64_t)r->map.mvLocalSurfels.size();  // the host mirror, refreshed every keyframe by default
}
int64_t adp_mapping_local(void *p, Surfel *out, int64_t cap) {
    const std::vector<Surfel> &v = ((AdpMapping *)p)->map.mvLocalSurfels;
    if (out && (int64_t)v.size() <= cap) memcpy(out, v.data(), sizeof(Surfel) * v.size());
    return (int64_t)v.size();
}
int64_t adp_mapping_inactive(void *p, Surfel *out, int64_t cap) {
    const std::vector<Surfel> &v = ((AdpMapping *)p)->map.mvInactiveSurfels;
    if (out && (int64_t)v.size() <= cap) memcpy(out, v.data(), sizeof(Surfel) * v.size());
    return (int64_t)v.size();
}

}  // extern "C"
