// ref_mapping_wrap.cpp -- C entry points around the REFERENCE's own src/SurfelMapping.cpp (TEST INFRASTRUCTURE).
//
// oracle/Makefile compiles /root/reference/src/SurfelMapping.cpp and src/SurfelFusion.cpp where they lie, unmodified, against
// oracle/ref_shim_cv/ (OpenCV / Eigen stand-ins, the sequential <thread>) and oracle/ref_shim_map/ (Map / MapPlane / pcl
// containers, a cv::FileStorage that answers from a table) into oracle/_ref/libmapping_ref.so.  A keyframe goes through
// SurfelMapping::InsertKeyFrame + ProcessNewKeyFrame (:137-192) exactly as in SurfelMapping::Run: the pose-graph bookkeeping,
// getAddRemovePoses / getDriftfreePoses (:306-352), moveAddSurfels (:194-304) and fuseMap with its compaction tail (:353-392)
// are all the reference's own code.  tests/test_oracle_ref.py drives the oracle (orc_move_add_surfels with the pose lists
// exported here, orc_surfel_fuse, orc_surfel_compact) next to it and requires Map::mvLocalSurfels and
// Map::mvInactiveSurfels to be identical after every keyframe.  Nothing else uses this library.
#include <cstdint>
#include <cstring>
#include <vector>

#define protected public
#include "SurfelMapping.h"
#undef protected

using namespace ORB_SLAM2;

struct RefMapping {
    Map map;
    SurfelMapping *sm;
    std::vector<int> lastAdd, lastRemove;
};

extern "C" {

void *ref_mapping_create(int w, int h, float fx, float fy, float cx, float cy, float far, float near) {
    std::map<std::string, double> &t = cv::FileStorage::table();
    t["Camera.fx"] = fx, t["Camera.fy"] = fy, t["Camera.cx"] = cx, t["Camera.cy"] = cy;
    t["Camera.width"] = w, t["Camera.height"] = h, t["Surfel.distanceFar"] = far, t["Surfel.distanceNear"] = near;
    RefMapping *r = new RefMapping();
    r->sm = new SurfelMapping(&r->map, "settings.yaml");
    return r;
}
void ref_mapping_destroy(void *p) {
    RefMapping *r = (RefMapping *)p;
    delete r->sm->mSurfelFusion;
    delete r->sm;
    delete r;
}

// One keyframe as SurfelMapping::Run processes it.  gray: 8-bit w x h followed by >= 3*w readable bytes; depth float w x h;
// membership int32 half resolution; Twc row-major 4x4; reference_index = the keyframe this one is linked to.
// Returns the size of Map::mvLocalSurfels afterwards.
int64_t ref_mapping_keyframe(void *p, uint8_t *gray, int w, int h, float *depth, int32_t *membership, const float *Twc, int reference_index) {
    RefMapping *r = (RefMapping *)p;
    cv::Mat image(h, w, CV_8UC1, gray, (size_t)w), dep(h, w, CV_32FC1, depth);
    cv::Mat mem((h + 1) / 2, (w + 1) / 2, CV_32SC1, membership), pose(4, 4, CV_32FC1);
    memcpy(pose.data, Twc, sizeof(float) * 16);
    // what moveAddSurfels is about to be told (getAddRemovePoses reads the state ProcessNewKeyFrame sets up first: replay
    // that set-up on copies)
    {
        std::vector<PoseElement> db = r->sm->posesDatabase;
        std::set<int> loc = r->sm->localSurfelsIndexs;
        PoseElement pe;
        const int index = (int)db.size();
        if (!db.empty()) {
            pe.linkedPoseIndex.push_back(reference_index);
            db[reference_index].linkedPoseIndex.push_back(index);
        }
        db.push_back(pe);
        loc.insert(index);
        std::swap(db, r->sm->posesDatabase), std::swap(loc, r->sm->localSurfelsIndexs);
        r->sm->getAddRemovePoses(reference_index, r->lastAdd, r->lastRemove);
        std::swap(db, r->sm->posesDatabase), std::swap(loc, r->sm->localSurfelsIndexs);
    }
    r->sm->InsertKeyFrame(image, dep, mem, pose, reference_index);
    r->sm->ProcessNewKeyFrame();
    return (int64_t)r->map.mvLocalSurfels.size();
}
int ref_mapping_last_lists(void *p, int32_t *add, int32_t *rem, int cap, int32_t *n_rem) {
    RefMapping *r = (RefMapping *)p;
    for (size_t i = 0; i < r->lastAdd.size() && (int)i < cap; i++) add[i] = r->lastAdd[i];
    for (size_t i = 0; i < r->lastRemove.size() && (int)i < cap; i++) rem[i] = r->lastRemove[i];
    *n_rem = (int)r->lastRemove.size();
    return (int)r->lastAdd.size();
}
void ref_mapping_set_local(void *p, const Surfel *s, int64_t n) { ((RefMapping *)p)->map.mvLocalSurfels.assign(s, s + n); }
int64_t ref_mapping_local(void *p, Surfel *out, int64_t cap) {
    const std::vector<Surfel> &v = ((RefMapping *)p)->map.mvLocalSurfels;
    if (out && (int64_t)v.size() <= cap) memcpy(out, v.data(), sizeof(Surfel) * v.size());
    return (int64_t)v.size();
}
int64_t ref_mapping_inactive(void *p, Surfel *out, int64_t cap) {
    const std::vector<Surfel> &v = ((RefMapping *)p)->map.mvInactiveSurfels;
    if (out && (int64_t)v.size() <= cap) memcpy(out, v.data(), sizeof(Surfel) * v.size());
    return (int64_t)v.size();
}

}  // extern "C"
