// ref_wrap.cpp -- C entry points around the REFERENCE's own src/SurfelFusion.cpp (TEST INFRASTRUCTURE).
//
// oracle/Makefile compiles /root/reference/src/SurfelFusion.cpp where it lies, unmodified, against the stand-in headers of
// oracle/ref_shim_cv/ (the build image has neither OpenCV nor Eigen) into oracle/_ref/libsurfel_ref.so.  What that library
// is: the reference's control flow and scalar arithmetic, line for line; what it is not: Eigen's kernels (the stand-in
// evaluates products left to right and the 4x4 inverse by cofactors) and real threads (the stand-in runs the ten
// slices in order).  tests/test_oracle_ref.py compares the oracle restatement against it; nothing else uses it.
#include <cstdint>
#include <cstring>
#include <vector>

#include <Eigen/Eigen>
#include <opencv2/opencv.hpp>
#include <thread>

#define private public  // the superpixel buffers are private members of the reference class
#include "SurfelFusion.h"
#undef private

extern "C" {

void *ref_surfel_create(int w, int h, float fx, float fy, float cx, float cy, float fuseFar, float fuseNear) {
    return new SurfelFusion(w, h, fx, fy, cx, cy, fuseFar, fuseNear);
}
void ref_surfel_destroy(void *p) { delete (SurfelFusion *)p; }

// gray: 8-bit, `gray_stride` bytes per row, followed by at least 3 * w readable bytes (the reference reads cv::Vec3b on
// it); depth: float metres, dense; membership: int32, half resolution, dense; Twc row-major 4x4.
// local: in/out, n_local surfels (the reference does not resize it); returns the number of new surfels (<= cap_new).
int ref_surfel_fuse(void *p, int ref, uint8_t *gray, int gray_stride, float *depth, int32_t *membership, const float *Twc,
                    Surfel *local, int64_t n_local, Surfel *new_out, int cap_new) {
    SurfelFusion *f = (SurfelFusion *)p;
    const int w = f->imageWidth, h = f->imageHeight;
    cv::Mat image(h, w, CV_8UC1, gray, (size_t)gray_stride), dep(h, w, CV_32FC1, depth);
    cv::Mat mem((h + 1) / 2, (w + 1) / 2, CV_32SC1, membership);
    Eigen::Matrix4f pose;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) pose(i, j) = Twc[4 * i + j];
    std::vector<Surfel> loc(local, local + n_local), nw;
    f->fuseInitializeMap(ref, image, dep, mem, pose, loc, nw);
    if ((int64_t)loc.size() != n_local) return -1;
    memcpy(local, loc.data(), sizeof(Surfel) * (size_t)n_local);
    const int n = (int)nw.size();
    for (int i = 0; i < n && i < cap_new; i++) new_out[i] = nw[i];
    return n;
}

// Timing form: the local map lives in a std::vector owned by this library, as Map::mvLocalSurfels does in the reference, so a
// fuse does not pay for copying the map in and out.  ref_surfel_set_map loads it; ref_surfel_fuse_resident runs
// fuseInitializeMap on it and returns the number of new surfels.
static std::vector<Surfel> g_resident, g_new;
void ref_surfel_set_map(const Surfel *local, int64_t n) { g_resident.assign(local, local + n); }
int64_t ref_surfel_map_size() { return (int64_t)g_resident.size(); }
// the tail of SurfelMapping::fuseMap (src/SurfelMapping.cpp:366-391; that file does not compile here, so these lines are a
// restatement): deleted slots refilled from the back of the deleted list with the new surfels, the rest swap-removed
int64_t ref_surfel_compact_resident() {
    std::vector<int> deleted;
    for (int i = 0; i < (int)g_resident.size(); i++)
        if (g_resident[i].updateTimes == 0) deleted.push_back(i);
    for (size_t i = 0; i < g_new.size(); i++) {
        if (g_new[i].updateTimes == 0) continue;
        if (!deleted.empty()) {
            g_resident[deleted.back()] = g_new[i];
            deleted.pop_back();
        } else {
            g_resident.push_back(g_new[i]);
        }
    }
    while (!deleted.empty()) {
        g_resident[deleted.back()] = g_resident.back();
        deleted.pop_back();
        g_resident.pop_back();
    }
    return (int64_t)g_resident.size();
}
int ref_surfel_fuse_resident(void *p, int ref, uint8_t *gray, int gray_stride, float *depth, int32_t *membership, const float *Twc) {
    SurfelFusion *f = (SurfelFusion *)p;
    const int w = f->imageWidth, h = f->imageHeight;
    cv::Mat image(h, w, CV_8UC1, gray, (size_t)gray_stride), dep(h, w, CV_32FC1, depth);
    cv::Mat mem((h + 1) / 2, (w + 1) / 2, CV_32SC1, membership);
    Eigen::Matrix4f pose;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) pose(i, j) = Twc[4 * i + j];
    f->fuseInitializeMap(ref, image, dep, mem, pose, g_resident, g_new);
    return (int)g_new.size();
}

int ref_surfel_index(void *p, int32_t *out) {
    SurfelFusion *f = (SurfelFusion *)p;
    memcpy(out, f->superpixelIndex.data(), sizeof(int32_t) * f->superpixelIndex.size());
    return (int)f->superpixelIndex.size();
}

// seeds as 18 x 4-byte fields in the order of the oracle's orc_seed / SEED_DTYPE
int ref_surfel_seeds(void *p, void *out) {
    SurfelFusion *f = (SurfelFusion *)p;
    struct Row {
        float x, y, size, normX, normY, normZ, posX, posY, posZ, viewCos, meanDepth, meanIntensity;
        int32_t r, g, b, fused, stable, use;
    };
    Row *o = (Row *)out;
    for (size_t i = 0; i < f->superpixelSeeds.size(); i++) {
        const SurfelFusion::SuperpixelSeed &s = f->superpixelSeeds[i];
        o[i] = Row{s.x, s.y, s.size, s.normX, s.normY, s.normZ, s.posX, s.posY, s.posZ, s.viewCos, s.meanDepth, s.meanIntensity,
                   s.r, s.g, s.b, (int32_t)s.fused, (int32_t)s.stable, (int32_t)s.use};
    }
    return (int)f->superpixelSeeds.size();
}

}  // extern "C"
