// SurfelFusion_msl.cpp -- SurfelFusion on the B200 front-end (drop-in for src/SurfelFusion.cpp).
// SurfelMapping::fuseMap (src/SurfelMapping.cpp:353-364) is unchanged.  This is the exact drop-in: the host
// vector Map::mvLocalSurfels stays authoritative, so every call uploads it, fuses on the device and downloads it
// again -- the upload whole (the host mutates the vector between calls: the fuseMap tail, moveAddSurfels), the download only for
// the surfels the call changed (msl_surfel_download_changed: ~30 % of the map).  INTEGRATION.md describes the device-resident mode
// (compaction and SurfelMapping::moveAddSurfels on the device, no round trip): adapters/SurfelMapping_msl.cpp.
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "SurfelFusion.h"
#include "msl_frontend.h"

static_assert(sizeof(Surfel) == sizeof(msl_surfel), "include/Surfel.h layout == msl_surfel");

namespace {
std::mutex g_mu;
std::unordered_map<const SurfelFusion *, msl_surfel_fusion *> g_h;  // one instance, SurfelMapping thread only
}

// handle of a SurfelFusion object, for adapters/SurfelMapping_msl.cpp (device-resident mode)
msl_surfel_fusion *msl_handle_of(const SurfelFusion *f) {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_h.at(f);
}

// The reference class has no destructor to hook (include/SurfelFusion.h): the owner releases the device map explicitly --
// call it from SurfelMapping's destructor / Stop().  Without it a SurfelFusion object that is destroyed
// leaks its device map, and an object later allocated at the same address would find a stale table entry (its constructor
// overwrites it).
void msl_release_of(const SurfelFusion *f) {
    msl_surfel_fusion *h = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_h.find(f);
        if (it == g_h.end()) return;
        h = it->second;
        g_h.erase(it);
    }
    msl_surfel_destroy(h);
}

// Device map capacity in surfels (56 B each + 8 B of queue / compaction planes): MSL_SURFEL_CAPACITY, default 32 Mi (2.4 GB).
static long long surfel_capacity() {
    if (const char *e = std::getenv("MSL_SURFEL_CAPACITY")) {
        const long long v = std::atoll(e);
        if (v > 0) return v;
    }
    return 32ll << 20;
}

SurfelFusion::SurfelFusion(int width, int height, float _fx, float _fy, float _cx, float _cy, float _fuseFar, float _fuseNear)
    : fx(_fx), fy(_fy), cx(_cx), cy(_cy), imageWidth(width), imageHeight(height), spWidth(width / SP_SIZE),
      spHeight(height / SP_SIZE), fuseFar(_fuseFar), fuseNear(_fuseNear) {  // declaration order of include/SurfelFusion.h:60-63
    msl_surfel_fusion *h = nullptr;
    if (msl_surfel_create(width, height, _fx, _fy, _cx, _cy, _fuseFar, _fuseNear, surfel_capacity(), 0, &h) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    msl_surfel_fusion *stale = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_h.find(this);
        if (it != g_h.end()) stale = it->second;  // an earlier object at this address that was never released
        g_h[this] = h;
    }
    if (stale) msl_surfel_destroy(stale);
}

void SurfelFusion::fuseInitializeMap(const int referenceFrameIndex, const cv::Mat &inputImage, const cv::Mat &inputDepth,
                                     const cv::Mat &inputPlaneMembershipImg, const Eigen::Matrix4f &pose,
                                     std::vector<Surfel> &localSurfels, std::vector<Surfel> &newSurfels) {
    msl_surfel_fusion *h;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        h = g_h.at(this);
    }
    CV_Assert(inputImage.type() == CV_8UC1 && inputDepth.type() == CV_32F && inputDepth.isContinuous() &&
              inputPlaneMembershipImg.type() == CV_32SC1 && inputPlaneMembershipImg.isContinuous());
    const Eigen::Matrix<float, 4, 4, Eigen::RowMajor> Twc = pose;  // the ABI takes row-major
    if (msl_surfel_upload_map(h, reinterpret_cast<const msl_surfel *>(localSurfels.data()), (int64_t)localSurfels.size()) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    newSurfels.resize((size_t)spWidth * spHeight);
    int64_t stats[4];
    if (msl_surfel_fuse(h, referenceFrameIndex, inputImage.data, (int)inputImage.step, inputDepth.ptr<float>(),
                        inputPlaneMembershipImg.ptr<int32_t>(), Twc.data(), reinterpret_cast<msl_surfel *>(newSurfels.data()),
                        (int)newSurfels.size(), /*compact=*/0, stats) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    newSurfels.resize((size_t)stats[0]);
    // dirty download: only the surfels this call updated (lastUpdate == referenceFrameIndex) or deleted (updateTimes == 0)
    // differ from what was uploaded -- about 30 % of the map (msl_surfel_download_changed); patched into the host vector
    static thread_local std::vector<int32_t> idx;
    static thread_local std::vector<msl_surfel> rec;
    if (idx.size() < localSurfels.size()) idx.resize(localSurfels.size()), rec.resize(localSurfels.size());
    int64_t n = 0;
    if (msl_surfel_download_changed(h, referenceFrameIndex, idx.data(), rec.data(), (int64_t)idx.size(), &n) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    Surfel *dst = localSurfels.data();
    const Surfel *src = reinterpret_cast<const Surfel *>(rec.data());
    for (int64_t k = 0; k < n; k++) dst[idx[k]] = src[k];
}
