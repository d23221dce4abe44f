// mappoint_wrap.cpp -- TEST HARNESS (tests/test_adapters_on_mock_abi.py): drives ComputeDistinctiveDescriptorsBatch of
// adapters/MapPoint_msl.cc on stand-in MapPoint / KeyFrame objects (oracle/ref_shim_match/slam_standins.hpp, force-included).
// n_kf keyframes with kf_rows[k] descriptor rows each (kf_desc: all rows, keyframe after keyframe), kf_bad flags; map point
// p observes (obs_kf[j], obs_row[j]) for j in [obs_off[p], obs_off[p+1]); mp_bad flags.  Output: the 32-byte descriptor
// every map point ends up with (zeros if the method returned without choosing one).
#include <cstdint>
#include <cstring>
#include <vector>

namespace ORB_SLAM2 {
void ComputeDistinctiveDescriptorsBatch(const std::vector<MapPoint *> &vpMPs);
}
using namespace ORB_SLAM2;

extern "C" int adp_distinctive(int n_kf, const int32_t *kf_rows, const uint8_t *kf_desc, const uint8_t *kf_bad, int n_mp,
                               const int32_t *obs_off, const int32_t *obs_kf, const int32_t *obs_row, const uint8_t *mp_bad,
                               uint8_t *out_desc) {
    orc_frame_geom g;
    memset(&g, 0, sizeof(g));
    std::vector<KeyFrame *> kfs;  // ONE allocation => ascending addresses in keyframe order (mObservations is keyed by pointer)
    char *raw = (char *)operator new(sizeof(KeyFrame) * (size_t)n_kf);
    size_t row0 = 0;
    for (int k = 0; k < n_kf; k++) {
        KeyFrame *kf = new (raw + sizeof(KeyFrame) * (size_t)k) KeyFrame(g, kf_rows[k], 8, 1.0f);
        kf->mDescriptors = cv::Mat(kf_rows[k], 32, CV_8UC1, (void *)(kf_desc + 32 * row0), 32).clone();
        kf->mbBadKF = kf_bad[k] != 0;
        row0 += kf_rows[k];
        kfs.push_back(kf);
    }
    std::vector<MapPoint> mps(n_mp);
    std::vector<MapPoint *> ptrs;
    for (int p = 0; p < n_mp; p++) {
        mps[p].mbBad = mp_bad[p] != 0;
        for (int j = obs_off[p]; j < obs_off[p + 1]; j++) mps[p].mObservations[kfs[obs_kf[j]]] = (size_t)obs_row[j];
        ptrs.push_back(&mps[p]);
    }
    ptrs.push_back(nullptr);  // the helper skips null entries
    ComputeDistinctiveDescriptorsBatch(ptrs);
    for (int p = 0; p < n_mp; p++) {
        if (mps[p].mDescriptor.empty())
            memset(out_desc + 32 * (size_t)p, 0, 32);
        else
            memcpy(out_desc + 32 * (size_t)p, mps[p].mDescriptor.ptr(), 32);
    }
    for (int k = 0; k < n_kf; k++) kfs[k]->~KeyFrame();
    operator delete(raw);
    return 0;
}
