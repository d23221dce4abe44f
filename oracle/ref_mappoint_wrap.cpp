// ref_mappoint_wrap.cpp -- C entry points around the REFERENCE's own src/MapPoint.cc (TEST INFRASTRUCTURE): compiled where it
// lies, unmodified, against oracle/ref_shim_cv/ and the KeyFrame / Frame / Map stand-ins of oracle/ref_shim_mp/ into
// oracle/_ref/libmappoint_ref.so.  Checked against it by tests/test_oracle_ref.py: the oracle's restatement of
// MapPoint::ComputeDistinctiveDescriptors (:210-263) and the PredictScale rule (:334-364) the matcher oracles use.
#include <cstdint>
#include <cstring>
#include <vector>

#include "MapPoint.h"
#include "ORBmatcher.h"
#include "msl_oracle.h"

using namespace ORB_SLAM2;

// the one ORBmatcher method src/MapPoint.cc calls (src/ORBmatcher.cc:835-849, pinned on its own by libmatch_ref.so)
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) { return orc_descriptor_distance(a.ptr(), b.ptr()); }

extern "C" {

// n_kf keyframes with kf_rows[k] descriptor rows each (kf_desc: all rows), kf_bad flags; map point p observes
// (obs_kf[j], obs_row[j]) for j in [obs_off[p], obs_off[p+1]).  Output: the 32-byte descriptor every point ends up with
// (zeros if ComputeDistinctiveDescriptors returned without choosing one).
int ref_distinctive(int n_kf, const int32_t *kf_rows, const uint8_t *kf_desc, const uint8_t *kf_bad, int n_mp, const int32_t *obs_off,
                    const int32_t *obs_kf, const int32_t *obs_row, uint8_t *out_desc) {
    Map map;
    std::vector<KeyFrame> kfs(n_kf);  // one allocation: ascending addresses in keyframe order (mObservations is keyed by pointer)
    size_t row0 = 0;
    for (int k = 0; k < n_kf; k++) {
        kfs[k].mnId = k;
        kfs[k].mDescriptors = cv::Mat(kf_rows[k], 32, CV_8UC1, (void *)(kf_desc + 32 * row0), 32).clone();
        kfs[k].mvuRight.assign(kf_rows[k], -1.0f);
        kfs[k].bad = kf_bad[k] != 0;
        row0 += kf_rows[k];
    }
    cv::Mat pos(3, 1, CV_32FC1);
    pos.setTo(0.0f);
    for (int p = 0; p < n_mp; p++) {
        MapPoint mp(pos, &kfs[0], &map);
        for (int j = obs_off[p]; j < obs_off[p + 1]; j++) mp.AddObservation(&kfs[obs_kf[j]], (size_t)obs_row[j]);
        mp.ComputeDistinctiveDescriptors();
        cv::Mat d = mp.GetDescriptor();
        if (d.empty())
            memset(out_desc + 32 * (size_t)p, 0, 32);
        else
            memcpy(out_desc + 32 * (size_t)p, d.ptr(), 32);
    }
    return 0;
}

// MapPoint::PredictScale(currentDist, KeyFrame*) and (.., Frame*) for a point whose mfMaxDistance is exactly max_dist (made
// through the Frame constructor: camera at the origin, point at (0, 0, max_dist), octave 0 with scale factor 1); also the
// two distance-invariance getters.  out: n x 2 levels; inv: {GetMinDistanceInvariance, GetMaxDistanceInvariance}.
int ref_predict_scale(float max_dist, float log_scale_factor, int n_levels, const float *top_scale_factor, int n, const float *dist,
                      int32_t *out, float *inv) {
    Map map;
    Frame F;
    F.mnScaleLevels = n_levels, F.mfLogScaleFactor = log_scale_factor;
    F.mvScaleFactors.assign(n_levels, 1.0f);
    F.mvScaleFactors[n_levels - 1] = *top_scale_factor;
    F.mvKeysUn.resize(1);
    F.mvKeysUn[0].octave = 0;
    F.mDescriptors = cv::Mat(1, 32, CV_8UC1);
    F.mDescriptors.setTo((uchar)0);
    F.Ow = cv::Mat(3, 1, CV_32FC1);
    F.Ow.setTo(0.0f);
    cv::Mat pos(3, 1, CV_32FC1);
    pos.setTo(0.0f);
    pos.at<float>(2, 0) = max_dist;
    MapPoint mp(pos, &map, &F, 0);
    KeyFrame K;
    K.mnScaleLevels = n_levels, K.mfLogScaleFactor = log_scale_factor;
    for (int i = 0; i < n; i++) out[2 * i] = mp.PredictScale(dist[i], &K), out[2 * i + 1] = mp.PredictScale(dist[i], &F);
    inv[0] = mp.GetMinDistanceInvariance(), inv[1] = mp.GetMaxDistanceInvariance();
    return 0;
}

}  // extern "C"
