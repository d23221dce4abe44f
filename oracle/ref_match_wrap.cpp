// ref_match_wrap.cpp -- C entry points around the REFERENCE's own src/ORBmatcher.cc (TEST INFRASTRUCTURE).
//
// oracle/Makefile compiles /root/reference/src/ORBmatcher.cc (and the vendored Thirdparty/DBoW2/DBoW2/FeatureVector.cpp)
// where they lie, unmodified, into oracle/_ref/libmatch_ref.so: OpenCV = oracle/ref_shim_cv/cvshim.hpp (cv::Mat as a
// container + the three pose products ORBmatcher forms, evaluated by the oracle's cv2-pinned gemm primitives); MapPoint /
// KeyFrame / Frame = the data-only stand-ins of oracle/ref_shim_match/slam_standins.hpp (force-included; the three
// include guards are pre-defined so that the reference's include/ORBmatcher.h finds them).  Every entry point takes the
// SAME flat arguments as the oracle function it checks (oracle/msl_oracle.h: orc_search_by_projection_frame, ...),
// builds the stand-in objects from them, calls the reference's method and flattens the result the same way -- with one
// difference: the reference resets a slot to NULL where the oracle records -3 ("assigned, then removed by the rotation
// check"), so both mean -1 here.  tests/test_oracle_ref.py compares the two; nothing else uses this library.
#include <cstdint>
#include <cstring>
#include <memory>
#include <set>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM2;

namespace {

cv::Mat vec3(const float *v) {
    cv::Mat m(3, 1, CV_32FC1);
    for (int i = 0; i < 3; i++) m.at<float>(i, 0) = v[i];
    return m;
}
cv::Mat mat44(const float *T) {
    cv::Mat m(4, 4, CV_32FC1);
    memcpy(m.data, T, sizeof(float) * 16);
    return m;
}
cv::Mat desc_rows(const uint8_t *d, int n) { return cv::Mat(n, 32, CV_8UC1, (void *)d, 32).clone(); }
cv::Mat desc_row(const uint8_t *d) { return cv::Mat(1, 32, CV_8UC1, (void *)d, 32).clone(); }

void fill_grid(GridGeom &G, const orc_frame_geom *g, const float *xy, const int32_t *octave, int n) {
    G.g = *g;
    G.xy.assign(xy, xy + 2 * (size_t)n);
    G.octave.assign(octave, octave + n);
}

void fill_frame(Frame &F, const orc_frame_geom *g, int n, const float *xy, const int32_t *octave, const float *angle,
                const float *uright, const uint8_t *desc) {
    F.fx = g->fx, F.fy = g->fy, F.cx = g->cx, F.cy = g->cy, F.mbf = g->mbf, F.mb = g->mb;
    F.mnMinX = g->mnMinX, F.mnMaxX = g->mnMaxX, F.mnMinY = g->mnMinY, F.mnMaxY = g->mnMaxY;  // statics, as in the reference
    Frame::mfGridElementWidthInv = g->gridWInv, Frame::mfGridElementHeightInv = g->gridHInv;
    F.N = n;
    F.mnScaleLevels = g->nlevels;
    F.mvScaleFactors.assign(g->scaleFactors, g->scaleFactors + 16);
    F.mfScaleFactor = g->scaleFactors[1];
    F.mfLogScaleFactor = 0;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) {
        cv::KeyPoint &k = F.mvKeysUn[i];
        k.pt.x = xy[2 * i], k.pt.y = xy[2 * i + 1], k.octave = octave ? octave[i] : 0, k.angle = angle ? angle[i] : -1;
    }
    F.mvKeys = F.mvKeysUn;
    F.mvuRight.assign(n, -1.0f);
    if (uright) F.mvuRight.assign(uright, uright + n);
    F.mDescriptors = desc_rows(desc, n);
    F.mvpMapPoints.assign(n, (MapPoint *)nullptr);
    F.mvbOutlier.assign(n, false);
    std::vector<int32_t> zero(n, 0);
    fill_grid(F.grid, g, xy, octave ? octave : zero.data(), n);
}

// slots occupied on entry point at this one (Observations() > 0)
MapPoint *occupied_dummy() {
    static MapPoint d;
    d.id = -2, d.nObs = 1;
    return &d;
}

void read_slots(const Frame &F, int32_t *cur_match) {
    for (int j = 0; j < F.N; j++) cur_match[j] = F.mvpMapPoints[j] ? F.mvpMapPoints[j]->id : -1;
}

void fill_featvec(DBoW2::FeatureVector &fv, int n_nodes, const uint32_t *id, const int32_t *off, const int32_t *feat) {
    for (int k = 0; k < n_nodes; k++)
        for (int j = off[k]; j < off[k + 1]; j++) fv.addFeature(id[k], (unsigned int)feat[j]);
}

}  // namespace

extern "C" {

// ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:835-849
int ref_descriptor_distance(const uint8_t *a, const uint8_t *b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th), src/ORBmatcher.cc:548-678
int ref_search_by_projection_frame(const orc_frame_geom *g, const float Tcw_cur[16], const float Tcw_last[16], float th,
                                   int check_orientation, int n_last, const uint8_t *last_has_mp, const uint8_t *last_outlier,
                                   const uint8_t *last_mp_obs, const float *last_mp_world, const uint8_t *last_mp_desc,
                                   const int32_t *last_octave, const float *last_angle, int n_cur, const float *cur_xy,
                                   const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                                   const uint8_t *cur_desc, const uint8_t *cur_occupied, int32_t *cur_match) {
    Frame Cur, Last;
    fill_frame(Cur, g, n_cur, cur_xy, cur_octave, cur_angle, cur_uright, cur_desc);
    Cur.mTcw = mat44(Tcw_cur);
    for (int j = 0; j < n_cur; j++)
        if (cur_occupied[j]) Cur.mvpMapPoints[j] = occupied_dummy();
    std::vector<float> zxy(2 * (size_t)n_last, 0.0f);
    std::vector<uint8_t> zd(32 * (size_t)n_last + 32, 0);
    fill_frame(Last, g, n_last, zxy.data(), last_octave, last_angle, nullptr, zd.data());
    Last.mTcw = mat44(Tcw_last);
    std::vector<MapPoint> mps(n_last);
    for (int i = 0; i < n_last; i++) {
        Last.mvbOutlier[i] = last_outlier[i] != 0;
        if (!last_has_mp[i]) continue;
        MapPoint &m = mps[i];
        m.id = i, m.nObs = last_mp_obs[i] ? 1 : 0;
        m.mWorldPos = vec3(last_mp_world + 3 * i);
        m.mDescriptor = desc_row(last_mp_desc + 32 * (size_t)i);
        Last.mvpMapPoints[i] = &m;
    }
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int n = matcher.SearchByProjection(Cur, Last, th);
    read_slots(Cur, cur_match);
    return n;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &, th), src/ORBmatcher.cc:40-117
int ref_search_by_projection_points(const orc_frame_geom *g, float th, float nnratio, int n_mp, const uint8_t *mp_valid,
                                    const uint8_t *mp_obs, const float *mp_proj_xyr, const int32_t *mp_level,
                                    const float *mp_viewcos, const uint8_t *mp_desc, int n_cur, const float *cur_xy,
                                    const int32_t *cur_octave, const float *cur_uright, const uint8_t *cur_desc,
                                    const uint8_t *cur_occupied, int32_t *cur_match) {
    Frame F;
    fill_frame(F, g, n_cur, cur_xy, cur_octave, nullptr, cur_uright, cur_desc);
    for (int j = 0; j < n_cur; j++)
        if (cur_occupied[j]) F.mvpMapPoints[j] = occupied_dummy();
    std::vector<MapPoint> mps(n_mp);
    std::vector<MapPoint *> ptrs(n_mp);
    for (int i = 0; i < n_mp; i++) {
        MapPoint &m = mps[i];
        m.id = i, m.nObs = mp_obs[i] ? 1 : 0;
        m.mbTrackInView = mp_valid[i] != 0;
        m.mTrackProjX = mp_proj_xyr[3 * i], m.mTrackProjY = mp_proj_xyr[3 * i + 1], m.mTrackProjXR = mp_proj_xyr[3 * i + 2];
        m.mnTrackScaleLevel = mp_level[i], m.mTrackViewCos = mp_viewcos[i];
        m.mDescriptor = desc_row(mp_desc + 32 * (size_t)i);
        ptrs[i] = &m;
    }
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, ptrs, th);
    read_slots(F, cur_match);
    return n;
}

// ORBmatcher::SearchByProjection(Frame &, KeyFrame *, const set<MapPoint*> &, th, ORBdist), src/ORBmatcher.cc:680-797
int ref_search_by_projection_keyframe(const orc_frame_geom *g, const float Tcw_cur[16], float th, int orb_dist,
                                      int check_orientation, float log_scale_factor, int n_kf, const uint8_t *kf_valid,
                                      const float *kf_mp_world, const uint8_t *kf_mp_desc, const float *kf_mp_dist,
                                      const float *kf_angle, int n_cur, const float *cur_xy, const int32_t *cur_octave,
                                      const float *cur_angle, const uint8_t *cur_desc, const uint8_t *cur_occupied,
                                      int32_t *cur_match) {
    Frame Cur;
    fill_frame(Cur, g, n_cur, cur_xy, cur_octave, cur_angle, nullptr, cur_desc);
    Cur.mTcw = mat44(Tcw_cur);
    Cur.mfLogScaleFactor = log_scale_factor;
    for (int j = 0; j < n_cur; j++)
        if (cur_occupied[j]) Cur.mvpMapPoints[j] = occupied_dummy();
    KeyFrame KF(*g, n_kf, g->nlevels, log_scale_factor);
    KF.mvKeysUn.resize(n_kf);
    KF.mvpMapPoints.assign(n_kf, (MapPoint *)nullptr);
    std::vector<MapPoint> mps(n_kf);
    for (int i = 0; i < n_kf; i++) {
        KF.mvKeysUn[i].angle = kf_angle[i];
        if (!kf_valid[i]) continue;
        MapPoint &m = mps[i];
        m.id = i;
        m.mWorldPos = vec3(kf_mp_world + 3 * i);
        m.mDescriptor = desc_row(kf_mp_desc + 32 * (size_t)i);
        m.mfMinDistance = kf_mp_dist[2 * i], m.mfMaxDistance = kf_mp_dist[2 * i + 1];
        KF.mvpMapPoints[i] = &m;
    }
    std::set<MapPoint *> found;  // kf_valid already excludes sAlreadyFound
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int n = matcher.SearchByProjection(Cur, &KF, found, th, orb_dist);
    read_slots(Cur, cur_match);
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame *, Frame &, vector<MapPoint*> &), src/ORBmatcher.cc:146-255
int ref_search_by_bow(float nnratio, int check_orientation, int n_nodes_kf, const uint32_t *kf_node_id,
                      const int32_t *kf_node_off, const int32_t *kf_node_feat, int n_nodes_f, const uint32_t *f_node_id,
                      const int32_t *f_node_off, const int32_t *f_node_feat, int n_kf, const uint8_t *kf_valid,
                      const uint8_t *kf_desc, const float *kf_angle, int n_f, const uint8_t *f_desc, const float *f_angle,
                      int32_t *f_match) {
    orc_frame_geom g;
    memset(&g, 0, sizeof(g));
    g.gridWInv = g.gridHInv = 1;
    KeyFrame KF(g, n_kf, 8, 1.0f);
    KF.mvKeysUn.resize(n_kf);
    KF.mvpMapPoints.assign(n_kf, (MapPoint *)nullptr);
    KF.mDescriptors = desc_rows(kf_desc, n_kf);
    fill_featvec(KF.mFeatVec, n_nodes_kf, kf_node_id, kf_node_off, kf_node_feat);
    std::vector<MapPoint> mps(n_kf);
    for (int i = 0; i < n_kf; i++) {
        KF.mvKeysUn[i].angle = kf_angle[i];
        mps[i].id = i;
        if (kf_valid[i]) KF.mvpMapPoints[i] = &mps[i];
    }
    Frame F;
    std::vector<float> zxy(2 * (size_t)n_f + 2, 0.0f);
    fill_frame(F, &g, n_f, zxy.data(), nullptr, f_angle, nullptr, f_desc);
    fill_featvec(F.mFeatVec, n_nodes_f, f_node_id, f_node_off, f_node_feat);
    std::vector<MapPoint *> out;
    ORBmatcher matcher(nnratio, check_orientation != 0);
    const int n = matcher.SearchByBoW(&KF, F, out);
    for (int j = 0; j < n_f; j++) f_match[j] = out[j] ? out[j]->id : -1;
    return n;
}

// ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:257-406 (+ CheckDistEpipolarLine :127-144)
int ref_search_for_triangulation(const float F12[9], const float Cw1[3], const float Tcw2[16], const float K2[4], int only_stereo,
                                 int check_orientation, int nlevels, const float *scale_factors2, const float *level_sigma2_2,
                                 int n_nodes1, const uint32_t *node_id1, const int32_t *node_off1, const int32_t *node_feat1,
                                 int n_nodes2, const uint32_t *node_id2, const int32_t *node_off2, const int32_t *node_feat2,
                                 int n1, const uint8_t *has_mp1, const float *uright1, const float *xy1, const float *angle1,
                                 const uint8_t *desc1, int n2, const uint8_t *has_mp2, const float *uright2, const float *xy2,
                                 const int32_t *octave2, const float *angle2, const uint8_t *desc2, int32_t *matches12) {
    orc_frame_geom g;
    memset(&g, 0, sizeof(g));
    g.fx = K2[0], g.fy = K2[1], g.cx = K2[2], g.cy = K2[3];
    KeyFrame KF1(g, n1, nlevels, 1.0f), KF2(g, n2, nlevels, 1.0f);
    MapPoint some;
    auto fill = [&](KeyFrame &K, int n, const uint8_t *has_mp, const float *uright, const float *xy, const int32_t *octave,
                    const float *angle, const uint8_t *desc) {
        K.mvKeysUn.resize(n);
        K.mvpMapPoints.assign(n, (MapPoint *)nullptr);
        for (int i = 0; i < n; i++) {
            cv::KeyPoint &k = K.mvKeysUn[i];
            k.pt.x = xy[2 * i], k.pt.y = xy[2 * i + 1], k.octave = octave ? octave[i] : 0, k.angle = angle[i];
            if (has_mp[i]) K.mvpMapPoints[i] = &some;
        }
        K.mvuRight.assign(uright, uright + n);
        K.mDescriptors = desc_rows(desc, n);
    };
    fill(KF1, n1, has_mp1, uright1, xy1, nullptr, angle1, desc1);
    fill(KF2, n2, has_mp2, uright2, xy2, octave2, angle2, desc2);
    fill_featvec(KF1.mFeatVec, n_nodes1, node_id1, node_off1, node_feat1);
    fill_featvec(KF2.mFeatVec, n_nodes2, node_id2, node_off2, node_feat2);
    KF1.Ow = vec3(Cw1);
    KF2.Tcw = mat44(Tcw2);
    KF2.mvScaleFactors.assign(scale_factors2, scale_factors2 + nlevels);
    KF2.mvLevelSigma2.assign(level_sigma2_2, level_sigma2_2 + nlevels);
    cv::Mat F(3, 3, CV_32FC1);
    memcpy(F.data, F12, sizeof(float) * 9);
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(0.6f, check_orientation != 0);
    const int n = matcher.SearchForTriangulation(&KF1, &KF2, F, pairs, only_stereo != 0);
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    for (size_t k = 0; k < pairs.size(); k++) matches12[pairs[k].first] = (int32_t)pairs[k].second;
    return n;
}

// ORBmatcher::Fuse(KeyFrame *, const vector<MapPoint*> &, th), src/ORBmatcher.cc:408-519.  The KeyFrame holds no map
// points, so every accepted candidate takes the AddObservation branch: fused_idx[i] = the keypoint index it was added
// at, -1 if the map point was not fused (the reference does not expose bestIdx / bestDist of the others).
int ref_fuse(const orc_frame_geom *g, const float Tcw[16], float th, float log_scale_factor, const float *inv_level_sigma2,
             int n_mp, const uint8_t *mp_valid, const float *mp_world, const float *mp_normal, const float *mp_dist,
             const uint8_t *mp_desc, int n_kf, const float *kf_xy, const int32_t *kf_octave, const float *kf_uright,
             const uint8_t *kf_desc, int32_t *fused_idx) {
    KeyFrame KF(*g, n_kf, g->nlevels, log_scale_factor);
    KF.mvKeysUn.resize(n_kf);
    KF.mvpMapPoints.assign(n_kf, (MapPoint *)nullptr);
    for (int i = 0; i < n_kf; i++) {
        cv::KeyPoint &k = KF.mvKeysUn[i];
        k.pt.x = kf_xy[2 * i], k.pt.y = kf_xy[2 * i + 1], k.octave = kf_octave[i];
    }
    KF.mvuRight.assign(kf_uright, kf_uright + n_kf);
    KF.mDescriptors = desc_rows(kf_desc, n_kf);
    KF.mvScaleFactors.assign(g->scaleFactors, g->scaleFactors + 16);
    KF.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + g->nlevels);
    fill_grid(KF.grid, g, kf_xy, kf_octave, n_kf);
    KF.Tcw = mat44(Tcw);
    {   // KeyFrame::SetPose (src/KeyFrame.cc:79-80): Rwc = Rcw.t(); Ow = -Rwc * tcw
        float R[9], t[3], o[3];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) R[3 * r + c] = Tcw[4 * r + c];
            t[r] = Tcw[4 * r + 3];
        }
        orc_cv_neg_rwc_times_t(R, t, o);
        KF.Ow = vec3(o);
    }
    std::vector<MapPoint> mps(n_mp);
    std::vector<MapPoint *> ptrs(n_mp);
    for (int i = 0; i < n_mp; i++) {
        MapPoint &m = mps[i];
        m.id = i;
        m.mWorldPos = vec3(mp_world + 3 * i), m.mNormalVector = vec3(mp_normal + 3 * i);
        m.mfMinDistance = mp_dist[2 * i], m.mfMaxDistance = mp_dist[2 * i + 1];
        m.mDescriptor = desc_row(mp_desc + 32 * (size_t)i);
        ptrs[i] = mp_valid[i] ? &m : nullptr;
    }
    ORBmatcher matcher(0.6f, true);
    const int n = matcher.Fuse(&KF, ptrs, th);
    for (int i = 0; i < n_mp; i++) fused_idx[i] = mps[i].addedAt;
    return n;
}

}  // extern "C"
