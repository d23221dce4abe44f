// ORBmatcher_msl.cc -- the tracking-time searches of ORB_SLAM2::ORBmatcher on the B200 front-end.
// Build together with the reference's src/ORBmatcher.cc compiled with -DMSL_FRONTEND, where the four
// definitions below (three SearchByProjection overloads + DescriptorDistance) are wrapped in
// `#ifndef MSL_FRONTEND` (INTEGRATION.md shows the patch); every other method (SearchByBoW, SearchForTriangulation, Fuse, ...) stays as is.  Callers are unchanged
// (src/Tracking.cc:956,963,1253,1262,1693,2006,2019).
#include <set>
#include <stdexcept>
#include <vector>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#include "msl_frontend.h"

namespace ORB_SLAM2 {

namespace {
msl_matcher *matcher() {
    static msl_matcher *m = nullptr;  // Tracking thread only (one ORBmatcher call at a time)
    if (!m && msl_matcher_create(4096, 4096, 1, 0, &m) != MSL_OK) throw std::runtime_error(msl_last_error());
    return m;
}
msl_frame_geom geom_of(const Frame &F) {
    msl_frame_geom g = {};
    g.fx = F.fx, g.fy = F.fy, g.cx = F.cx, g.cy = F.cy;
    g.mnMinX = Frame::mnMinX, g.mnMinY = Frame::mnMinY, g.mnMaxX = Frame::mnMaxX, g.mnMaxY = Frame::mnMaxY;
    g.gridWInv = Frame::mfGridElementWidthInv, g.gridHInv = Frame::mfGridElementHeightInv;
    g.mb = F.mb, g.mbf = F.mbf;
    g.nlevels = (int)F.mvScaleFactors.size();
    for (int i = 0; i < g.nlevels && i < 16; i++) g.scaleFactors[i] = F.mvScaleFactors[i];
    return g;
}
struct CurArrays {
    std::vector<float> xy, angle, uright;
    std::vector<int32_t> octave;
    std::vector<uint8_t> occ;
    explicit CurArrays(const Frame &F) : xy(2 * F.N), angle(F.N), uright(F.N), octave(F.N), occ(F.N) {
        for (int j = 0; j < F.N; j++) {
            xy[2 * j] = F.mvKeysUn[j].pt.x, xy[2 * j + 1] = F.mvKeysUn[j].pt.y;
            angle[j] = F.mvKeysUn[j].angle, octave[j] = F.mvKeysUn[j].octave, uright[j] = F.mvuRight[j];
            occ[j] = F.mvpMapPoints[j] && F.mvpMapPoints[j]->Observations() > 0;
        }
    }
};
}  // namespace

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {  // src/ORBmatcher.cc:835-849
    const uint32_t *pa = a.ptr<uint32_t>(), *pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th) {  // :548-678
    const int nl = LastFrame.N, nc = CurrentFrame.N;
    std::vector<uint8_t> has(nl), outl(nl), obs(nl), desc((size_t)nl * 32);
    std::vector<float> world(3 * (size_t)nl), ang(nl);
    std::vector<int32_t> oct(nl);
    for (int i = 0; i < nl; i++) {
        MapPoint *p = LastFrame.mvpMapPoints[i];
        has[i] = p != nullptr, outl[i] = LastFrame.mvbOutlier[i];
        oct[i] = LastFrame.mvKeys[i].octave, ang[i] = LastFrame.mvKeysUn[i].angle;
        if (!p) continue;
        obs[i] = p->Observations() > 0;
        cv::Mat x = p->GetWorldPos();
        world[3 * i] = x.at<float>(0), world[3 * i + 1] = x.at<float>(1), world[3 * i + 2] = x.at<float>(2);
        memcpy(&desc[(size_t)i * 32], p->GetDescriptor().ptr(), 32);
    }
    CurArrays C(CurrentFrame);
    std::vector<int32_t> match(nc);
    int32_t nmatches = 0;
    const msl_frame_geom g = geom_of(CurrentFrame);
    cv::Mat Tc, Tl;
    CurrentFrame.mTcw.convertTo(Tc, CV_32F);
    LastFrame.mTcw.convertTo(Tl, CV_32F);
    if (msl_search_by_projection_frame(matcher(), &g, Tc.ptr<float>(), Tl.ptr<float>(), th, mbCheckOrientation, nl, has.data(),
                                       outl.data(), obs.data(), world.data(), desc.data(), oct.data(), ang.data(), nc,
                                       C.xy.data(), C.octave.data(), C.angle.data(), C.uright.data(),
                                       CurrentFrame.mDescriptors.ptr(), C.occ.data(), match.data(), &nmatches) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int j = 0; j < nc; j++) {
        if (match[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[match[j]];
        else if (match[j] == -3) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint *>(NULL);
    }
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th) {  // :40-117
    const int nm = (int)vpMapPoints.size(), nc = F.N;
    std::vector<uint8_t> valid(nm), obs(nm), desc((size_t)nm * 32);
    std::vector<float> proj(3 * (size_t)nm), vcos(nm);
    std::vector<int32_t> lvl(nm);
    for (int k = 0; k < nm; k++) {
        MapPoint *p = vpMapPoints[k];
        valid[k] = p->mbTrackInView && !p->isBad();
        if (!valid[k]) continue;
        obs[k] = p->Observations() > 0;
        proj[3 * k] = p->mTrackProjX, proj[3 * k + 1] = p->mTrackProjY, proj[3 * k + 2] = p->mTrackProjXR;
        lvl[k] = p->mnTrackScaleLevel, vcos[k] = p->mTrackViewCos;
        memcpy(&desc[(size_t)k * 32], p->GetDescriptor().ptr(), 32);
    }
    CurArrays C(F);
    std::vector<int32_t> match(nc);
    int32_t nmatches = 0;
    const msl_frame_geom g = geom_of(F);
    if (msl_search_by_projection_points(matcher(), &g, th, mfNNratio, nm, valid.data(), obs.data(), proj.data(), lvl.data(),
                                        vcos.data(), desc.data(), nc, C.xy.data(), C.octave.data(), C.uright.data(),
                                        F.mDescriptors.ptr(), C.occ.data(), match.data(), &nmatches) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int j = 0; j < nc; j++)
        if (match[j] >= 0) F.mvpMapPoints[j] = vpMapPoints[match[j]];
    return nmatches;
}

namespace {
// MapPoint::mfMinDistance / mfMaxDistance are protected (include/MapPoint.h:135-136) and only exposed scaled by
// 0.8f / 1.2f; PredictScale needs the raw value.  A derived struct may name them, which yields plain
// pointers-to-member of MapPoint -- no change to the reference header.
struct MapPointDistances : MapPoint {
    static float MapPoint::*minPtr() { return &MapPointDistances::mfMinDistance; }
    static float MapPoint::*maxPtr() { return &MapPointDistances::mfMaxDistance; }
};
}  // namespace

int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound,
                                   const float th, const int ORBdist) {  // :680-797
    const std::vector<MapPoint *> vpMPs = pKF->GetMapPointMatches();
    const int nk = (int)vpMPs.size(), nc = CurrentFrame.N;
    std::vector<uint8_t> valid(nk), desc((size_t)nk * 32), occ(nc);
    std::vector<float> world(3 * (size_t)nk), dist(2 * (size_t)nk), ang(nk);
    for (int i = 0; i < nk; i++) {
        MapPoint *p = vpMPs[i];
        ang[i] = pKF->mvKeysUn[i].angle;
        valid[i] = p && !p->isBad() && !sAlreadyFound.count(p);
        if (!valid[i]) continue;
        cv::Mat x = p->GetWorldPos();
        world[3 * i] = x.at<float>(0), world[3 * i + 1] = x.at<float>(1), world[3 * i + 2] = x.at<float>(2);
        dist[2 * i] = p->*MapPointDistances::minPtr(), dist[2 * i + 1] = p->*MapPointDistances::maxPtr();
        memcpy(&desc[(size_t)i * 32], p->GetDescriptor().ptr(), 32);
    }
    CurArrays C(CurrentFrame);
    for (int j = 0; j < nc; j++) occ[j] = CurrentFrame.mvpMapPoints[j] != nullptr;  // :741-742
    std::vector<int32_t> match(nc);
    int32_t nmatches = 0;
    msl_frame_geom g = geom_of(CurrentFrame);
    g.nlevels = CurrentFrame.mnScaleLevels;
    cv::Mat Tc;
    CurrentFrame.mTcw.convertTo(Tc, CV_32F);
    if (msl_search_by_projection_keyframe(matcher(), &g, Tc.ptr<float>(), th, ORBdist, mbCheckOrientation,
                                          CurrentFrame.mfLogScaleFactor, nk, valid.data(), world.data(), desc.data(),
                                          dist.data(), ang.data(), nc, C.xy.data(), C.octave.data(), C.angle.data(),
                                          CurrentFrame.mDescriptors.ptr(), occ.data(), match.data(), &nmatches) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int j = 0; j < nc; j++) {
        if (match[j] >= 0) CurrentFrame.mvpMapPoints[j] = vpMPs[match[j]];
        else if (match[j] == -3) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint *>(NULL);
    }
    return nmatches;
}

}  // namespace ORB_SLAM2
