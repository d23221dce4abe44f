// ORBmatcher_msl.cc -- the tracking-time searches of ORB_SLAM2::ORBmatcher on the B200 front-end.
// Build together with the reference's src/ORBmatcher.cc compiled with -DMSL_FRONTEND, where the seven
// definitions below (three SearchByProjection overloads, DescriptorDistance, SearchByBoW, SearchForTriangulation, Fuse)
// are wrapped in `#ifndef MSL_FRONTEND` (INTEGRATION.md shows the patch); the remaining methods (CheckDistEpipolarLine,
// RadiusByViewingCos, ComputeThreeMaxima) stay as they are.  Callers are unchanged (src/Tracking.cc:859,956,963,1158,1253,
// 1262,1693,1942,2006,2019; src/LocalMapping.cc:351,549,569).
#include <cstring>
#include <set>
#include <stdexcept>
#include <utility>
#include <vector>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#include "msl_frontend.h"

namespace ORB_SLAM2 {

namespace {
// One handle (stream + scratch arena) per calling thread: the Tracking thread runs the SearchByProjection overloads and
// SearchByBoW while the LocalMapping thread runs SearchForTriangulation and Fuse (src/LocalMapping.cc:351,549,569), and a
// handle is not thread-safe (include/msl_frontend.h).  The handles live as long as their threads.
msl_matcher *matcher() {
    struct Holder {
        msl_matcher *m = nullptr;
        ~Holder() {
            if (m) msl_matcher_destroy(m);
        }
    };
    static thread_local Holder h;
    if (!h.m && msl_matcher_create(4096, 4096, 1, 0, &h.m) != MSL_OK) throw std::runtime_error(msl_last_error());
    return h.m;
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

namespace {
// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>) in the CSR form of the C ABI
struct FeatVecCsr {
    std::vector<uint32_t> id;
    std::vector<int32_t> off, feat;
    explicit FeatVecCsr(const DBoW2::FeatureVector &fv) {
        off.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            id.push_back(it->first);
            for (size_t k = 0; k < it->second.size(); k++) feat.push_back((int32_t)it->second[k]);
            off.push_back((int32_t)feat.size());
        }
    }
};
msl_frame_geom geom_of(const KeyFrame *pKF) {
    msl_frame_geom g = {};
    g.fx = pKF->fx, g.fy = pKF->fy, g.cx = pKF->cx, g.cy = pKF->cy;
    g.mnMinX = pKF->mnMinX, g.mnMinY = pKF->mnMinY, g.mnMaxX = pKF->mnMaxX, g.mnMaxY = pKF->mnMaxY;
    g.gridWInv = pKF->mfGridElementWidthInv, g.gridHInv = pKF->mfGridElementHeightInv;
    g.mb = pKF->mb, g.mbf = pKF->mbf;
    g.nlevels = pKF->mnScaleLevels;
    for (int i = 0; i < g.nlevels && i < 16; i++) g.scaleFactors[i] = pKF->mvScaleFactors[i];
    return g;
}
}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches) {  // :146-255
    const std::vector<MapPoint *> vpMapPointsKF = pKF->GetMapPointMatches();
    const int nk = (int)vpMapPointsKF.size(), nf = F.N;
    vpMapPointMatches = std::vector<MapPoint *>(nf, static_cast<MapPoint *>(NULL));
    std::vector<uint8_t> valid(nk);
    std::vector<float> kang(nk), fang(nf);
    for (int i = 0; i < nk; i++) {
        MapPoint *p = vpMapPointsKF[i];
        valid[i] = p && !p->isBad();
        kang[i] = pKF->mvKeysUn[i].angle;
    }
    for (int j = 0; j < nf; j++) fang[j] = F.mvKeys[j].angle;
    const FeatVecCsr a(pKF->mFeatVec), b(F.mFeatVec);
    std::vector<int32_t> match(nf);
    int32_t nmatches = 0;
    if (msl_search_by_bow(matcher(), mfNNratio, mbCheckOrientation, (int)a.id.size(), a.id.data(), a.off.data(), a.feat.data(),
                          (int)b.id.size(), b.id.data(), b.off.data(), b.feat.data(), nk, valid.data(), pKF->mDescriptors.ptr(),
                          kang.data(), nf, F.mDescriptors.ptr(), fang.data(), match.data(), &nmatches) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int j = 0; j < nf; j++)
        if (match[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match[j]];
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo) {  // :257-406
    const int n1 = pKF1->N, n2 = pKF2->N;
    std::vector<uint8_t> has1(n1), has2(n2);
    std::vector<float> xy1(2 * (size_t)n1), ang1(n1), xy2(2 * (size_t)n2), ang2(n2);
    std::vector<int32_t> oct2(n2);
    for (int i = 0; i < n1; i++) {
        has1[i] = pKF1->GetMapPoint(i) != nullptr;
        xy1[2 * i] = pKF1->mvKeysUn[i].pt.x, xy1[2 * i + 1] = pKF1->mvKeysUn[i].pt.y, ang1[i] = pKF1->mvKeysUn[i].angle;
    }
    for (int j = 0; j < n2; j++) {
        has2[j] = pKF2->GetMapPoint(j) != nullptr;
        xy2[2 * j] = pKF2->mvKeysUn[j].pt.x, xy2[2 * j + 1] = pKF2->mvKeysUn[j].pt.y, ang2[j] = pKF2->mvKeysUn[j].angle;
        oct2[j] = pKF2->mvKeysUn[j].octave;
    }
    cv::Mat Cw = pKF1->GetCameraCenter(), T2 = pKF2->GetPose(), F;
    F12.convertTo(F, CV_32F);
    F = F.clone();  // dense row-major 3x3
    cv::Mat T2f;
    T2.convertTo(T2f, CV_32F);
    T2f = T2f.clone();
    const float Cw1[3] = {Cw.at<float>(0), Cw.at<float>(1), Cw.at<float>(2)};
    const float K2[4] = {pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy};
    const FeatVecCsr a(pKF1->mFeatVec), b(pKF2->mFeatVec);
    std::vector<int32_t> m12(n1);
    int32_t nmatches = 0;
    if (msl_search_for_triangulation(matcher(), F.ptr<float>(), Cw1, T2f.ptr<float>(), K2, bOnlyStereo, mbCheckOrientation,
                                     pKF2->mnScaleLevels, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                                     (int)a.id.size(), a.id.data(), a.off.data(), a.feat.data(), (int)b.id.size(), b.id.data(),
                                     b.off.data(), b.feat.data(), n1, has1.data(), pKF1->mvuRight.data(), xy1.data(), ang1.data(),
                                     pKF1->mDescriptors.ptr(), n2, has2.data(), pKF2->mvuRight.data(), xy2.data(), oct2.data(),
                                     ang2.data(), pKF2->mDescriptors.ptr(), m12.data(), &nmatches) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < n1; i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
    return nmatches;
}

int ORBmatcher::Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th) {  // :408-546
    const int nm = (int)vpMapPoints.size(), nk = pKF->N;
    std::vector<uint8_t> valid(nm), desc((size_t)nm * 32);
    std::vector<float> world(3 * (size_t)nm), normal(3 * (size_t)nm), dist(2 * (size_t)nm), xy(2 * (size_t)nk);
    std::vector<int32_t> oct(nk);
    for (int i = 0; i < nm; i++) {
        MapPoint *p = vpMapPoints[i];
        valid[i] = p && !p->isBad() && !p->IsInKeyFrame(pKF);
        if (!valid[i]) continue;
        cv::Mat x = p->GetWorldPos(), nrm = p->GetNormal();
        for (int k = 0; k < 3; k++) world[3 * i + k] = x.at<float>(k), normal[3 * i + k] = nrm.at<float>(k);
        dist[2 * i] = p->*MapPointDistances::minPtr(), dist[2 * i + 1] = p->*MapPointDistances::maxPtr();
        memcpy(&desc[(size_t)i * 32], p->GetDescriptor().ptr(), 32);
    }
    for (int j = 0; j < nk; j++) xy[2 * j] = pKF->mvKeysUn[j].pt.x, xy[2 * j + 1] = pKF->mvKeysUn[j].pt.y, oct[j] = pKF->mvKeysUn[j].octave;
    cv::Mat T;
    pKF->GetPose().convertTo(T, CV_32F);
    T = T.clone();
    const msl_frame_geom g = geom_of(pKF);
    std::vector<int32_t> bestIdx(nm), bestDist(nm);
    int32_t n = 0;
    if (msl_fuse_search(matcher(), &g, T.ptr<float>(), th, pKF->mfLogScaleFactor, pKF->mvInvLevelSigma2.data(), nm, valid.data(),
                        world.data(), normal.data(), dist.data(), desc.data(), nk, xy.data(), oct.data(), pKF->mvuRight.data(),
                        pKF->mDescriptors.ptr(), bestIdx.data(), bestDist.data(), &n) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    // :521-541 on the pointer graph, in map-point order.  A map point listed twice is in the KeyFrame after its first
    // visit (AddObservation), which the reference re-tests per iteration (:425).
    int nFused = 0;
    for (int i = 0; i < nm; i++) {
        if (!valid[i] || bestDist[i] > TH_LOW) continue;
        MapPoint *pMP = vpMapPoints[i];
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(pMP, bestIdx[i]);
        }
        nFused++;
    }
    return nFused;
}

}  // namespace ORB_SLAM2
