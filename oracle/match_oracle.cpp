// match_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, see msl_oracle.h) restating
// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:835-849), the three Frame-side
// ORBmatcher::SearchByProjection overloads (:40-117, :548-678, :680-797), SearchByBoW (:146-255),
// SearchForTriangulation (:257-406) with CheckDistEpipolarLine (:127-144), the search part of Fuse (:408-546),
// ComputeThreeMaxima (:799-830) and the Frame / KeyFrame grid they search through (src/Frame.cc:155-168, 332-381,
// 418-427; src/KeyFrame.cc:469-504).
// The Frame/MapPoint object graph is flattened into arrays (see msl_oracle.h).  cv::Mat products are evaluated as
// cv::gemm does for 3x3 / 3x1 CV_32F operands -- PINNED bit-exactly against cv2.gemm 4.13.0 in
// tests/test_oracle_primitives.py::test_cv_gemm_semantics:
//   * flags == 0 (R * x + t, -Rwc * t): the small-matrix path, FLOAT products and sums a0*b0 + a1*b1 + a2*b2, then
//     (float)(t0 * alpha + c * beta) in double;
//   * a transposed operand (-Rcw.t() * tcw): the general path, double accumulation, one rounding;
//   * cv::norm of a 3-vector: double accumulation of squares, sqrt in double.
#include "msl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

const int TH_HIGH = 100;      // src/ORBmatcher.cc:33
const int TH_LOW = 50;        // :34
const int HISTO_LENGTH = 30;  // :35
const int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:53-54

int descriptor_distance(const uint8_t *a, const uint8_t *b) {
    const int32_t *pa = (const int32_t *)a, *pb = (const int32_t *)b;
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
};

// Frame::PosInGrid :418-427
bool pos_in_grid(const orc_frame_geom *g, float x, float y, int &posX, int &posY) {
    posX = (int)std::round((x - g->mnMinX) * g->gridWInv);
    posY = (int)std::round((y - g->mnMinY) * g->gridHInv);
    if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) return false;
    return true;
}

// Frame::AssignFeaturesToGrid :155-168
void assign_grid(const orc_frame_geom *g, const float *kp_xy, int n, Grid &G) {
    for (int i = 0; i < n; i++) {
        int gx, gy;
        if (pos_in_grid(g, kp_xy[2 * i], kp_xy[2 * i + 1], gx, gy)) G.cell[gx][gy].push_back(i);
    }
}

// Frame::GetFeaturesInArea :332-381
void features_in_area(const orc_frame_geom *g, const Grid &G, const float *kp_xy, const int32_t *kp_octave, float x,
                      float y, float r, int minLevel, int maxLevel, std::vector<int> &vIndices) {
    vIndices.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - g->mnMinX - r) * g->gridWInv));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min((int)GRID_COLS - 1, (int)std::ceil((x - g->mnMinX + r) * g->gridWInv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - g->mnMinY - r) * g->gridHInv));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min((int)GRID_ROWS - 1, (int)std::ceil((y - g->mnMinY + r) * g->gridHInv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<int> &vCell = G.cell[ix][iy];
            for (size_t j = 0; j < vCell.size(); j++) {
                const int k = vCell[j];
                if (bCheckLevels) {
                    if (kp_octave[k] < minLevel) continue;
                    if (maxLevel >= 0)
                        if (kp_octave[k] > maxLevel) continue;
                }
                const float distx = kp_xy[2 * k] - x, disty = kp_xy[2 * k + 1] - y;
                if (std::fabs(distx) < r && std::fabs(disty) < r) vIndices.push_back(k);
            }
        }
}

// ORBmatcher::ComputeThreeMaxima :799-830
void three_maxima(const std::vector<int> *histo, int L, int &ind1, int &ind2, int &ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) {
            max3 = max2, max2 = max1, max1 = s;
            ind3 = ind2, ind2 = ind1, ind1 = i;
        } else if (s > max2) {
            max3 = max2, max2 = s;
            ind3 = ind2, ind2 = i;
        } else if (s > max3) {
            max3 = s, ind3 = i;
        }
    }
    if (max2 < 0.1f * (float)max1) {
        ind2 = -1, ind3 = -1;
    } else if (max3 < 0.1f * (float)max1) {
        ind3 = -1;
    }
}

// cv::Mat (3x3 float) * (3x1 float) + (3x1 float) as ONE cv::gemm with flags == 0: OpenCV's small-matrix path
// (modules/core/src/matmul: len == 3) forms the row sum in float, t0 = a0*b0 + a1*b1 + a2*b2, and stores
// (float)(t0*alpha + c*beta) evaluated in double.
inline float gemm_row(const float *R, const float *x, float t) {
    const float t0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
    return (float)((double)t0 * 1.0 + (double)t * 1.0);
}
// row r of -R.t() * t (GEMM_1_T, alpha = -1): the general path, double accumulation, one rounding
inline float gemm_neg_rt_row(const float *R, int r, const float *t) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)R[k * 3 + r] * (double)t[k];
    return (float)(-1.0 * s);
}
// row r of -Rwc * t with Rwc = Rcw.t() materialised first (KeyFrame::SetPose, src/KeyFrame.cc:79-80): flags == 0, alpha = -1
inline float gemm_neg_rwc_row(const float *Rcw, int r, const float *t) {
    const float t0 = Rcw[0 * 3 + r] * t[0] + Rcw[1 * 3 + r] * t[1] + Rcw[2 * 3 + r] * t[2];
    return (float)((double)t0 * -1.0);
}
// cv::norm(v) of a 3x1 CV_32F
inline double norm3(const float *v) {
    double s2 = 0;
    for (int k = 0; k < 3; k++) s2 += (double)v[k] * (double)v[k];
    return std::sqrt(s2);
}

}  // namespace

extern "C" {

int orc_descriptor_distance(const uint8_t *a, const uint8_t *b) { return descriptor_distance(a, b); }

// the cv::Mat arithmetic conventions above, exported so that tests can pin them against cv2 (R row-major 3x3)
void orc_cv_rx_plus_t(const float R[9], const float x[3], const float t[3], float out[3]) {
    for (int r = 0; r < 3; r++) out[r] = gemm_row(R + 3 * r, x, t[r]);
}
void orc_cv_neg_rt_times_t(const float R[9], const float t[3], float out[3]) {
    for (int r = 0; r < 3; r++) out[r] = gemm_neg_rt_row(R, r, t);
}
void orc_cv_neg_rwc_times_t(const float Rcw[9], const float t[3], float out[3]) {
    for (int r = 0; r < 3; r++) out[r] = gemm_neg_rwc_row(Rcw, r, t);
}
double orc_cv_norm3(const float v[3]) { return norm3(v); }

// brute-force best / second-best over all train descriptors (ties: lowest index); the CPU counterpart of the
// all-pairs Hamming search of BASELINE.json config 2
void orc_hamming_best2(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *best_idx, int32_t *best_dist,
                       int32_t *second_dist) {
    for (int i = 0; i < nq; i++) {
        int bd = 257, bi = -1, sd = 257;
        for (int j = 0; j < nt; j++) {
            const int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < bd) sd = bd, bd = d, bi = j;
            else if (d < sd) sd = d;
        }
        best_idx[i] = bi, best_dist[i] = bd > 256 ? 256 : bd, second_dist[i] = sd > 256 ? 256 : sd;
    }
}

int orc_features_in_area(const orc_frame_geom *g, const float *kp_xy, const int32_t *kp_octave, int n, float x, float y,
                         float r, int minLevel, int maxLevel, int32_t *out, int cap) {
    Grid G;
    assign_grid(g, kp_xy, n, G);
    std::vector<int> v;
    features_in_area(g, G, kp_xy, kp_octave, x, y, r, minLevel, maxLevel, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

int orc_search_by_projection_frame(const orc_frame_geom *g, const float Tcw_cur[16], const float Tcw_last[16], float th,
                                   int check_orientation, int n_last, const uint8_t *last_has_mp,
                                   const uint8_t *last_outlier, const uint8_t *last_mp_obs, const float *last_mp_world,
                                   const uint8_t *last_mp_desc, const int32_t *last_octave, const float *last_angle,
                                   int n_cur, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle,
                                   const float *cur_uright, const uint8_t *cur_desc, const uint8_t *cur_occupied,
                                   int32_t *cur_match) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    Grid G;
    assign_grid(g, cur_xy, n_cur, G);
    std::vector<uint8_t> blocked(cur_occupied, cur_occupied + n_cur);  // mvpMapPoints[i2] && Observations()>0
    for (int j = 0; j < n_cur; j++) cur_match[j] = cur_occupied[j] ? -2 : -1;
    // :554-568  twc = -Rcw^T tcw;  tlc = Rlw twc + tlw
    float Rcw[9], tcw[3], Rlw[9], tlw[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rcw[r * 3 + c] = Tcw_cur[r * 4 + c], Rlw[r * 3 + c] = Tcw_last[r * 4 + c];
        tcw[r] = Tcw_cur[r * 4 + 3], tlw[r] = Tcw_last[r * 4 + 3];
    }
    float twc[3];
    for (int r = 0; r < 3; r++) twc[r] = gemm_neg_rt_row(Rcw, r, tcw);
    const float tlc2 = gemm_row(Rlw + 6, twc, tlw[2]);
    const bool bForward = tlc2 > g->mb;
    const bool bBackward = -tlc2 > g->mb;
    std::vector<int> vIndices2;
    for (int i = 0; i < n_last; i++) {
        if (!last_has_mp[i]) continue;
        if (last_outlier[i]) continue;
        const float *x3Dw = last_mp_world + 3 * i;
        const float xc = gemm_row(Rcw, x3Dw, tcw[0]);
        const float yc = gemm_row(Rcw + 3, x3Dw, tcw[1]);
        const float invzc = (float)(1.0 / gemm_row(Rcw + 6, x3Dw, tcw[2]));
        if (invzc < 0) continue;
        float u = g->fx * xc * invzc + g->cx;
        float v = g->fy * yc * invzc + g->cy;
        if (u < g->mnMinX || u > g->mnMaxX) continue;
        if (v < g->mnMinY || v > g->mnMaxY) continue;
        int nLastOctave = last_octave[i];
        float radius = th * g->scaleFactors[nLastOctave];
        if (bForward)
            features_in_area(g, G, cur_xy, cur_octave, u, v, radius, nLastOctave, -1, vIndices2);
        else if (bBackward)
            features_in_area(g, G, cur_xy, cur_octave, u, v, radius, 0, nLastOctave, vIndices2);
        else
            features_in_area(g, G, cur_xy, cur_octave, u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t *dMP = last_mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); k++) {
            const int i2 = vIndices2[k];
            if (blocked[i2]) continue;
            if (cur_uright[i2] > 0) {
                const float ur = u - g->mbf * invzc;
                const float er = std::fabs(ur - cur_uright[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(dMP, cur_desc + 32 * (size_t)i2);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx2 = i2;
            }
        }
        if (bestDist <= TH_HIGH) {
            cur_match[bestIdx2] = i;
            blocked[bestIdx2] = last_mp_obs[i];
            nmatches++;
            if (check_orientation) {
                float rot = last_angle[i] - cur_angle[bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    cur_match[rotHist[i][j]] = -3;  // = static_cast<MapPoint*>(NULL)
                    nmatches--;
                }
    }
    return nmatches;
}

int orc_search_by_projection_points(const orc_frame_geom *g, float th, float nnratio, int n_mp, const uint8_t *mp_valid,
                                    const uint8_t *mp_obs, const float *mp_proj_xyr, const int32_t *mp_level,
                                    const float *mp_viewcos, const uint8_t *mp_desc, int n_cur, const float *cur_xy,
                                    const int32_t *cur_octave, const float *cur_uright, const uint8_t *cur_desc,
                                    const uint8_t *cur_occupied, int32_t *cur_match) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    Grid G;
    assign_grid(g, cur_xy, n_cur, G);
    std::vector<uint8_t> blocked(cur_occupied, cur_occupied + n_cur);
    for (int j = 0; j < n_cur; j++) cur_match[j] = cur_occupied[j] ? -2 : -1;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < n_mp; iMP++) {
        if (!mp_valid[iMP]) continue;  // !mbTrackInView || isBad()
        const int nPredictedLevel = mp_level[iMP];
        float r = (mp_viewcos[iMP] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos :119-124
        if (bFactor) r *= th;
        features_in_area(g, G, cur_xy, cur_octave, mp_proj_xyr[3 * iMP], mp_proj_xyr[3 * iMP + 1],
                         r * g->scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t *MPdescriptor = mp_desc + 32 * (size_t)iMP;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const int idx = vIndices[k];
            if (blocked[idx]) continue;
            if (cur_uright[idx] > 0) {
                const float er = std::fabs(mp_proj_xyr[3 * iMP + 2] - cur_uright[idx]);
                if (er > r * g->scaleFactors[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(MPdescriptor, cur_desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestLevel2 = bestLevel;
                bestLevel = cur_octave[idx];
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = cur_octave[idx];
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            cur_match[bestIdx] = iMP;
            blocked[bestIdx] = mp_obs[iMP];
            nmatches++;
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
// src/ORBmatcher.cc:680-797 (relocalisation) with MapPoint::PredictScale (src/MapPoint.cc:350-364) and
// MapPoint::GetMin/MaxDistanceInvariance (:324-332).  Differences from the Frame-Frame overload that are kept:
// no invzc<0 rejection, any non-NULL slot blocks (:741-742, no Observations() test), no uRight test, threshold
// ORBdist instead of TH_HIGH, level window nPredictedLevel-1..+1.
int orc_search_by_projection_keyframe(const orc_frame_geom *g, const float Tcw_cur[16], float th, int orb_dist,
                                      int check_orientation, float log_scale_factor, int n_kf, const uint8_t *kf_valid,
                                      const float *kf_mp_world, const uint8_t *kf_mp_desc, const float *kf_mp_dist,
                                      const float *kf_angle, int n_cur, const float *cur_xy, const int32_t *cur_octave,
                                      const float *cur_angle, const uint8_t *cur_desc, const uint8_t *cur_occupied,
                                      int32_t *cur_match) {
    int nmatches = 0;
    float Rcw[9], tcw[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rcw[r * 3 + c] = Tcw_cur[r * 4 + c];
        tcw[r] = Tcw_cur[r * 4 + 3];
    }
    float Ow[3];  // :686  Ow = -Rcw.t() * tcw (one gemm, transposed operand)
    for (int r = 0; r < 3; r++) Ow[r] = gemm_neg_rt_row(Rcw, r, tcw);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    Grid G;
    assign_grid(g, cur_xy, n_cur, G);
    std::vector<uint8_t> blocked(cur_occupied, cur_occupied + n_cur);  // mvpMapPoints[i2] != NULL
    for (int j = 0; j < n_cur; j++) cur_match[j] = cur_occupied[j] ? -2 : -1;
    std::vector<int> vIndices2;
    for (int i = 0; i < n_kf; i++) {
        if (!kf_valid[i]) continue;  // pMP && !pMP->isBad() && !sAlreadyFound.count(pMP)
        const float *x3Dw = kf_mp_world + 3 * i;
        const float xc = gemm_row(Rcw, x3Dw, tcw[0]);
        const float yc = gemm_row(Rcw + 3, x3Dw, tcw[1]);
        const float invzc = (float)(1.0 / gemm_row(Rcw + 6, x3Dw, tcw[2]));
        const float u = g->fx * xc * invzc + g->cx;
        const float v = g->fy * yc * invzc + g->cy;
        if (u < g->mnMinX || u > g->mnMaxX) continue;
        if (v < g->mnMinY || v > g->mnMaxY) continue;
        // :717-718  PO = x3Dw - Ow; dist3D = cv::norm(PO)  (CV_32F L2 norm: double accumulation, sqrt in double)
        const float PO[3] = {x3Dw[0] - Ow[0], x3Dw[1] - Ow[1], x3Dw[2] - Ow[2]};
        const float dist3D = (float)norm3(PO);
        const float mfMinDistance = kf_mp_dist[2 * i], mfMaxDistance = kf_mp_dist[2 * i + 1];
        const float maxDistance = 1.2f * mfMaxDistance;  // src/MapPoint.cc:329-332
        const float minDistance = 0.8f * mfMinDistance;  // :324-327
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        // MapPoint::PredictScale(dist3D, &CurrentFrame) :350-364 (std::log(float), std::ceil(float))
        const float ratio = mfMaxDistance / dist3D;
        int nPredictedLevel = (int)std::ceil(std::log(ratio) / log_scale_factor);
        if (nPredictedLevel < 0) nPredictedLevel = 0;
        else if (nPredictedLevel >= g->nlevels) nPredictedLevel = g->nlevels - 1;
        const float radius = th * g->scaleFactors[nPredictedLevel];
        features_in_area(g, G, cur_xy, cur_octave, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t *dMP = kf_mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); k++) {
            const int i2 = vIndices2[k];
            if (blocked[i2]) continue;
            const int dist = descriptor_distance(dMP, cur_desc + 32 * (size_t)i2);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx2 = i2;
            }
        }
        if (bestDist <= orb_dist && bestIdx2 >= 0) {  // bestIdx2 == -1 with ORBdist >= 256 would index [-1] in the reference
            cur_match[bestIdx2] = i;
            blocked[bestIdx2] = 1;
            nmatches++;
            if (check_orientation) {
                float rot = kf_angle[i] - cur_angle[bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    cur_match[rotHist[i][j]] = -3;
                    nmatches--;
                }
    }
    return nmatches;
}


// ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches), src/ORBmatcher.cc:146-255.
// The two DBoW2::FeatureVector maps (node id -> feature indices, ascending node id) arrive in CSR form; the merge
// loop with lower_bound (:164-233) visits exactly the node ids present in both, in ascending order.
int orc_search_by_bow(float nnratio, int check_orientation, int n_nodes_kf, const uint32_t *kf_node_id,
                      const int32_t *kf_node_off, const int32_t *kf_node_feat, int n_nodes_f, const uint32_t *f_node_id,
                      const int32_t *f_node_off, const int32_t *f_node_feat, int n_kf, const uint8_t *kf_valid,
                      const uint8_t *kf_desc, const float *kf_angle, int n_f, const uint8_t *f_desc, const float *f_angle,
                      int32_t *f_match) {
    (void)n_kf;
    for (int j = 0; j < n_f; j++) f_match[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < n_nodes_kf && b < n_nodes_f) {
        if (kf_node_id[a] == f_node_id[b]) {
            for (int iKF = kf_node_off[a]; iKF < kf_node_off[a + 1]; iKF++) {
                const int realIdxKF = kf_node_feat[iKF];
                if (!kf_valid[realIdxKF]) continue;  // !pMP || pMP->isBad()
                const uint8_t *dKF = kf_desc + 32 * (size_t)realIdxKF;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = f_node_off[b]; iF < f_node_off[b + 1]; iF++) {
                    const int realIdxF = f_node_feat[iF];
                    if (f_match[realIdxF] >= 0) continue;  // vpMapPointMatches[realIdxF] != NULL
                    const int dist = descriptor_distance(dKF, f_desc + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) {
                        bestDist2 = bestDist1;
                        bestDist1 = dist;
                        bestIdxF = realIdxF;
                    } else if (dist < bestDist2) {
                        bestDist2 = dist;
                    }
                }
                if (bestDist1 <= TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        f_match[bestIdxF] = realIdxKF;
                        if (check_orientation) {
                            float rot = kf_angle[realIdxKF] - f_angle[bestIdxF];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)std::round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(bestIdxF);
                        }
                        nmatches++;
                    }
                }
            }
            a++, b++;
        } else if (kf_node_id[a] < f_node_id[b]) {
            while (a < n_nodes_kf && kf_node_id[a] < f_node_id[b]) a++;  // lower_bound
        } else {
            while (b < n_nodes_f && f_node_id[b] < kf_node_id[a]) b++;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    f_match[rotHist[i][j]] = -3;  // = static_cast<MapPoint*>(NULL)
                    nmatches--;
                }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vMatchedPairs, bOnlyStereo),
// src/ORBmatcher.cc:257-406, with CheckDistEpipolarLine :127-144.  vbMatched2 is never set in the reference (:276, :322),
// so a KeyFrame-2 feature can be the match of several KeyFrame-1 features; `dist > bestDist` (:336) lets a later
// candidate of equal distance replace an earlier one.  matches12[idx1] = idx2, -1 none, -3 removed by the rotation check.
int orc_search_for_triangulation(const float F12[9], const float Cw1[3], const float Tcw2[16], const float K2[4],
                                 int only_stereo, int check_orientation, int nlevels, const float *scale_factors2,
                                 const float *level_sigma2_2, int n_nodes1, const uint32_t *node_id1,
                                 const int32_t *node_off1, const int32_t *node_feat1, int n_nodes2,
                                 const uint32_t *node_id2, const int32_t *node_off2, const int32_t *node_feat2, int n1,
                                 const uint8_t *has_mp1, const float *uright1, const float *xy1, const float *angle1,
                                 const uint8_t *desc1, int n2, const uint8_t *has_mp2, const float *uright2,
                                 const float *xy2, const int32_t *octave2, const float *angle2, const uint8_t *desc2,
                                 int32_t *matches12) {
    (void)nlevels, (void)n2;
    // :263-270 epipole in the second image: C2 = R2w * Cw + t2w (one gemm)
    float R2w[9], t2w[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) R2w[r * 3 + c] = Tcw2[r * 4 + c];
        t2w[r] = Tcw2[r * 4 + 3];
    }
    const float C2[3] = {gemm_row(R2w, Cw1, t2w[0]), gemm_row(R2w + 3, Cw1, t2w[1]), gemm_row(R2w + 6, Cw1, t2w[2])};
    const float invz = 1.0f / C2[2];
    const float ex = K2[0] * C2[0] * invz + K2[2];
    const float ey = K2[1] * C2[1] * invz + K2[3];
    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < n_nodes1 && b < n_nodes2) {
        if (node_id1[a] == node_id2[b]) {
            for (int i1 = node_off1[a]; i1 < node_off1[a + 1]; i1++) {
                const int idx1 = node_feat1[i1];
                if (has_mp1[idx1]) continue;
                const bool bStereo1 = uright1[idx1] >= 0;
                if (only_stereo)
                    if (!bStereo1) continue;
                const float kp1x = xy1[2 * idx1], kp1y = xy1[2 * idx1 + 1];
                const uint8_t *d1 = desc1 + 32 * (size_t)idx1;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = node_off2[b]; i2 < node_off2[b + 1]; i2++) {
                    const int idx2 = node_feat2[i2];
                    if (has_mp2[idx2]) continue;  // vbMatched2[idx2] is always false
                    const bool bStereo2 = uright2[idx2] >= 0;
                    if (only_stereo)
                        if (!bStereo2) continue;
                    const int dist = descriptor_distance(d1, desc2 + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const float kp2x = xy2[2 * idx2], kp2y = xy2[2 * idx2 + 1];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2x, distey = ey - kp2y;
                        if (distex * distex + distey * distey < 100 * scale_factors2[octave2[idx2]]) continue;
                    }
                    // CheckDistEpipolarLine :127-144
                    const float la = kp1x * F12[0] + kp1y * F12[3] + F12[6];
                    const float lb = kp1x * F12[1] + kp1y * F12[4] + F12[7];
                    const float lc = kp1x * F12[2] + kp1y * F12[5] + F12[8];
                    const float num = la * kp2x + lb * kp2y + lc;
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * level_sigma2_2[octave2[idx2]]) {
                        bestIdx2 = idx2;
                        bestDist = dist;
                    }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_orientation) {
                        float rot = angle1[idx1] - angle2[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            a++, b++;
        } else if (node_id1[a] < node_id2[b]) {
            while (a < n_nodes1 && node_id1[a] < node_id2[b]) a++;
        } else {
            while (b < n_nodes2 && node_id2[b] < node_id1[a]) b++;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    matches12[rotHist[i][j]] = -3;  // the reference writes -1 (:391)
                    nmatches--;
                }
    }
    return nmatches;
}

// The search part of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th), src/ORBmatcher.cc:408-519:
// per map point the best keypoint of the KeyFrame (bestIdx, bestDist).  What follows in the reference (:522-541,
// Replace / AddObservation / AddMapPoint on the pointer graph, whenever bestDist <= TH_LOW) stays with the caller; it
// does not feed back into the search of later map points.  KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:469-504) is the
// Frame grid query without a level filter; KeyFrame::IsInImage :540-542; MapPoint::PredictScale(dist, KeyFrame*)
// src/MapPoint.cc:334-348.  mp_dist = (mfMinDistance, mfMaxDistance).  Returns the number of map points with
// bestDist <= TH_LOW (= nFused of the reference when no map point is listed twice).
int orc_fuse_search(const orc_frame_geom *g, const float Tcw[16], float th, float log_scale_factor,
                    const float *inv_level_sigma2, int n_mp, const uint8_t *mp_valid, const float *mp_world,
                    const float *mp_normal, const float *mp_dist, const uint8_t *mp_desc, int n_kf, const float *kf_xy,
                    const int32_t *kf_octave, const float *kf_uright, const uint8_t *kf_desc, int32_t *best_idx,
                    int32_t *best_dist) {
    float Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rcw[r * 3 + c] = Tcw[r * 4 + c];
        tcw[r] = Tcw[r * 4 + 3];
    }
    for (int r = 0; r < 3; r++) Ow[r] = gemm_neg_rwc_row(Rcw, r, tcw);  // KeyFrame::SetPose: Rwc = Rcw.t(); Ow = -Rwc * tcw
    Grid G;
    assign_grid(g, kf_xy, n_kf, G);
    std::vector<int> vIndices;
    int nFused = 0;
    for (int i = 0; i < n_mp; i++) {
        best_idx[i] = -1, best_dist[i] = 256;
        if (!mp_valid[i]) continue;  // !pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)
        const float *p3Dw = mp_world + 3 * i;
        const float pc0 = gemm_row(Rcw, p3Dw, tcw[0]), pc1 = gemm_row(Rcw + 3, p3Dw, tcw[1]), pc2 = gemm_row(Rcw + 6, p3Dw, tcw[2]);
        if (pc2 < 0.0f) continue;
        const float invz = 1 / pc2;
        const float x = pc0 * invz, y = pc1 * invz;
        const float u = g->fx * x + g->cx, v = g->fy * y + g->cy;
        if (!(u >= g->mnMinX && u < g->mnMaxX && v >= g->mnMinY && v < g->mnMaxY)) continue;
        const float ur = u - g->mbf * invz;
        const float maxDistance = 1.2f * mp_dist[2 * i + 1], minDistance = 0.8f * mp_dist[2 * i];
        const float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        const float dist3D = (float)norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        double dot = 0;  // cv::Mat::dot on 3 floats: double products, double accumulation
        for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)mp_normal[3 * i + k];
        if (dot < 0.5 * dist3D) continue;
        const float ratio = mp_dist[2 * i + 1] / dist3D;
        int nPredictedLevel = (int)std::ceil(std::log(ratio) / log_scale_factor);
        if (nPredictedLevel < 0) nPredictedLevel = 0;
        else if (nPredictedLevel >= g->nlevels) nPredictedLevel = g->nlevels - 1;
        const float radius = th * g->scaleFactors[nPredictedLevel];
        features_in_area(g, G, kf_xy, kf_octave, u, v, radius, -1, -1, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t *dMP = mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const int idx = vIndices[k];
            const int kpLevel = kf_octave[idx];
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const float kpx = kf_xy[2 * idx], kpy = kf_xy[2 * idx + 1];
            if (kf_uright[idx] >= 0) {
                const float ex = u - kpx, ey = v - kpy, er = ur - kf_uright[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = u - kpx, ey = v - kpy;
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
            }
            const int dist = descriptor_distance(dMP, kf_desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx = idx;
            }
        }
        best_idx[i] = bestIdx, best_dist[i] = bestDist;
        if (bestDist <= TH_LOW) nFused++;
    }
    return nFused;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263) for a batch of map points: point k owns the
// descriptors desc[off[k] .. off[k+1]) (the observations in non-bad KeyFrames, in std::map order).  best_idx[k] = BestIdx
// (index inside the point's list; -1 for a point without descriptors, which the reference leaves untouched),
// best_median[k] = BestMedian.
void orc_distinctive_descriptors(int n_points, const int32_t *off, const uint8_t *desc, int32_t *best_idx,
                                 int32_t *best_median) {
    for (int p = 0; p < n_points; p++) {
        const size_t N = (size_t)(off[p + 1] - off[p]);
        best_idx[p] = -1, best_median[p] = 0x7fffffff;
        if (N == 0) continue;
        const uint8_t *D = desc + 32 * (size_t)off[p];
        std::vector<float> Distances(N * N);  // float Distances[N][N] in the reference
        for (size_t i = 0; i < N; i++) {
            Distances[i * N + i] = 0;
            for (size_t j = i + 1; j < N; j++) {
                const int distij = descriptor_distance(D + 32 * i, D + 32 * j);
                Distances[i * N + j] = (float)distij;
                Distances[j * N + i] = (float)distij;
            }
        }
        int BestMedian = 0x7fffffff, BestIdx = 0;
        for (size_t i = 0; i < N; i++) {
            std::vector<int> vDists(Distances.begin() + i * N, Distances.begin() + (i + 1) * N);
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) {
                BestMedian = median;
                BestIdx = (int)i;
            }
        }
        best_idx[p] = BestIdx, best_median[p] = BestMedian;
    }
}

}  // extern "C"
