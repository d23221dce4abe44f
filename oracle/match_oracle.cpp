// match_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, see msl_oracle.h) restating
// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:835-849), the three Frame-side
// ORBmatcher::SearchByProjection overloads (:40-117, :548-678, :680-797), ComputeThreeMaxima (:799-830) and the
// Frame grid they search through (src/Frame.cc:155-168, 332-381, 418-427).
// The Frame/MapPoint object graph is flattened into arrays (see msl_oracle.h); cv::Mat products are
// evaluated as OpenCV's gemm does for CV_32F (double accumulation, one rounding) -- "parity unpinned".
#include "msl_oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

const int TH_HIGH = 100;      // src/ORBmatcher.cc:33
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

// cv::Mat (3x3 float) * (3x1 float) + (3x1 float) as one gemm: double accumulation, single rounding
inline float gemm_row(const float *R, const float *x, float t) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)R[k] * (double)x[k];
    return (float)(s + (double)t);
}

}  // namespace

extern "C" {

int orc_descriptor_distance(const uint8_t *a, const uint8_t *b) { return descriptor_distance(a, b); }

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
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)(-Rcw[k * 3 + r]) * (double)tcw[k];
        twc[r] = (float)s;
    }
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
    float Ow[3];  // :686  Ow = -Rcw.t() * tcw (one gemm)
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)(-Rcw[k * 3 + r]) * (double)tcw[k];
        Ow[r] = (float)s;
    }
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
        double s2 = 0;
        for (int k = 0; k < 3; k++) {
            const float po = x3Dw[k] - Ow[k];
            s2 += (double)po * (double)po;
        }
        const float dist3D = (float)std::sqrt(s2);
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

}  // extern "C"
