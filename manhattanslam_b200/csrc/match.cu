// match.cu -- B200-native descriptor matching (replaces ORBmatcher::DescriptorDistance and the window
// searches ORBmatcher::SearchByProjection of src/ORBmatcher.cc:40-117, 548-678, 680-797, 799-849 with the
// Frame grid semantics of src/Frame.cc:155-168, 332-381, 418-427 of razayunus/ManhattanSLAM; the vocabulary-node
// searches SearchByBoW :146-255 and SearchForTriangulation :257-406, and the search part of Fuse :408-519).
//
//   k_hamming_best2     all-pairs 256-bit Hamming: one warp per query, train descriptors staged through
//                       shared memory in 256-descriptor tiles, xor + __popc on 8 x u32, warp-shuffle
//                       (best, index, second) reduction; ties resolve to the lowest train index
//   k_hamming_all_pairs full nq x nt distance matrix (uint16)
//   k_search            one CTA per Frame pair: (1) the 64x48 feature grid as a sorted (cell, index) list --
//                       reproducing GetFeaturesInArea's candidate order (ix asc, iy asc, insertion asc);
//                       (2) per-query projection / window / level predicate; (3) the reference's greedy,
//                       order-dependent slot blocking (Observations()>0) solved as a fixed point:
//                       query i sees slot j blocked iff an earlier query i' < i holds it -- iterate until
//                       no assignment changes (unique solution by induction on i); (4) rotation histogram
//                       + ComputeThreeMaxima pruning.
//   k_node_search       SearchByBoW / SearchForTriangulation: the two FeatureVectors in CSR form; one warp per
//                       vocabulary node of the first (binary search for the same id in the second), queries of a node
//                       in order (SearchByBoW's "slot already matched" only couples queries of one node, because a
//                       feature belongs to exactly one node), lanes over the node's candidates, warp-shuffle
//                       reduction under the reference's tie rules; rotation histogram + ComputeThreeMaxima.
//   k_fuse_search       Fuse: the KeyFrame grid as in k_search, one thread per map point (projection, distance /
//                       viewing-angle / scale tests, chi-square gate per candidate, strict-< best).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "msl_common.cuh"

using namespace msl;

namespace {

constexpr int TH_HIGH = 100;      // src/ORBmatcher.cc:33
constexpr int TH_LOW = 50;        // :34
constexpr int HISTO_LENGTH = 30;  // :35
constexpr int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:53-54
constexpr int NCELLS = GRID_COLS * GRID_ROWS;
constexpr int MAXK = 4096;  // keypoints per frame handled by k_search

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256)
    k_hamming_best2(const uint8_t *__restrict__ q, int nqMax, int qRows, const uint8_t *__restrict__ t, int ntMax, int tRows,
                    const int32_t *__restrict__ qCounts, const int32_t *__restrict__ tCounts, int32_t *__restrict__ bestIdx,
                    int32_t *__restrict__ bestDist, int32_t *__restrict__ secondDist) {
    __shared__ uint4 tile[256 * 2];
    const int b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int qi = blockIdx.x * 8 + wid;
    // qRows/tRows = descriptor rows reserved per batch entry; counts (optional) = rows actually filled
    const int nq = qCounts ? min(qCounts[b], nqMax) : nqMax, nt = tCounts ? min(tCounts[b], ntMax) : ntMax;
    if (blockIdx.x * 8 >= nq) return;
    const uint4 *Q = (const uint4 *)(q + (size_t)b * qRows * 32);
    const uint4 *T = (const uint4 *)(t + (size_t)b * tRows * 32);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (qi < nq) q0 = __ldg(Q + 2 * qi), q1 = __ldg(Q + 2 * qi + 1);
    int bd = 257, bi = -1, sd = 257;
    for (int base = 0; base < nt; base += 256) {
        const int m = min(256, nt - base);
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * m; k += 256) tile[k] = __ldg(T + 2 * base + k);
        __syncthreads();
        if (qi < nq)
            for (int j = lane; j < m; j += 32) {
                const int d = hamming256(q0, q1, tile[2 * j], tile[2 * j + 1]);
                if (d < bd) {
                    sd = bd;
                    bd = d;
                    bi = base + j;
                } else if (d < sd)
                    sd = d;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int obd = __shfl_xor_sync(0xffffffffu, bd, o), obi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int osd = __shfl_xor_sync(0xffffffffu, sd, o);
        const int ns = min(max(bd, obd), min(sd, osd));
        if (obd < bd || (obd == bd && (unsigned)obi < (unsigned)bi)) bd = obd, bi = obi;
        sd = ns;
    }
    if (qi < nq && lane == 0) {
        const size_t o = (size_t)b * qRows + qi;
        bestIdx[o] = bi;
        bestDist[o] = bd > 256 ? 256 : bd;
        secondDist[o] = sd > 256 ? 256 : sd;
    }
}

__global__ void __launch_bounds__(256)
    k_hamming_all_pairs(const uint8_t *__restrict__ q, int nq, const uint8_t *__restrict__ t, int nt, uint16_t *__restrict__ dist) {
    const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= nt) return;
    const uint4 *Q = (const uint4 *)(q + ((size_t)b * nq + i) * 32);
    const uint4 *T = (const uint4 *)(t + ((size_t)b * nt + j) * 32);
    dist[((size_t)b * nq + i) * nt + j] = (uint16_t)hamming256(__ldg(Q), __ldg(Q + 1), __ldg(T), __ldg(T + 1));
}

// -------------------------------------------------------------------------------- window search
struct SearchArgs {
    msl_frame_geom g;
    int mode;                 // 0: Frame-Frame (:548-678), 1: Frame-MapPoints (:40-117), 2: Frame-KeyFrame (:680-797)
    float th, nnratio;
    int checkOri;
    float Rcw[9], tcw[3];
    float Ow[3], logScale;    // mode 2: camera centre, Frame::mfLogScaleFactor
    int distTh;               // TH_HIGH (modes 0, 1) or ORBdist (mode 2)
    int bForward, bBackward;
    int nq, nc;
    // query side
    const uint8_t *q_valid, *q_obs, *q_desc;
    const float *q_f3;        // modes 0, 2: world xyz; mode 1: (projX, projY, projXR)
    const float *q_f2;        // mode 2: (mfMinDistance, mfMaxDistance)
    const int32_t *q_level;   // mode 0: last octave; mode 1: predicted level
    const float *q_aux;       // mode 0: last angle; mode 1: view cos
    // current frame
    const float *c_xy, *c_angle, *c_uright;
    const int32_t *c_octave;
    const uint8_t *c_desc, *c_occ;
    int32_t *c_match, *nmatches;
    // scratch (global)
    int32_t *assign, *bin;
};

// cv::Mat (3x3) * (3x1) + (3x1), CV_32F, as ONE cv::gemm with flags == 0: OpenCV's small-matrix path forms the row sum in
// float, t0 = a0*b0 + a1*b1 + a2*b2, and stores (float)(t0*alpha + c*beta) evaluated in double (pinned against cv2.gemm,
// tests/test_oracle_primitives.py::test_cv_gemm_semantics).  Built with -fmad=false: no contraction.
__host__ __device__ __forceinline__ float gemm_row(const float *R, const float *x, float t) {
    const float t0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
    return (float)((double)t0 * 1.0 + (double)t * 1.0);
}

__device__ __forceinline__ void bitonic_sort_u32(uint32_t *a, int n2) {  // n2 = power of two, whole CTA
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t x = a[i], y = a[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) a[i] = y, a[l] = x;
                }
            }
            __syncthreads();
        }
}

__device__ void search_body(const SearchArgs &A) {
    __shared__ uint32_t keys[MAXK];          // (cell << 12 | index), sorted; 0xffffffff = not in grid
    __shared__ unsigned short cellStart[NCELLS + 1];
    __shared__ int blockedAt[MAXK];           // first query index holding the slot (INF = free)
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_changed, s_n, s_ind[3];
    const int tid = threadIdx.x, nt = blockDim.x;
    const msl_frame_geom &g = A.g;
    // ---- (1) grid: AssignFeaturesToGrid (src/Frame.cc:155-168) as a sorted list
    int n2 = 1;
    while (n2 < A.nc) n2 <<= 1;
    for (int i = tid; i < n2; i += nt) {
        uint32_t key = 0xffffffffu;
        if (i < A.nc) {
            const int px = (int)roundf((A.c_xy[2 * i] - g.mnMinX) * g.gridWInv);   // PosInGrid :418-427
            const int py = (int)roundf((A.c_xy[2 * i + 1] - g.mnMinY) * g.gridHInv);
            if (!(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS)) key = ((uint32_t)(px * GRID_ROWS + py) << 12) | (uint32_t)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    bitonic_sort_u32(keys, n2);
    for (int c = tid; c <= NCELLS; c += nt) {  // first position with cell >= c
        int lo = 0, hi = A.nc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((keys[mid] >> 12) < (uint32_t)c) lo = mid + 1; else hi = mid;
        }
        cellStart[c] = (unsigned short)lo;
    }
    for (int j = tid; j < A.nc; j += nt) blockedAt[j] = A.c_occ[j] ? -1 : 0x7fffffff;
    for (int i = tid; i < A.nq; i += nt) A.assign[i] = -1;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();

    // ---- (2)+(3) search rounds
    for (int round = 0; round < A.nq + 2; round++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < A.nq; i += nt) {
            int best = -1;
            if (A.q_valid[i]) {
                float u, v, radius, urRef, erLim;
                int minL, maxL;
                bool ok = true;
                if (A.mode == 0) {
                    const float *x3Dw = A.q_f3 + 3 * i;
                    const float xc = gemm_row(A.Rcw, x3Dw, A.tcw[0]);
                    const float yc = gemm_row(A.Rcw + 3, x3Dw, A.tcw[1]);
                    const float invzc = (float)(1.0 / (double)gemm_row(A.Rcw + 6, x3Dw, A.tcw[2]));
                    if (invzc < 0) ok = false;
                    u = g.fx * xc * invzc + g.cx;
                    v = g.fy * yc * invzc + g.cy;
                    if (u < g.mnMinX || u > g.mnMaxX) ok = false;
                    if (v < g.mnMinY || v > g.mnMaxY) ok = false;
                    const int oct = A.q_level[i];
                    radius = A.th * g.scaleFactors[oct];
                    if (A.bForward) minL = oct, maxL = -1;
                    else if (A.bBackward) minL = 0, maxL = oct;
                    else minL = oct - 1, maxL = oct + 1;
                    urRef = u - g.mbf * invzc;
                    erLim = radius;
                } else if (A.mode == 2) {
                    const float *x3Dw = A.q_f3 + 3 * i;
                    const float xc = gemm_row(A.Rcw, x3Dw, A.tcw[0]);
                    const float yc = gemm_row(A.Rcw + 3, x3Dw, A.tcw[1]);
                    const float invzc = (float)(1.0 / (double)gemm_row(A.Rcw + 6, x3Dw, A.tcw[2]));  // no invzc<0 test (:705)
                    u = g.fx * xc * invzc + g.cx;
                    v = g.fy * yc * invzc + g.cy;
                    if (u < g.mnMinX || u > g.mnMaxX) ok = false;
                    if (v < g.mnMinY || v > g.mnMaxY) ok = false;
                    // :717-718 cv::norm(x3Dw - Ow): float differences, double accumulation, double sqrt
                    double s2 = 0;
                    for (int k = 0; k < 3; k++) {
                        const float po = x3Dw[k] - A.Ow[k];
                        s2 += (double)po * (double)po;
                    }
                    const float dist3D = (float)sqrt(s2);
                    const float minD = A.q_f2[2 * i], maxD = A.q_f2[2 * i + 1];
                    if (dist3D < 0.8f * minD || dist3D > 1.2f * maxD) ok = false;  // src/MapPoint.cc:324-332
                    // MapPoint::PredictScale (src/MapPoint.cc:350-364): ceil(logf(ratio) / logScale).  logf is taken as the
                    // rounded fp64 logarithm (glibc's logf is correctly rounded in all but ~1e-8 of its inputs)
                    const float ratio = maxD / dist3D;
                    int lvl = 0;
                    if (ok) {
                        lvl = (int)ceilf((float)log((double)ratio) / A.logScale);
                        if (lvl < 0) lvl = 0;
                        else if (lvl >= g.nlevels) lvl = g.nlevels - 1;
                    }
                    radius = A.th * g.scaleFactors[lvl];
                    minL = lvl - 1, maxL = lvl + 1;
                    urRef = 0.f, erLim = 0.f;
                } else {
                    const int lvl = A.q_level[i];
                    float r = (A.q_aux[i] > 0.998f) ? 2.5f : 4.0f;  // RadiusByViewingCos :119-124
                    if (A.th != 1.0f) r *= A.th;
                    u = A.q_f3[3 * i], v = A.q_f3[3 * i + 1];
                    radius = r * g.scaleFactors[lvl];
                    minL = lvl - 1, maxL = lvl;
                    urRef = A.q_f3[3 * i + 2];
                    erLim = r * g.scaleFactors[lvl];
                }
                if (ok) {
                    // GetFeaturesInArea, src/Frame.cc:332-381
                    const int cx0 = max(0, (int)floorf((u - g.mnMinX - radius) * g.gridWInv));
                    const int cx1 = min(GRID_COLS - 1, (int)ceilf((u - g.mnMinX + radius) * g.gridWInv));
                    const int cy0 = max(0, (int)floorf((v - g.mnMinY - radius) * g.gridHInv));
                    const int cy1 = min(GRID_ROWS - 1, (int)ceilf((v - g.mnMinY + radius) * g.gridHInv));
                    if (cx0 < GRID_COLS && cx1 >= 0 && cy0 < GRID_ROWS && cy1 >= 0) {
                        const bool checkLevels = (minL > 0) || (maxL >= 0);
                        const uint4 *Q = (const uint4 *)(A.q_desc + 32 * (size_t)i);
                        const uint4 q0 = Q[0], q1 = Q[1];
                        int bestDist = 256, bestDist2 = 256, bestLevel = -1, bestLevel2 = -1;
                        for (int ix = cx0; ix <= cx1; ix++) {
                            const int s0 = cellStart[ix * GRID_ROWS + cy0], s1 = cellStart[ix * GRID_ROWS + cy1 + 1];
                            for (int s = s0; s < s1; s++) {
                                const int k = keys[s] & 0xfff;
                                const int oct = A.c_octave[k];
                                if (checkLevels) {
                                    if (oct < minL) continue;
                                    if (maxL >= 0 && oct > maxL) continue;
                                }
                                const float dx = A.c_xy[2 * k] - u, dy = A.c_xy[2 * k + 1] - v;
                                if (!(fabsf(dx) < radius && fabsf(dy) < radius)) continue;
                                if (blockedAt[k] < i) continue;  // slot holds a MapPoint with Observations()>0
                                const float ur = A.c_uright[k];
                                if (A.mode != 2 && ur > 0 && fabsf(urRef - ur) > erLim) continue;
                                const uint4 *T = (const uint4 *)(A.c_desc + 32 * (size_t)k);
                                const int d = hamming256(q0, q1, T[0], T[1]);
                                if (d < bestDist) {
                                    bestDist2 = bestDist, bestLevel2 = bestLevel;
                                    bestDist = d, bestLevel = oct, best = k;
                                } else if (d < bestDist2) {
                                    bestLevel2 = oct, bestDist2 = d;
                                }
                            }
                        }
                        if (!(bestDist <= A.distTh)) best = -1;
                        else if (A.mode == 1 && bestLevel == bestLevel2 && (float)bestDist > A.nnratio * (float)bestDist2) best = -1;
                    }
                }
            }
            if (best != A.assign[i]) {
                A.assign[i] = best;
                s_changed = 1;
            }
        }
        __syncthreads();
        if (!s_changed) break;
        for (int j = tid; j < A.nc; j += nt) blockedAt[j] = A.c_occ[j] ? -1 : 0x7fffffff;
        __syncthreads();
        for (int i = tid; i < A.nq; i += nt) {
            const int k = A.assign[i];
            if (k >= 0 && A.q_obs[i]) atomicMin(&blockedAt[k], i);
        }
        __syncthreads();
    }

    // ---- (4) matches, rotation histogram (:639-675) and ComputeThreeMaxima (:799-830)
    int *lastAssign = blockedAt;  // reuse: last query assigned to each slot (-1 none)
    __syncthreads();
    for (int j = tid; j < A.nc; j += nt) lastAssign[j] = -1;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < A.nq; i += nt) {
        const int k = A.assign[i];
        int bin = -1;
        if (k >= 0) {
            mine++;
            atomicMax(&lastAssign[k], i);
            if (A.mode != 1 && A.checkOri) {
                float rot = A.q_aux[i] - A.c_angle[k];
                if (rot < 0.0f) rot += 360.0f;
                bin = (int)roundf(rot * (1.0f / HISTO_LENGTH));
                if (bin == HISTO_LENGTH) bin = 0;
                atomicAdd(&hist[bin], 1);
            }
        }
        A.bin[i] = bin;
    }
    if (mine) atomicAdd(&s_n, mine);
    __syncthreads();
    if (tid == 0) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        if (A.mode != 1 && A.checkOri) {
            int max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < HISTO_LENGTH; i++) {
                const int s = hist[i];
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
            if ((float)max2 < 0.1f * (float)max1) ind2 = -1, ind3 = -1;
            else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        }
        s_ind[0] = ind1, s_ind[1] = ind2, s_ind[2] = ind3;
    }
    __syncthreads();
    for (int j = tid; j < A.nc; j += nt) A.c_match[j] = A.c_occ[j] ? -2 : lastAssign[j];
    __syncthreads();
    if (A.mode != 1 && A.checkOri) {
        int pruned = 0;
        for (int i = tid; i < A.nq; i += nt) {
            const int b = A.bin[i];
            if (b >= 0 && b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) {
                A.c_match[A.assign[i]] = -3;  // mvpMapPoints[..] = NULL
                pruned++;
            }
        }
        if (pruned) atomicSub(&s_n, pruned);
    }
    __syncthreads();
    if (tid == 0) *A.nmatches = s_n;
}

__global__ void __launch_bounds__(1024) k_search(SearchArgs A) { search_body(A); }
// one CTA per recorded call of a deferred batch (msl_matcher_batch_begin / _end)
__global__ void __launch_bounds__(1024) k_search_many(const SearchArgs *__restrict__ args) {
    __shared__ SearchArgs A;
    if (threadIdx.x < sizeof(SearchArgs) / 4) reinterpret_cast<uint32_t *>(&A)[threadIdx.x] = reinterpret_cast<const uint32_t *>(args + blockIdx.x)[threadIdx.x];
    __syncthreads();
    search_body(A);
}

// ---- SearchByProjection(CurrentFrame, LastFrame, th) for a BATCH of consecutive frame pairs straight from the device-resident
// output of the extractor and the frame glue (pair p: Last = frame p, Current = frame p + 1).
// k_track_last is the Last-frame side of Tracking::TrackWithMotionModel: Tracking::UpdateLastFrame (src/Tracking.cc:1052-1104)
// gives the RGB-D keypoints of the last frame "visual odometry" MapPoints -- sorted by depth, all with z <= mThDepth, at
// least the 100 closest -- at Frame::UnprojectStereo (src/Frame.cc:515-526), x3Dw = mRwc * x3Dc + mOw as cv::Mat
// arithmetic (gemm_row).  Such points have no observations, so they never block a slot (q_obs = 0), and a fresh current
// frame has no occupied slot.
struct TrackPose {      // per frame: Rwc (= Rcw^T), Ow, and for the pair whose CURRENT frame this is: Rcw, tcw, direction flags
    float Rwc[9], Ow[3], Rcw[9], tcw[3];
    int bForward, bBackward;
};

__global__ void __launch_bounds__(256)
    k_track_last(const msl_keypoint *__restrict__ kps, int rows, const int32_t *__restrict__ counts, const float *__restrict__ xyUn,
                 const float *__restrict__ kdepth, const TrackPose *__restrict__ poses, float cx, float cy, float invfx, float invfy,
                 float thDepth, uint8_t *__restrict__ valid, float *__restrict__ world, int32_t *__restrict__ octave,
                 float *__restrict__ angle) {
    __shared__ float sz[MAXK];
    __shared__ int s_cut;
    const int f = blockIdx.x, tid = threadIdx.x, n = min(counts[f], min(rows, MAXK));
    const size_t o = (size_t)f * rows;
    for (int i = tid; i < n; i += blockDim.x) sz[i] = kdepth[o + i];
    if (tid == 0) s_cut = 0x7fffffff;
    __syncthreads();
    // rank of every keypoint with z > 0 in std::sort order of (z, index) pairs (:1064-1073)
    for (int i = tid; i < n; i += blockDim.x) {
        const float z = sz[i];
        int rank = -1;
        if (z > 0) {
            rank = 0;
            for (int j = 0; j < n; j++) {
                const float zj = sz[j];
                rank += (zj > 0) && (zj < z || (zj == z && j < i));
            }
            // the loop over the sorted list stops after the first point with z > mThDepth once more than 100 were taken (:1102)
            if (z > thDepth && rank + 1 > 100) atomicMin(&s_cut, rank);
        }
        octave[o + i] = kps[o + i].octave;
        angle[o + i] = kps[o + i].angle;
        world[3 * (o + i)] = __int_as_float(rank);  // parked until the cut is known
    }
    __syncthreads();
    const int cut = s_cut;
    const TrackPose &T = poses[f];
    for (int i = tid; i < n; i += blockDim.x) {
        const int rank = __float_as_int(world[3 * (o + i)]);
        const bool sel = rank >= 0 && rank <= cut;
        valid[o + i] = sel ? 1 : 0;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f;
        if (sel) {
            const float z = sz[i], u = xyUn[2 * (o + i)], v = xyUn[2 * (o + i) + 1];
            const float x3[3] = {(u - cx) * z * invfx, (v - cy) * z * invfy, z};
            w0 = gemm_row(T.Rwc, x3, T.Ow[0]), w1 = gemm_row(T.Rwc + 3, x3, T.Ow[1]), w2 = gemm_row(T.Rwc + 6, x3, T.Ow[2]);
        }
        world[3 * (o + i)] = w0, world[3 * (o + i) + 1] = w1, world[3 * (o + i) + 2] = w2;
    }
}

struct SearchBatch {
    msl_frame_geom g;
    float th;
    int checkOri, rows;
    const int32_t *counts;
    const TrackPose *poses;
    const uint8_t *valid, *zeros, *desc;
    const float *world, *xyUn, *angle, *uright;
    const int32_t *octave;
    int32_t *match, *nmatches, *assign, *bin;
};

__global__ void __launch_bounds__(1024) k_search_batch(SearchBatch Bt) {
    __shared__ SearchArgs A;
    const int pr = blockIdx.x;  // Last = frame pr, Current = frame pr + 1
    if (threadIdx.x == 0) {
        const size_t l = (size_t)pr * Bt.rows, c = l + Bt.rows;
        const TrackPose &T = Bt.poses[pr + 1];
        A.g = Bt.g, A.mode = 0, A.th = Bt.th, A.nnratio = 0, A.checkOri = Bt.checkOri, A.distTh = TH_HIGH;
        for (int k = 0; k < 9; k++) A.Rcw[k] = T.Rcw[k];
        for (int k = 0; k < 3; k++) A.tcw[k] = T.tcw[k], A.Ow[k] = 0.f;
        A.logScale = 0.f, A.bForward = T.bForward, A.bBackward = T.bBackward;
        A.nq = min(Bt.counts[pr], min(Bt.rows, MAXK)), A.nc = min(Bt.counts[pr + 1], min(Bt.rows, MAXK));
        A.q_valid = Bt.valid + l, A.q_obs = Bt.zeros, A.q_desc = Bt.desc + 32 * l, A.q_f3 = Bt.world + 3 * l, A.q_f2 = nullptr;
        A.q_level = Bt.octave + l, A.q_aux = Bt.angle + l;
        A.c_xy = Bt.xyUn + 2 * c, A.c_angle = Bt.angle + c, A.c_uright = Bt.uright + c, A.c_octave = Bt.octave + c;
        A.c_desc = Bt.desc + 32 * c, A.c_occ = Bt.zeros;
        A.c_match = Bt.match + l, A.nmatches = Bt.nmatches + pr, A.assign = Bt.assign + l, A.bin = Bt.bin + l;
    }
    __syncthreads();
    search_body(A);
}

// ------------------------------------------------------------- vocabulary-node searches (SearchByBoW, SearchForTriangulation)
__device__ __forceinline__ void three_maxima(const int *hist, int &ind1, int &ind2, int &ind3) {  // :799-830
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = hist[i];
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
    if ((float)max2 < 0.1f * (float)max1) ind2 = -1, ind3 = -1;
    else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
}

struct NodeArgs {
    int mode;  // 0: SearchByBoW (A = KeyFrame, B = Frame; out indexed by B), 1: SearchForTriangulation (A = KF1, B = KF2; out by A)
    float nnratio;
    int checkOri, onlyStereo;
    int nNodesA, nNodesB, nA, nB;
    const uint32_t *idA, *idB;
    const int32_t *offA, *featA, *offB, *featB;
    const uint8_t *a_flag, *a_desc;  // mode 0: kf_valid; mode 1: has_mp1
    const float *a_angle, *a_xy, *a_uright;
    const uint8_t *b_flag, *b_desc;  // mode 1: has_mp2
    const float *b_angle, *b_xy, *b_uright;
    const int32_t *b_octave;
    float F12[9], ex, ey, scale2[16], sigma2[16];
    int32_t *out, *bin, *nmatches;
};

__device__ __forceinline__ void node_search_body(const NodeArgs &A) {
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_n, s_ind[3];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int nOut = A.mode == 0 ? A.nB : A.nA;
    for (int j = tid; j < nOut; j += blockDim.x) A.out[j] = -1, A.bin[j] = -1;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const float factor = 1.0f / HISTO_LENGTH;
    for (int na = wid; na < A.nNodesA; na += nw) {
        // the node of the same id on the B side (the reference's merge with lower_bound visits exactly the common ids)
        const uint32_t id = A.idA[na];
        int lo = 0, hi = A.nNodesB;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A.idB[mid] < id) lo = mid + 1; else hi = mid;
        }
        if (lo >= A.nNodesB || A.idB[lo] != id) continue;
        const int b0 = A.offB[lo], b1 = A.offB[lo + 1];
        for (int ia = A.offA[na]; ia < A.offA[na + 1]; ia++) {
            const int idxA = A.featA[ia];
            if (A.mode == 0) {
                if (!A.a_flag[idxA]) continue;  // !pMP || pMP->isBad()
            } else {
                if (A.a_flag[idxA]) continue;   // pMP1 != NULL
            }
            const bool stereoA = A.mode == 1 && A.a_uright[idxA] >= 0;
            if (A.mode == 1 && A.onlyStereo && !stereoA) continue;
            const uint4 *Q = (const uint4 *)(A.a_desc + 32 * (size_t)idxA);
            const uint4 q0 = __ldg(Q), q1 = __ldg(Q + 1);
            int bd = A.mode == 0 ? 256 : TH_LOW + 1, bp = -1, sd = 256;
            float la = 0, lb = 0, lc = 0, den = 0, ax = 0, ay = 0;
            if (A.mode == 1) {  // epipolar line of kp1 in the second image (CheckDistEpipolarLine :127-144)
                ax = A.a_xy[2 * idxA], ay = A.a_xy[2 * idxA + 1];
                la = ax * A.F12[0] + ay * A.F12[3] + A.F12[6];
                lb = ax * A.F12[1] + ay * A.F12[4] + A.F12[7];
                lc = ax * A.F12[2] + ay * A.F12[5] + A.F12[8];
                den = la * la + lb * lb;
            }
            for (int ib = b0 + lane; ib < b1; ib += 32) {
                const int idxB = A.featB[ib];
                if (A.mode == 0) {
                    if (A.out[idxB] >= 0) continue;  // vpMapPointMatches[realIdxF] set by an earlier query of this node
                } else {
                    if (A.b_flag[idxB]) continue;     // pMP2 != NULL (vbMatched2 is never set in the reference)
                    if (A.onlyStereo && !(A.b_uright[idxB] >= 0)) continue;
                }
                const uint4 *T = (const uint4 *)(A.b_desc + 32 * (size_t)idxB);
                const int d = hamming256(q0, q1, __ldg(T), __ldg(T + 1));
                if (A.mode == 0) {
                    if (d < bd) sd = bd, bd = d, bp = ib;
                    else if (d < sd) sd = d;
                } else {
                    if (d > TH_LOW) continue;
                    const float bx = A.b_xy[2 * idxB], by = A.b_xy[2 * idxB + 1];
                    const int oct = A.b_octave[idxB];
                    if (!stereoA && !(A.b_uright[idxB] >= 0)) {
                        const float distex = A.ex - bx, distey = A.ey - by;
                        if (distex * distex + distey * distey < 100.0f * A.scale2[oct]) continue;
                    }
                    const float num = la * bx + lb * by + lc;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (!((double)dsqr < 3.84 * (double)A.sigma2[oct])) continue;
                    if (d <= bd) bd = d, bp = ib;  // `dist > bestDist` (:336): the later of two equal candidates wins
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int obd = __shfl_xor_sync(0xffffffffu, bd, o), obp = __shfl_xor_sync(0xffffffffu, bp, o);
                const int osd = __shfl_xor_sync(0xffffffffu, sd, o);
                if (A.mode == 0) {
                    const int ns = min(max(bd, obd), min(sd, osd));
                    if (obd < bd || (obd == bd && (unsigned)obp < (unsigned)bp)) bd = obd, bp = obp;
                    sd = ns;
                } else {
                    if (obd < bd || (obd == bd && obp > bp)) bd = obd, bp = obp;
                }
            }
            bool hit;
            if (A.mode == 0) hit = bp >= 0 && bd <= TH_LOW && (float)bd < A.nnratio * (float)sd;
            else hit = bp >= 0;
            if (hit && lane == 0) {
                const int idxB = A.featB[bp];
                const int slot = A.mode == 0 ? idxB : idxA;
                A.out[slot] = A.mode == 0 ? idxA : idxB;
                atomicAdd(&s_n, 1);
                if (A.checkOri) {
                    float rot = A.a_angle[idxA] - A.b_angle[idxB];
                    if (rot < 0.0f) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    A.bin[slot] = bin;
                    atomicAdd(&hist[bin], 1);
                }
            }
            __syncwarp();  // lane 0's write of out[] is visible to the warp's next query
        }
    }
    __syncthreads();
    if (A.checkOri) {
        if (tid == 0) three_maxima(hist, s_ind[0], s_ind[1], s_ind[2]);
        __syncthreads();
        int pruned = 0;
        for (int j = tid; j < nOut; j += blockDim.x) {
            const int b = A.bin[j];
            if (b >= 0 && b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) {
                A.out[j] = -3;
                pruned++;
            }
        }
        if (pruned) atomicSub(&s_n, pruned);
        __syncthreads();
    }
    if (tid == 0) *A.nmatches = s_n;
}
__global__ void __launch_bounds__(1024) k_node_search(const NodeArgs *__restrict__ args) {  // one CTA per recorded call
    __shared__ NodeArgs A;
    if (threadIdx.x < sizeof(NodeArgs) / 4) reinterpret_cast<uint32_t *>(&A)[threadIdx.x] = reinterpret_cast<const uint32_t *>(args + blockIdx.x)[threadIdx.x];
    __syncthreads();
    node_search_body(A);
}

// ------------------------------------------------------------------------------- Fuse (search part)
struct FuseArgs {
    msl_frame_geom g;
    float th, logScale;
    float Rcw[9], tcw[3], Ow[3], invSigma2[16];
    int nq, nc;
    const uint8_t *q_valid, *q_desc;
    const float *q_world, *q_normal, *q_dist;
    const float *c_xy, *c_uright;
    const int32_t *c_octave;
    const uint8_t *c_desc;
    int32_t *bestIdx, *bestDist, *nFused;
};

__device__ __forceinline__ void fuse_search_body(const FuseArgs &A) {
    __shared__ uint32_t keys[MAXK];  // (cell << 12 | index), sorted; 0xffffffff = not in grid
    __shared__ unsigned short cellStart[NCELLS + 1];
    __shared__ int s_n;
    const int tid = threadIdx.x, nt = blockDim.x;
    const msl_frame_geom &g = A.g;
    // the KeyFrame's grid (copied from its Frame: AssignFeaturesToGrid, src/Frame.cc:155-168) as a sorted list
    int n2 = 1;
    while (n2 < A.nc) n2 <<= 1;
    for (int i = tid; i < n2; i += nt) {
        uint32_t key = 0xffffffffu;
        if (i < A.nc) {
            const int px = (int)roundf((A.c_xy[2 * i] - g.mnMinX) * g.gridWInv);
            const int py = (int)roundf((A.c_xy[2 * i + 1] - g.mnMinY) * g.gridHInv);
            if (!(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS)) key = ((uint32_t)(px * GRID_ROWS + py) << 12) | (uint32_t)i;
        }
        keys[i] = key;
    }
    if (tid == 0) s_n = 0;
    __syncthreads();
    bitonic_sort_u32(keys, n2);
    for (int c = tid; c <= NCELLS; c += nt) {
        int lo = 0, hi = A.nc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((keys[mid] >> 12) < (uint32_t)c) lo = mid + 1; else hi = mid;
        }
        cellStart[c] = (unsigned short)lo;
    }
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < A.nq; i += nt) {
        int best = -1, bestDist = 256;
        if (A.q_valid[i]) {
            const float *p3Dw = A.q_world + 3 * i;
            const float pc0 = gemm_row(A.Rcw, p3Dw, A.tcw[0]), pc1 = gemm_row(A.Rcw + 3, p3Dw, A.tcw[1]);
            const float pc2 = gemm_row(A.Rcw + 6, p3Dw, A.tcw[2]);
            bool ok = !(pc2 < 0.0f);  // :431
            const float invz = 1.0f / pc2;
            const float x = pc0 * invz, y = pc1 * invz;
            const float u = g.fx * x + g.cx, v = g.fy * y + g.cy;
            if (!(u >= g.mnMinX && u < g.mnMaxX && v >= g.mnMinY && v < g.mnMaxY)) ok = false;  // KeyFrame::IsInImage
            const float ur = u - g.mbf * invz;
            const float minD = 0.8f * A.q_dist[2 * i], maxD = 1.2f * A.q_dist[2 * i + 1];
            float PO[3];
            double s2 = 0;
            for (int k = 0; k < 3; k++) {
                PO[k] = p3Dw[k] - A.Ow[k];
                s2 += (double)PO[k] * (double)PO[k];
            }
            const float dist3D = (float)sqrt(s2);
            if (dist3D < minD || dist3D > maxD) ok = false;
            double dot = 0;
            for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)A.q_normal[3 * i + k];
            if (dot < 0.5 * (double)dist3D) ok = false;  // viewing angle > 60 deg (:459-462)
            if (ok) {
                // MapPoint::PredictScale(dist3D, pKF) src/MapPoint.cc:334-348 (logf as the rounded fp64 logarithm, see k_search)
                const float ratio = A.q_dist[2 * i + 1] / dist3D;
                int lvl = (int)ceilf((float)log((double)ratio) / A.logScale);
                if (lvl < 0) lvl = 0;
                else if (lvl >= g.nlevels) lvl = g.nlevels - 1;
                const float radius = A.th * g.scaleFactors[lvl];
                // KeyFrame::GetFeaturesInArea src/KeyFrame.cc:469-504
                const int cx0 = max(0, (int)floorf((u - g.mnMinX - radius) * g.gridWInv));
                const int cx1 = min(GRID_COLS - 1, (int)ceilf((u - g.mnMinX + radius) * g.gridWInv));
                const int cy0 = max(0, (int)floorf((v - g.mnMinY - radius) * g.gridHInv));
                const int cy1 = min(GRID_ROWS - 1, (int)ceilf((v - g.mnMinY + radius) * g.gridHInv));
                if (cx0 < GRID_COLS && cx1 >= 0 && cy0 < GRID_ROWS && cy1 >= 0) {
                    const uint4 *Q = (const uint4 *)(A.q_desc + 32 * (size_t)i);
                    const uint4 q0 = Q[0], q1 = Q[1];
                    for (int ix = cx0; ix <= cx1; ix++) {
                        const int s0 = cellStart[ix * GRID_ROWS + cy0], s1 = cellStart[ix * GRID_ROWS + cy1 + 1];
                        for (int s = s0; s < s1; s++) {
                            const int k = keys[s] & 0xfff;
                            const float kpx = A.c_xy[2 * k], kpy = A.c_xy[2 * k + 1];
                            if (!(fabsf(kpx - u) < radius && fabsf(kpy - v) < radius)) continue;
                            const int kpLevel = A.c_octave[k];
                            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
                            const float ex = u - kpx, ey = v - kpy;
                            const float kpr = A.c_uright[k];
                            if (kpr >= 0) {  // chi-square gate, stereo (:488-499) / mono (:500-509)
                                const float er = ur - kpr;
                                const float e2 = ex * ex + ey * ey + er * er;
                                if ((double)(e2 * A.invSigma2[kpLevel]) > 7.8) continue;
                            } else {
                                const float e2 = ex * ex + ey * ey;
                                if ((double)(e2 * A.invSigma2[kpLevel]) > 5.99) continue;
                            }
                            const uint4 *T = (const uint4 *)(A.c_desc + 32 * (size_t)k);
                            const int d = hamming256(q0, q1, T[0], T[1]);
                            if (d < bestDist) bestDist = d, best = k;
                        }
                    }
                }
            }
        }
        A.bestIdx[i] = best, A.bestDist[i] = bestDist;
        if (bestDist <= TH_LOW) mine++;
    }
    if (mine) atomicAdd(&s_n, mine);
    __syncthreads();
    if (tid == 0) *A.nFused = s_n;
}
__global__ void __launch_bounds__(1024) k_fuse_search(const FuseArgs *__restrict__ args) {  // one CTA per recorded call
    __shared__ FuseArgs A;
    if (threadIdx.x < sizeof(FuseArgs) / 4) reinterpret_cast<uint32_t *>(&A)[threadIdx.x] = reinterpret_cast<const uint32_t *>(args + blockIdx.x)[threadIdx.x];
    __syncthreads();
    fuse_search_body(A);
}

// --------------------------------------------------- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263), batched
// One CTA (4 warps) per map point, one warp per row i of the point's N x N distance matrix.  The row's median -- the
// element of rank r = (int)(0.5 * (N - 1)) of the sorted row, diagonal 0 included -- is found without sorting: it is
// the smallest v in [0, 256] with #{j : d(i, j) <= v} > r, by bisection on v with one warp-wide count per step (at most
// 9 steps; the distances are recomputed, 8 xor + popc each).  BestIdx = the first row with the smallest median (:247-251).
__global__ void __launch_bounds__(128)
    k_distinctive(const int32_t *__restrict__ off, const uint8_t *__restrict__ desc, int32_t *__restrict__ bestIdx,
                  int32_t *__restrict__ bestMedian) {
    __shared__ int s_med[4], s_row[4];
    const int p = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int o0 = off[p], N = off[p + 1] - o0;
    if (N <= 0) {  // uniform for the CTA: the reference returns without touching the map point
        if (threadIdx.x == 0) bestIdx[p] = -1, bestMedian[p] = 0x7fffffff;
        return;
    }
    const uint4 *D = (const uint4 *)(desc + 32 * (size_t)o0);
    const int r = (int)(0.5 * (double)(N - 1));
    int myMed = 0x7fffffff, myRow = -1;
    for (int i = wid; i < N; i += 4) {
        const uint4 a0 = __ldg(D + 2 * i), a1 = __ldg(D + 2 * i + 1);
        int lo = 0, hi = 256;  // N > r elements are <= 256: the answer lies in [lo, hi]
        while (lo < hi) {      // lo, hi are warp-uniform
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j = lane; j < N; j += 32) c += hamming256(a0, a1, __ldg(D + 2 * j), __ldg(D + 2 * j + 1)) <= mid;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c > r) hi = mid;
            else lo = mid + 1;
        }
        if (lo < myMed) myMed = lo, myRow = i;  // a warp's rows ascend: strict < keeps the first
    }
    if (lane == 0) s_med[wid] = myMed, s_row[wid] = myRow;
    __syncthreads();
    if (threadIdx.x == 0) {
        int bm = 0x7fffffff, br = -1;
        for (int w = 0; w < 4; w++)
            if (s_row[w] >= 0 && (br < 0 || s_med[w] < bm || (s_med[w] == bm && s_row[w] < br))) bm = s_med[w], br = s_row[w];
        bestIdx[p] = br, bestMedian[p] = bm;
    }
}

}  // namespace

struct msl_matcher {
    int maxQ, maxT, maxBatch, device;
    cudaStream_t stream = nullptr;
    uint8_t *d_q = nullptr, *d_t = nullptr;
    int32_t *d_bi = nullptr, *d_bd = nullptr, *d_sd = nullptr;
    uint16_t *d_dist = nullptr;
    size_t distCap = 0;
    uint8_t *d_scr = nullptr;  // arena for the window-search arrays
    uint8_t *h_scr = nullptr;  // its pinned host mirror: a call's inputs are packed here and go up in ONE copy
    size_t scrCap = 0;
    uint8_t *d_dd = nullptr;   // msl_distinctive_descriptors: offsets | descriptors | outputs (grown on demand)
    size_t ddCap = 0;
    uint8_t *d_trk = nullptr;  // msl_search_by_projection_frames_dev: per-frame scratch (grown on demand)
    size_t trkCap = 0;
    // deferred execution (msl_matcher_batch_begin / _end; a single call is a batch of one): the calls recorded so far
    bool batching = false;
    size_t batchOff = 0;  // arena bytes taken by the recorded calls
    std::vector<SearchArgs> bSearch;
    std::vector<NodeArgs> bNode;
    std::vector<FuseArgs> bFuse;
    struct Out {
        void *host;
        const void *dev;
        size_t bytes;
    };
    struct Pending {
        const void *first;  // the call's outputs are contiguous in the arena: [first, first + bytes)
        size_t bytes;
        Out out[3];
    };
    std::vector<Pending> pending;
    // optional timing of an execution on the stream (msl_matcher_set_timing): upload + kernels + download
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float lastMs = 0.f;
    int lastCalls = 0;
};

static void matcher_free(msl_matcher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    void *ptrs[] = {m->d_q, m->d_t, m->d_bi, m->d_bd, m->d_sd, m->d_dist, m->d_scr, m->d_dd, m->d_trk};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (m->h_scr) cudaFreeHost(m->h_scr);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

// Bump allocator over the scratch arena.  A search takes a dozen small host arrays: copied one by one from pageable memory
// they cost more than the kernel (each cudaMemcpyAsync of a pageable buffer is a staged, effectively synchronous copy), so
// put() packs them into the arena's pinned host mirror at the offsets they will have on the device and flush() sends the
// packed range up in ONE copy; the outputs (allocated with get() after the inputs, so they are contiguous) come back in
// one copy into the mirror (fetch) and are handed out from there.
struct Arena {
    uint8_t *base;
    size_t off, cap;
    cudaStream_t st;
    uint8_t *hbase;
    bool ok = true;
    size_t inEnd = 0;
    template <typename T>
    const T *put(const T *h, size_t n) {
        off = align_up(off, 16);
        if (off + n * sizeof(T) > cap) { ok = false; return nullptr; }
        T *d = (T *)(base + off);
        if (n) memcpy(hbase + off, h, n * sizeof(T));
        off += n * sizeof(T);
        inEnd = off;
        return d;
    }
    template <typename T>
    T *get(size_t n) {
        off = align_up(off, 16);
        if (off + n * sizeof(T) > cap) { ok = false; return nullptr; }
        T *d = (T *)(base + off);
        off += n * sizeof(T);
        return d;
    }
    bool flush() {  // every put() so far -> device
        if (!ok) return false;
        if (inEnd && cudaMemcpyAsync(base, hbase, inEnd, cudaMemcpyHostToDevice, st) != cudaSuccess) ok = false;
        return ok;
    }
    // [first, first + bytes) of the arena -> the mirror, then wait; host(p) is the mirror address of device address p
    bool fetch(const void *first, size_t bytes) {
        const size_t o = (const uint8_t *)first - base;
        if (cudaMemcpyAsync(hbase + o, first, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
        return cudaStreamSynchronize(st) == cudaSuccess;
    }
    template <typename T>
    const T *host(const T *d) const { return (const T *)(hbase + ((const uint8_t *)d - base)); }
};

// The scratch arena of the searches grows with the call: the reference hands SearchByProjection the whole of
// mvpLocalMapPoints (src/Tracking.cc:1693) and Fuse the map points of every neighbour keyframe (src/LocalMapping.cc:569) --
// tens of thousands of points -- so only the per-frame keypoint count (MAXK, the kernels' shared-memory grid) is a hard
// limit; max_queries / max_train of msl_matcher_create size the initial allocation and the Hamming batch buffers.
static int matcher_execute(msl_matcher *m);
static int ensure_scratch(msl_matcher *m, size_t n_a, size_t n_b) {
    const size_t need = (n_a + n_b) * 96 + 4096;
    if (m->batchOff + need <= m->scrCap) return MSL_OK;
    size_t want = need + need / 2;
    if (!m->pending.empty()) {
        // a deferred batch has filled the arena: run what is recorded, start over at offset 0 -- and quadruple the arena
        // (up to 256 MB) so that the caller's next batch of this size runs as ONE execution
        const int rc = matcher_execute(m);
        if (rc) return rc;
        want = std::max(want, std::min(m->scrCap * 4, (size_t)256 << 20));
    } else if (need <= m->scrCap)
        return MSL_OK;
    if (want <= m->scrCap) return MSL_OK;
    MSL_CUDA(cudaStreamSynchronize(m->stream));
    if (m->d_scr) cudaFree(m->d_scr);
    if (m->h_scr) cudaFreeHost(m->h_scr);
    m->d_scr = nullptr, m->h_scr = nullptr, m->scrCap = 0;
    const size_t cap = want;
    MSL_CUDA(cudaMalloc((void **)&m->d_scr, cap));
    MSL_CUDA(cudaMallocHost((void **)&m->h_scr, cap));
    m->scrCap = cap;
    return MSL_OK;
}

static_assert(sizeof(SearchArgs) <= 1024 && sizeof(NodeArgs) <= 1024 && sizeof(FuseArgs) <= 1024 && sizeof(SearchArgs) % 4 == 0 &&
                  sizeof(NodeArgs) % 4 == 0 && sizeof(FuseArgs) % 4 == 0,
              "argument records: loaded by one CTA pass, and the 4096 spare bytes per call of ensure_scratch cover them");
// Runs every recorded call: the packed inputs and the argument records go up in one copy, one launch per kind of search with
// one CTA per call, the outputs come back with one copy per call into the pinned mirror and are handed to the callers'
// arrays after a single synchronisation.
static int matcher_execute(msl_matcher *m) {
    if (m->pending.empty()) {
        m->batchOff = 0;
        return MSL_OK;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    ar.inEnd = m->batchOff;
    const SearchArgs *dS = m->bSearch.empty() ? nullptr : ar.put(m->bSearch.data(), m->bSearch.size());
    const NodeArgs *dN = m->bNode.empty() ? nullptr : ar.put(m->bNode.data(), m->bNode.size());
    const FuseArgs *dF = m->bFuse.empty() ? nullptr : ar.put(m->bFuse.data(), m->bFuse.size());
    const size_t nS = m->bSearch.size(), nN = m->bNode.size(), nF = m->bFuse.size();
    std::vector<msl_matcher::Pending> pend;
    pend.swap(m->pending);
    m->bSearch.clear(), m->bNode.clear(), m->bFuse.clear();
    m->batchOff = 0;
    if (m->timing) MSL_CUDA(cudaEventRecord(m->ev0, m->stream));
    if (!ar.flush()) return fail(MSL_ERR_CUDA, "matcher: scratch arena / input copy failure");
    if (nS) {
        k_search_many<<<(unsigned)nS, 1024, 0, m->stream>>>(dS);
        MSL_LAUNCH_CHECK();
    }
    if (nN) {
        k_node_search<<<(unsigned)nN, 1024, 0, m->stream>>>(dN);
        MSL_LAUNCH_CHECK();
    }
    if (nF) {
        k_fuse_search<<<(unsigned)nF, 1024, 0, m->stream>>>(dF);
        MSL_LAUNCH_CHECK();
    }
    for (const auto &p : pend)
        MSL_CUDA(cudaMemcpyAsync(m->h_scr + ((const uint8_t *)p.first - m->d_scr), p.first, p.bytes, cudaMemcpyDeviceToHost, m->stream));
    if (m->timing) MSL_CUDA(cudaEventRecord(m->ev1, m->stream));
    MSL_CUDA(cudaStreamSynchronize(m->stream));
    if (m->timing) {
        MSL_CUDA(cudaEventElapsedTime(&m->lastMs, m->ev0, m->ev1));
        m->lastCalls = (int)pend.size();
    }
    for (const auto &p : pend)
        for (const auto &o : p.out)
            if (o.host && o.bytes) memcpy(o.host, m->h_scr + ((const uint8_t *)o.dev - m->d_scr), o.bytes);
    return MSL_OK;
}

// records a call (its inputs are packed, its outputs allocated); executes at once unless a deferred batch is open.
// The three arrays of `outs` are adjacent in the arena, in this order.
static int matcher_record(msl_matcher *m, const Arena &ar, msl_matcher::Out o0, msl_matcher::Out o1, msl_matcher::Out o2) {
    if (!ar.ok) return fail(MSL_ERR_CUDA, "matcher: scratch arena overflow");
    msl_matcher::Pending p;
    p.first = o0.dev;
    p.bytes = (const uint8_t *)o2.dev + o2.bytes - (const uint8_t *)o0.dev;
    p.out[0] = o0, p.out[1] = o1, p.out[2] = o2;
    m->pending.push_back(p);
    m->batchOff = align_up(ar.off, 256);
    return m->batching ? MSL_OK : matcher_execute(m);
}

// c_match and nmatches are adjacent in the arena (allocated in that order by every caller)
static int run_search(msl_matcher *m, Arena &ar, SearchArgs &A, int32_t *cur_match, int32_t *nmatches) {
    m->bSearch.push_back(A);
    const int rc = matcher_record(m, ar, {cur_match, A.c_match, sizeof(int32_t) * A.nc}, {nmatches, A.nmatches, sizeof(int32_t)}, {nullptr, A.nmatches, sizeof(int32_t)});
    if (rc) m->bSearch.clear(), m->bNode.clear(), m->bFuse.clear(), m->pending.clear(), m->batchOff = 0;
    return rc;
}

extern "C" {

int msl_matcher_create(int max_queries, int max_train, int max_batch, int device, msl_matcher **out) {
    if (!out) return fail(MSL_ERR_INVALID, "msl_matcher_create: null out");
    *out = nullptr;
    if (max_queries < 1 || max_train < 1 || max_batch < 1) return fail(MSL_ERR_INVALID, "msl_matcher_create: parameter out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0)
        return fail(MSL_ERR_CUDA, "msl_matcher_create: no usable CUDA device (there is no CPU fallback)");
    MSL_CUDA(cudaSetDevice(device));
    msl_matcher *m = new msl_matcher();
    m->maxQ = max_queries, m->maxT = max_train, m->maxBatch = max_batch, m->device = device;
    const size_t B = max_batch;
    cudaError_t e = cudaMalloc((void **)&m->d_q, B * max_queries * 32);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_t, B * max_train * 32);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_bi, B * max_queries * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_bd, B * max_queries * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_sd, B * max_queries * 4);
    m->scrCap = (size_t)(max_queries + max_train) * 96 + 4096;
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_scr, m->scrCap);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&m->h_scr, m->scrCap);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        matcher_free(m);
        return fail(MSL_ERR_CUDA, std::string("msl_matcher_create: ") + cudaGetErrorString(e));
    }
    *out = m;
    return MSL_OK;
}

void msl_matcher_destroy(msl_matcher *m) { matcher_free(m); }
void *msl_matcher_stream(msl_matcher *m) { return m ? (void *)m->stream : nullptr; }
int msl_matcher_sync(msl_matcher *m) {
    if (!m) return fail(MSL_ERR_INVALID, "null handle");
    MSL_CUDA(cudaSetDevice(m->device));
    MSL_CUDA(cudaStreamSynchronize(m->stream));
    return MSL_OK;
}

int msl_matcher_set_timing(msl_matcher *m, int on) {
    if (!m) return fail(MSL_ERR_INVALID, "null handle");
    MSL_CUDA(cudaSetDevice(m->device));
    if (on && !m->ev0) {
        MSL_CUDA(cudaEventCreate(&m->ev0));
        MSL_CUDA(cudaEventCreate(&m->ev1));
    }
    m->timing = on != 0;
    return MSL_OK;
}
int msl_matcher_last_execution(msl_matcher *m, double *device_ms, int *calls) {
    if (!m || !device_ms || !calls) return fail(MSL_ERR_INVALID, "msl_matcher_last_execution: null argument");
    *device_ms = m->lastMs, *calls = m->lastCalls;
    return MSL_OK;
}
int msl_matcher_batch_begin(msl_matcher *m) {
    if (!m) return fail(MSL_ERR_INVALID, "null handle");
    if (m->batching) return fail(MSL_ERR_STATE, "msl_matcher_batch_begin: a batch is already open");
    m->batching = true;
    return MSL_OK;
}
int msl_matcher_batch_end(msl_matcher *m) {
    if (!m) return fail(MSL_ERR_INVALID, "null handle");
    if (!m->batching) return fail(MSL_ERR_STATE, "msl_matcher_batch_end: no open batch");
    m->batching = false;
    MSL_CUDA(cudaSetDevice(m->device));
    const int rc = matcher_execute(m);
    if (rc) m->bSearch.clear(), m->bNode.clear(), m->bFuse.clear(), m->pending.clear(), m->batchOff = 0;
    return rc;
}

int msl_hamming_best2_dev(msl_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, int batch,
                          int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_dist) {
    if (!m || !d_q || !d_t || !d_best_idx || !d_best_dist || !d_second_dist) return fail(MSL_ERR_INVALID, "msl_hamming_best2_dev: null argument");
    if (nq < 1 || nt < 0 || batch < 1) return fail(MSL_ERR_INVALID, "msl_hamming_best2_dev: bad size");
    MSL_CUDA(cudaSetDevice(m->device));
    k_hamming_best2<<<dim3(cdiv(nq, 8), batch), 256, 0, m->stream>>>(d_q, nq, nq, d_t, nt, nt, nullptr, nullptr, d_best_idx,
                                                                      d_best_dist, d_second_dist);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

int msl_hamming_best2_counts_dev(msl_matcher *m, const uint8_t *d_q, const uint8_t *d_t, int rows, const int32_t *d_qcounts,
                                 const int32_t *d_tcounts, int batch, int32_t *d_best_idx, int32_t *d_best_dist,
                                 int32_t *d_second_dist, void *stream) {
    if (!m || !d_q || !d_t || !d_best_idx || !d_best_dist || !d_second_dist) return fail(MSL_ERR_INVALID, "msl_hamming_best2_counts_dev: null argument");
    if (rows < 1 || batch < 1) return fail(MSL_ERR_INVALID, "msl_hamming_best2_counts_dev: bad size");
    MSL_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    k_hamming_best2<<<dim3(cdiv(rows, 8), batch), 256, 0, st>>>(d_q, rows, rows, d_t, rows, rows, d_qcounts, d_tcounts, d_best_idx,
                                                                 d_best_dist, d_second_dist);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

int msl_hamming_best2(msl_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt, int batch, int32_t *best_idx,
                      int32_t *best_dist, int32_t *second_dist) {
    if (!m || !q || !t || !best_idx || !best_dist || !second_dist) return fail(MSL_ERR_INVALID, "msl_hamming_best2: null argument");
    if (nq < 1 || nq > m->maxQ || nt < 1 || nt > m->maxT || batch < 1 || batch > m->maxBatch) return fail(MSL_ERR_INVALID, "msl_hamming_best2: bad size");
    MSL_CUDA(cudaSetDevice(m->device));
    MSL_CUDA(cudaMemcpyAsync(m->d_q, q, (size_t)batch * nq * 32, cudaMemcpyHostToDevice, m->stream));
    MSL_CUDA(cudaMemcpyAsync(m->d_t, t, (size_t)batch * nt * 32, cudaMemcpyHostToDevice, m->stream));
    int rc = msl_hamming_best2_dev(m, m->d_q, nq, m->d_t, nt, batch, m->d_bi, m->d_bd, m->d_sd);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(best_idx, m->d_bi, (size_t)batch * nq * 4, cudaMemcpyDeviceToHost, m->stream));
    MSL_CUDA(cudaMemcpyAsync(best_dist, m->d_bd, (size_t)batch * nq * 4, cudaMemcpyDeviceToHost, m->stream));
    MSL_CUDA(cudaMemcpyAsync(second_dist, m->d_sd, (size_t)batch * nq * 4, cudaMemcpyDeviceToHost, m->stream));
    MSL_CUDA(cudaStreamSynchronize(m->stream));
    return MSL_OK;
}

int msl_hamming_all_pairs(msl_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt, int batch, uint16_t *dist) {
    if (!m || !q || !t || !dist) return fail(MSL_ERR_INVALID, "msl_hamming_all_pairs: null argument");
    if (nq < 1 || nq > m->maxQ || nt < 1 || nt > m->maxT || batch < 1 || batch > m->maxBatch) return fail(MSL_ERR_INVALID, "msl_hamming_all_pairs: bad size");
    MSL_CUDA(cudaSetDevice(m->device));
    const size_t need = (size_t)batch * nq * nt * 2;
    if (need > m->distCap) {
        if (m->d_dist) cudaFree(m->d_dist);
        m->d_dist = nullptr;
        MSL_CUDA(cudaMalloc((void **)&m->d_dist, need));
        m->distCap = need;
    }
    MSL_CUDA(cudaMemcpyAsync(m->d_q, q, (size_t)batch * nq * 32, cudaMemcpyHostToDevice, m->stream));
    MSL_CUDA(cudaMemcpyAsync(m->d_t, t, (size_t)batch * nt * 32, cudaMemcpyHostToDevice, m->stream));
    k_hamming_all_pairs<<<dim3(cdiv(nt, 256), nq, batch), 256, 0, m->stream>>>(m->d_q, nq, m->d_t, nt, m->d_dist);
    MSL_LAUNCH_CHECK();
    MSL_CUDA(cudaMemcpyAsync(dist, m->d_dist, need, cudaMemcpyDeviceToHost, m->stream));
    MSL_CUDA(cudaStreamSynchronize(m->stream));
    return MSL_OK;
}

int msl_search_by_projection_frame(msl_matcher *m, const msl_frame_geom *geom, const float Tcw_cur[16],
                                   const float Tcw_last[16], float th, int check_orientation, int n_last,
                                   const uint8_t *last_has_mp, const uint8_t *last_outlier, const uint8_t *last_mp_obs,
                                   const float *last_mp_world, const uint8_t *last_mp_desc, const int32_t *last_octave,
                                   const float *last_angle, int n_cur, const float *cur_xy, const int32_t *cur_octave,
                                   const float *cur_angle, const float *cur_uright, const uint8_t *cur_desc,
                                   const uint8_t *cur_occupied, int32_t *cur_match, int32_t *nmatches) {
    if (!m || !geom || !Tcw_cur || !Tcw_last || !cur_match || !nmatches) return fail(MSL_ERR_INVALID, "msl_search_by_projection_frame: null argument");
    if (n_last < 0 || n_cur < 0 || n_cur > MAXK || n_last > 0x7ffffff)
        return fail(MSL_ERR_INVALID, "msl_search_by_projection_frame: too many keypoints for this handle");
    if (n_cur == 0 || n_last == 0) {
        for (int j = 0; j < n_cur; j++) cur_match[j] = cur_occupied[j] ? -2 : -1;
        *nmatches = 0;
        return MSL_OK;
    }
    MSL_CUDA(cudaSetDevice(m->device));
    SearchArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *geom, A.mode = 0, A.th = th, A.nnratio = 0, A.checkOri = check_orientation, A.nq = n_last, A.nc = n_cur;
    A.distTh = TH_HIGH;
    // :554-568: twc = -Rcw.t() * tcw (transposed operand: cv::gemm's general path, double accumulation, one rounding);
    // tlc = Rlw * twc + tlw (flags == 0: the small-matrix path, gemm_row)
    float Rlw[9], tlw[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) A.Rcw[r * 3 + c] = Tcw_cur[r * 4 + c], Rlw[r * 3 + c] = Tcw_last[r * 4 + c];
        A.tcw[r] = Tcw_cur[r * 4 + 3], tlw[r] = Tcw_last[r * 4 + 3];
    }
    float twc[3];
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)(-A.Rcw[k * 3 + r]) * (double)A.tcw[k];
        twc[r] = (float)s;
    }
    const float tlc2 = gemm_row(Rlw + 6, twc, tlw[2]);
    A.bForward = tlc2 > geom->mb;
    A.bBackward = -tlc2 > geom->mb;
    std::vector<uint8_t> valid(n_last);
    for (int i = 0; i < n_last; i++) valid[i] = last_has_mp[i] && !last_outlier[i];
    {
        const int rc_ = ensure_scratch(m, (size_t)(n_last), (size_t)(n_cur));
        if (rc_) return rc_;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    A.q_valid = ar.put(valid.data(), n_last);
    A.q_obs = ar.put(last_mp_obs, n_last);
    A.q_desc = ar.put(last_mp_desc, (size_t)n_last * 32);
    A.q_f3 = ar.put(last_mp_world, (size_t)n_last * 3);
    A.q_level = ar.put(last_octave, n_last);
    A.q_aux = ar.put(last_angle, n_last);
    A.c_xy = ar.put(cur_xy, (size_t)n_cur * 2);
    A.c_angle = ar.put(cur_angle, n_cur);
    A.c_uright = ar.put(cur_uright, n_cur);
    A.c_octave = ar.put(cur_octave, n_cur);
    A.c_desc = ar.put(cur_desc, (size_t)n_cur * 32);
    A.c_occ = ar.put(cur_occupied, n_cur);
    A.c_match = ar.get<int32_t>(n_cur);
    A.nmatches = ar.get<int32_t>(1);
    A.assign = ar.get<int32_t>(n_last);
    A.bin = ar.get<int32_t>(n_last);
    if (!ar.ok) return fail(MSL_ERR_CUDA, "msl_search_by_projection_frame: scratch arena / copy failure");
    return run_search(m, ar, A, cur_match, nmatches);
}

int msl_search_by_projection_frames_dev(msl_matcher *m, const msl_frame_geom *geom, float th, int check_orientation, float th_depth,
                                        const msl_keypoint *d_kps, const uint8_t *d_desc, int rows, const int32_t *d_counts,
                                        int n_frames, const float *d_xy_un, const float *d_uright, const float *d_kdepth,
                                        const float *Tcw, int32_t *d_cur_match, int32_t *d_nmatches, void *stream) {
    if (!m || !geom || !d_kps || !d_desc || !d_counts || !d_xy_un || !d_uright || !d_kdepth || !Tcw || !d_cur_match || !d_nmatches)
        return fail(MSL_ERR_INVALID, "msl_search_by_projection_frames_dev: null argument");
    if (n_frames < 2 || rows < 1 || rows > MAXK) return fail(MSL_ERR_INVALID, "msl_search_by_projection_frames_dev: bad size");
    MSL_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    const size_t F = (size_t)n_frames, R = (size_t)rows;
    // scratch: poses | valid | zeros | world | octave | angle | assign | bin
    const size_t oPose = 0, oValid = align_up(oPose + F * sizeof(TrackPose), 256), oZero = align_up(oValid + F * R, 256),
                 oWorld = align_up(oZero + R, 256), oOct = align_up(oWorld + F * R * 12, 256), oAng = align_up(oOct + F * R * 4, 256),
                 oAsg = align_up(oAng + F * R * 4, 256), oBin = align_up(oAsg + F * R * 4, 256), total = oBin + F * R * 4;
    if (total > m->trkCap) {
        MSL_CUDA(cudaStreamSynchronize(st));
        if (m->d_trk) cudaFree(m->d_trk);
        m->d_trk = nullptr, m->trkCap = 0;
        MSL_CUDA(cudaMalloc((void **)&m->d_trk, total));
        m->trkCap = total;
    }
    // per-frame pose blocks as the reference forms them: Rwc = Rcw.t(), Ow = -Rcw.t() * tcw (transposed operand: cv::gemm's
    // general path, double accumulation, one rounding; src/Frame.cc:305-311); per pair tlc = Rlw * twc + tlw (:554-568)
    std::vector<TrackPose> tp(F);
    for (size_t f = 0; f < F; f++) {
        const float *T = Tcw + 16 * f;
        TrackPose &P = tp[f];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) P.Rcw[r * 3 + c] = T[r * 4 + c], P.Rwc[c * 3 + r] = T[r * 4 + c];
            P.tcw[r] = T[r * 4 + 3];
        }
        for (int r = 0; r < 3; r++) {
            double sum = 0;
            for (int k = 0; k < 3; k++) sum += (double)(-P.Rcw[k * 3 + r]) * (double)P.tcw[k];
            P.Ow[r] = (float)sum;
        }
        P.bForward = P.bBackward = 0;
        if (f > 0) {  // this frame is the Current frame of pair f - 1
            const float *Tl = Tcw + 16 * (f - 1);
            const float Rlw2[3] = {Tl[8], Tl[9], Tl[10]};
            const float tlc2 = gemm_row(Rlw2, P.Ow, Tl[11]);
            P.bForward = tlc2 > geom->mb, P.bBackward = -tlc2 > geom->mb;
        }
    }
    MSL_CUDA(cudaMemcpyAsync(m->d_trk + oPose, tp.data(), F * sizeof(TrackPose), cudaMemcpyHostToDevice, st));
    MSL_CUDA(cudaMemsetAsync(m->d_trk + oZero, 0, R, st));
    uint8_t *valid = m->d_trk + oValid;
    float *world = (float *)(m->d_trk + oWorld), *angle = (float *)(m->d_trk + oAng);
    int32_t *octave = (int32_t *)(m->d_trk + oOct);
    const float invfx = 1.0f / geom->fx, invfy = 1.0f / geom->fy;  // Frame::invfx / invfy (src/Frame.cc:123-124)
    k_track_last<<<n_frames, 256, 0, st>>>(d_kps, rows, d_counts, d_xy_un, d_kdepth, (const TrackPose *)(m->d_trk + oPose), geom->cx,
                                           geom->cy, invfx, invfy, th_depth, valid, world, octave, angle);
    MSL_LAUNCH_CHECK();
    SearchBatch Bt;
    Bt.g = *geom, Bt.th = th, Bt.checkOri = check_orientation, Bt.rows = rows, Bt.counts = d_counts;
    Bt.poses = (const TrackPose *)(m->d_trk + oPose), Bt.valid = valid, Bt.zeros = m->d_trk + oZero, Bt.desc = d_desc;
    Bt.world = world, Bt.xyUn = d_xy_un, Bt.angle = angle, Bt.uright = d_uright, Bt.octave = octave;
    Bt.match = d_cur_match, Bt.nmatches = d_nmatches, Bt.assign = (int32_t *)(m->d_trk + oAsg), Bt.bin = (int32_t *)(m->d_trk + oBin);
    k_search_batch<<<n_frames - 1, 1024, 0, st>>>(Bt);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

// Host form of the batched search: extractor output, depth frames and poses in host memory.  Uploads them, runs the frame
// glue (Frame::ComputeStereoFromRGBD, undistortion-free camera: mvKeysUn = mvKeys), the Last-frame side and the search on
// the matcher's stream, downloads the match tables.  `glue` supplies the glue kernels (any handle created for this frame size).
int msl_search_by_projection_frames(msl_matcher *m, msl_glue *glue, const msl_frame_geom *geom, float th, int check_orientation,
                                    float th_depth, const msl_keypoint *kps, const uint8_t *desc, int rows, const int32_t *counts,
                                    int n_frames, const float *depth, int w, int h, const float *Tcw, int32_t *cur_match,
                                    int32_t *nmatches) {
    if (!m || !glue || !geom || !kps || !desc || !counts || !depth || !Tcw || !cur_match || !nmatches)
        return fail(MSL_ERR_INVALID, "msl_search_by_projection_frames: null argument");
    if (n_frames < 2 || rows < 1 || rows > MAXK || w < 1 || h < 1) return fail(MSL_ERR_INVALID, "msl_search_by_projection_frames: bad size");
    MSL_CUDA(cudaSetDevice(m->device));
    const size_t F = (size_t)n_frames, R = (size_t)rows, npx = (size_t)w * h;
    const size_t oK = 0, oD = align_up(oK + F * R * sizeof(msl_keypoint), 256), oC = align_up(oD + F * R * 32, 256),
                 oZ = align_up(oC + F * 4, 256), oXY = align_up(oZ + F * npx * 4, 256), oUR = align_up(oXY + F * R * 8, 256),
                 oKD = align_up(oUR + F * R * 4, 256), oCM = align_up(oKD + F * R * 4, 256), oNM = align_up(oCM + F * R * 4, 256),
                 total = oNM + F * 4;
    if (total > m->ddCap) {  // (shares the grow-on-demand staging buffer of msl_distinctive_descriptors)
        MSL_CUDA(cudaStreamSynchronize(m->stream));
        if (m->d_dd) cudaFree(m->d_dd);
        m->d_dd = nullptr, m->ddCap = 0;
        MSL_CUDA(cudaMalloc((void **)&m->d_dd, total));
        m->ddCap = total;
    }
    uint8_t *b = m->d_dd;
    cudaStream_t st = m->stream;
    MSL_CUDA(cudaMemcpyAsync(b + oK, kps, F * R * sizeof(msl_keypoint), cudaMemcpyHostToDevice, st));
    MSL_CUDA(cudaMemcpyAsync(b + oD, desc, F * R * 32, cudaMemcpyHostToDevice, st));
    MSL_CUDA(cudaMemcpyAsync(b + oC, counts, F * 4, cudaMemcpyHostToDevice, st));
    MSL_CUDA(cudaMemcpyAsync(b + oZ, depth, F * npx * 4, cudaMemcpyHostToDevice, st));
    const float K4[4] = {geom->fx, geom->fy, geom->cx, geom->cy};
    int rc = msl_glue_keypoints_dev(glue, (const msl_keypoint *)(b + oK), rows, (const int32_t *)(b + oC), n_frames, K4, nullptr,
                                    (const float *)(b + oZ), geom->mbf, (float *)(b + oXY), (float *)(b + oUR), (float *)(b + oKD), st);
    if (rc) return rc;
    rc = msl_search_by_projection_frames_dev(m, geom, th, check_orientation, th_depth, (const msl_keypoint *)(b + oK), b + oD, rows,
                                             (const int32_t *)(b + oC), n_frames, (const float *)(b + oXY), (const float *)(b + oUR),
                                             (const float *)(b + oKD), Tcw, (int32_t *)(b + oCM), (int32_t *)(b + oNM), st);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(cur_match, b + oCM, (F - 1) * R * 4, cudaMemcpyDeviceToHost, st));
    MSL_CUDA(cudaMemcpyAsync(nmatches, b + oNM, (F - 1) * 4, cudaMemcpyDeviceToHost, st));
    MSL_CUDA(cudaStreamSynchronize(st));
    return MSL_OK;
}

int msl_search_by_projection_points(msl_matcher *m, const msl_frame_geom *geom, float th, float nnratio, int n_mp,
                                    const uint8_t *mp_valid, const uint8_t *mp_obs, const float *mp_proj_xyr,
                                    const int32_t *mp_level, const float *mp_viewcos, const uint8_t *mp_desc, int n_cur,
                                    const float *cur_xy, const int32_t *cur_octave, const float *cur_uright,
                                    const uint8_t *cur_desc, const uint8_t *cur_occupied, int32_t *cur_match,
                                    int32_t *nmatches) {
    if (!m || !geom || !cur_match || !nmatches) return fail(MSL_ERR_INVALID, "msl_search_by_projection_points: null argument");
    if (n_mp < 0 || n_cur < 0 || n_cur > MAXK || n_mp > 0x7ffffff)
        return fail(MSL_ERR_INVALID, "msl_search_by_projection_points: too many keypoints for this handle");
    if (n_cur == 0 || n_mp == 0) {
        for (int j = 0; j < n_cur; j++) cur_match[j] = cur_occupied[j] ? -2 : -1;
        *nmatches = 0;
        return MSL_OK;
    }
    MSL_CUDA(cudaSetDevice(m->device));
    SearchArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *geom, A.mode = 1, A.th = th, A.nnratio = nnratio, A.checkOri = 0, A.nq = n_mp, A.nc = n_cur;
    A.distTh = TH_HIGH;
    {
        const int rc_ = ensure_scratch(m, (size_t)(n_mp), (size_t)(n_cur));
        if (rc_) return rc_;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    A.q_valid = ar.put(mp_valid, n_mp);
    A.q_obs = ar.put(mp_obs, n_mp);
    A.q_desc = ar.put(mp_desc, (size_t)n_mp * 32);
    A.q_f3 = ar.put(mp_proj_xyr, (size_t)n_mp * 3);
    A.q_level = ar.put(mp_level, n_mp);
    A.q_aux = ar.put(mp_viewcos, n_mp);
    A.c_xy = ar.put(cur_xy, (size_t)n_cur * 2);
    A.c_angle = A.c_xy;  // unused in this mode
    A.c_uright = ar.put(cur_uright, n_cur);
    A.c_octave = ar.put(cur_octave, n_cur);
    A.c_desc = ar.put(cur_desc, (size_t)n_cur * 32);
    A.c_occ = ar.put(cur_occupied, n_cur);
    A.c_match = ar.get<int32_t>(n_cur);
    A.nmatches = ar.get<int32_t>(1);
    A.assign = ar.get<int32_t>(n_mp);
    A.bin = ar.get<int32_t>(n_mp);
    if (!ar.ok) return fail(MSL_ERR_CUDA, "msl_search_by_projection_points: scratch arena / copy failure");
    return run_search(m, ar, A, cur_match, nmatches);
}

int msl_search_by_projection_keyframe(msl_matcher *m, const msl_frame_geom *geom, const float Tcw_cur[16], float th,
                                      int orb_dist, int check_orientation, float log_scale_factor, int n_kf,
                                      const uint8_t *kf_valid, const float *kf_mp_world, const uint8_t *kf_mp_desc,
                                      const float *kf_mp_dist, const float *kf_angle, int n_cur, const float *cur_xy,
                                      const int32_t *cur_octave, const float *cur_angle, const uint8_t *cur_desc,
                                      const uint8_t *cur_occupied, int32_t *cur_match, int32_t *nmatches) {
    if (!m || !geom || !Tcw_cur || !cur_match || !nmatches) return fail(MSL_ERR_INVALID, "msl_search_by_projection_keyframe: null argument");
    if (n_kf < 0 || n_cur < 0 || n_cur > MAXK || n_kf > 0x7ffffff)
        return fail(MSL_ERR_INVALID, "msl_search_by_projection_keyframe: too many keypoints for this handle");
    if (geom->nlevels < 1 || geom->nlevels > 16 || !(log_scale_factor > 0.f) || orb_dist < 0 || orb_dist > 255)
        return fail(MSL_ERR_INVALID, "msl_search_by_projection_keyframe: parameter out of range");
    if (n_cur == 0 || n_kf == 0) {
        for (int j = 0; j < n_cur; j++) cur_match[j] = cur_occupied[j] ? -2 : -1;
        *nmatches = 0;
        return MSL_OK;
    }
    MSL_CUDA(cudaSetDevice(m->device));
    SearchArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *geom, A.mode = 2, A.th = th, A.nnratio = 0, A.checkOri = check_orientation, A.nq = n_kf, A.nc = n_cur;
    A.distTh = orb_dist, A.logScale = log_scale_factor;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) A.Rcw[r * 3 + c] = Tcw_cur[r * 4 + c];
        A.tcw[r] = Tcw_cur[r * 4 + 3];
    }
    for (int r = 0; r < 3; r++) {  // :686 Ow = -Rcw.t() * tcw (transposed operand: general path, double accumulation, one rounding)
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)(-A.Rcw[k * 3 + r]) * (double)A.tcw[k];
        A.Ow[r] = (float)s;
    }
    std::vector<uint8_t> ones(n_kf, 1);  // every slot assigned by this overload blocks later queries (:741-742)
    {
        const int rc_ = ensure_scratch(m, (size_t)(n_kf), (size_t)(n_cur));
        if (rc_) return rc_;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    A.q_valid = ar.put(kf_valid, n_kf);
    A.q_obs = ar.put(ones.data(), n_kf);
    A.q_desc = ar.put(kf_mp_desc, (size_t)n_kf * 32);
    A.q_f3 = ar.put(kf_mp_world, (size_t)n_kf * 3);
    A.q_f2 = ar.put(kf_mp_dist, (size_t)n_kf * 2);
    A.q_level = nullptr;
    A.q_aux = ar.put(kf_angle, n_kf);
    A.c_xy = ar.put(cur_xy, (size_t)n_cur * 2);
    A.c_angle = ar.put(cur_angle, n_cur);
    A.c_uright = A.c_angle;  // unused in this mode
    A.c_octave = ar.put(cur_octave, n_cur);
    A.c_desc = ar.put(cur_desc, (size_t)n_cur * 32);
    A.c_occ = ar.put(cur_occupied, n_cur);
    A.c_match = ar.get<int32_t>(n_cur);
    A.nmatches = ar.get<int32_t>(1);
    A.assign = ar.get<int32_t>(n_kf);
    A.bin = ar.get<int32_t>(n_kf);
    if (!ar.ok) return fail(MSL_ERR_CUDA, "msl_search_by_projection_keyframe: scratch arena / copy failure");
    return run_search(m, ar, A, cur_match, nmatches);
}

static bool csr_ok(int nNodes, const uint32_t *id, const int32_t *off, const int32_t *feat, int nFeat) {
    if (nNodes < 0) return false;
    if (nNodes == 0) return true;
    if (!id || !off || off[0] != 0) return false;
    for (int k = 0; k < nNodes; k++) {
        if (off[k + 1] < off[k]) return false;
        if (k && !(id[k - 1] < id[k])) return false;  // std::map order
    }
    if (off[nNodes] && !feat) return false;
    for (int e = 0; e < off[nNodes]; e++)
        if (feat[e] < 0 || feat[e] >= nFeat) return false;
    return true;
}

// out, bin and nmatches are adjacent in the arena (allocated in that order by both callers)
static int run_node_search(msl_matcher *m, Arena &ar, NodeArgs &A, int nOut, int32_t *out, int32_t *nmatches) {
    m->bNode.push_back(A);
    const int rc = matcher_record(m, ar, {out, A.out, sizeof(int32_t) * nOut}, {nullptr, A.bin, 0}, {nmatches, A.nmatches, sizeof(int32_t)});
    if (rc) m->bSearch.clear(), m->bNode.clear(), m->bFuse.clear(), m->pending.clear(), m->batchOff = 0;
    return rc;
}

int msl_search_by_bow(msl_matcher *m, float nnratio, int check_orientation, int n_nodes_kf, const uint32_t *kf_node_id,
                      const int32_t *kf_node_off, const int32_t *kf_node_feat, int n_nodes_f, const uint32_t *f_node_id,
                      const int32_t *f_node_off, const int32_t *f_node_feat, int n_kf, const uint8_t *kf_valid,
                      const uint8_t *kf_desc, const float *kf_angle, int n_f, const uint8_t *f_desc, const float *f_angle,
                      int32_t *f_match, int32_t *nmatches) {
    if (!m || !f_match || !nmatches) return fail(MSL_ERR_INVALID, "msl_search_by_bow: null argument");
    if (n_kf < 0 || n_f < 0 || n_kf > 0x7ffffff || n_f > 0x7ffffff) return fail(MSL_ERR_INVALID, "msl_search_by_bow: too many keypoints for this handle");
    if (!csr_ok(n_nodes_kf, kf_node_id, kf_node_off, kf_node_feat, n_kf) || !csr_ok(n_nodes_f, f_node_id, f_node_off, f_node_feat, n_f))
        return fail(MSL_ERR_INVALID, "msl_search_by_bow: malformed feature vector (ids must ascend, offsets must be monotone, indices in range)");
    if (n_kf == 0 || n_f == 0 || n_nodes_kf == 0 || n_nodes_f == 0) {
        for (int j = 0; j < n_f; j++) f_match[j] = -1;
        *nmatches = 0;
        return MSL_OK;
    }
    if (!kf_valid || !kf_desc || !kf_angle || !f_desc || !f_angle) return fail(MSL_ERR_INVALID, "msl_search_by_bow: null argument");
    MSL_CUDA(cudaSetDevice(m->device));
    NodeArgs A;
    memset(&A, 0, sizeof(A));
    A.mode = 0, A.nnratio = nnratio, A.checkOri = check_orientation, A.nNodesA = n_nodes_kf, A.nNodesB = n_nodes_f;
    A.nA = n_kf, A.nB = n_f;
    {
        const int rc_ = ensure_scratch(m, (size_t)(n_kf + n_nodes_kf + n_nodes_f), (size_t)(n_f));
        if (rc_) return rc_;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    A.idA = ar.put(kf_node_id, n_nodes_kf), A.offA = ar.put(kf_node_off, n_nodes_kf + 1);
    A.featA = ar.put(kf_node_feat, kf_node_off[n_nodes_kf]);
    A.idB = ar.put(f_node_id, n_nodes_f), A.offB = ar.put(f_node_off, n_nodes_f + 1);
    A.featB = ar.put(f_node_feat, f_node_off[n_nodes_f]);
    A.a_flag = ar.put(kf_valid, n_kf), A.a_desc = ar.put(kf_desc, (size_t)n_kf * 32), A.a_angle = ar.put(kf_angle, n_kf);
    A.b_desc = ar.put(f_desc, (size_t)n_f * 32), A.b_angle = ar.put(f_angle, n_f);
    A.out = ar.get<int32_t>(n_f), A.bin = ar.get<int32_t>(n_f), A.nmatches = ar.get<int32_t>(1);
    if (!ar.ok) return fail(MSL_ERR_CUDA, "msl_search_by_bow: scratch arena / copy failure");
    return run_node_search(m, ar, A, n_f, f_match, nmatches);
}

int msl_search_for_triangulation(msl_matcher *m, const float F12[9], const float Cw1[3], const float Tcw2[16],
                                 const float K2[4], int only_stereo, int check_orientation, int nlevels,
                                 const float *scale_factors2, const float *level_sigma2_2, int n_nodes1,
                                 const uint32_t *node_id1, const int32_t *node_off1, const int32_t *node_feat1,
                                 int n_nodes2, const uint32_t *node_id2, const int32_t *node_off2,
                                 const int32_t *node_feat2, int n1, const uint8_t *has_mp1, const float *uright1,
                                 const float *xy1, const float *angle1, const uint8_t *desc1, int n2,
                                 const uint8_t *has_mp2, const float *uright2, const float *xy2, const int32_t *octave2,
                                 const float *angle2, const uint8_t *desc2, int32_t *matches12, int32_t *nmatches) {
    if (!m || !F12 || !Cw1 || !Tcw2 || !K2 || !scale_factors2 || !level_sigma2_2 || !matches12 || !nmatches)
        return fail(MSL_ERR_INVALID, "msl_search_for_triangulation: null argument");
    if (n1 < 0 || n2 < 0 || n1 > 0x7ffffff || n2 > 0x7ffffff) return fail(MSL_ERR_INVALID, "msl_search_for_triangulation: too many keypoints for this handle");
    if (nlevels < 1 || nlevels > 16) return fail(MSL_ERR_INVALID, "msl_search_for_triangulation: nlevels out of range");
    if (!csr_ok(n_nodes1, node_id1, node_off1, node_feat1, n1) || !csr_ok(n_nodes2, node_id2, node_off2, node_feat2, n2))
        return fail(MSL_ERR_INVALID, "msl_search_for_triangulation: malformed feature vector");
    if (n1 == 0 || n2 == 0 || n_nodes1 == 0 || n_nodes2 == 0) {
        for (int i = 0; i < n1; i++) matches12[i] = -1;
        *nmatches = 0;
        return MSL_OK;
    }
    if (!has_mp1 || !uright1 || !xy1 || !angle1 || !desc1 || !has_mp2 || !uright2 || !xy2 || !octave2 || !angle2 || !desc2)
        return fail(MSL_ERR_INVALID, "msl_search_for_triangulation: null argument");
    for (int j = 0; j < n2; j++)
        if (octave2[j] < 0 || octave2[j] >= nlevels) return fail(MSL_ERR_INVALID, "msl_search_for_triangulation: octave out of range");
    MSL_CUDA(cudaSetDevice(m->device));
    NodeArgs A;
    memset(&A, 0, sizeof(A));
    A.mode = 1, A.checkOri = check_orientation, A.onlyStereo = only_stereo, A.nNodesA = n_nodes1, A.nNodesB = n_nodes2;
    A.nA = n1, A.nB = n2;
    memcpy(A.F12, F12, sizeof(float) * 9);
    for (int l = 0; l < nlevels; l++) A.scale2[l] = scale_factors2[l], A.sigma2[l] = level_sigma2_2[l];
    {   // :263-270 epipole in the second image: C2 = R2w * Cw + t2w (one cv gemm, small-matrix path)
        float C2[3];
        for (int r = 0; r < 3; r++) C2[r] = gemm_row(Tcw2 + 4 * r, Cw1, Tcw2[4 * r + 3]);
        const float invz = 1.0f / C2[2];
        A.ex = K2[0] * C2[0] * invz + K2[2];
        A.ey = K2[1] * C2[1] * invz + K2[3];
    }
    {
        const int rc_ = ensure_scratch(m, (size_t)(n1 + n_nodes1 + n_nodes2), (size_t)(n2));
        if (rc_) return rc_;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    A.idA = ar.put(node_id1, n_nodes1), A.offA = ar.put(node_off1, n_nodes1 + 1), A.featA = ar.put(node_feat1, node_off1[n_nodes1]);
    A.idB = ar.put(node_id2, n_nodes2), A.offB = ar.put(node_off2, n_nodes2 + 1), A.featB = ar.put(node_feat2, node_off2[n_nodes2]);
    A.a_flag = ar.put(has_mp1, n1), A.a_desc = ar.put(desc1, (size_t)n1 * 32), A.a_angle = ar.put(angle1, n1);
    A.a_xy = ar.put(xy1, (size_t)n1 * 2), A.a_uright = ar.put(uright1, n1);
    A.b_flag = ar.put(has_mp2, n2), A.b_desc = ar.put(desc2, (size_t)n2 * 32), A.b_angle = ar.put(angle2, n2);
    A.b_xy = ar.put(xy2, (size_t)n2 * 2), A.b_uright = ar.put(uright2, n2), A.b_octave = ar.put(octave2, n2);
    A.out = ar.get<int32_t>(n1), A.bin = ar.get<int32_t>(n1), A.nmatches = ar.get<int32_t>(1);
    if (!ar.ok) return fail(MSL_ERR_CUDA, "msl_search_for_triangulation: scratch arena / copy failure");
    return run_node_search(m, ar, A, n1, matches12, nmatches);
}

int msl_fuse_search(msl_matcher *m, const msl_frame_geom *geom, const float Tcw[16], float th, float log_scale_factor,
                    const float *inv_level_sigma2, int n_mp, const uint8_t *mp_valid, const float *mp_world,
                    const float *mp_normal, const float *mp_dist, const uint8_t *mp_desc, int n_kf, const float *kf_xy,
                    const int32_t *kf_octave, const float *kf_uright, const uint8_t *kf_desc, int32_t *best_idx,
                    int32_t *best_dist, int32_t *nfused) {
    if (!m || !geom || !Tcw || !inv_level_sigma2 || !best_idx || !best_dist || !nfused) return fail(MSL_ERR_INVALID, "msl_fuse_search: null argument");
    if (n_mp < 0 || n_kf < 0 || n_kf > MAXK || n_mp > 0x7ffffff) return fail(MSL_ERR_INVALID, "msl_fuse_search: too many keypoints for this handle");
    if (geom->nlevels < 1 || geom->nlevels > 16 || !(log_scale_factor > 0.f)) return fail(MSL_ERR_INVALID, "msl_fuse_search: parameter out of range");
    if (n_mp == 0 || n_kf == 0) {
        for (int i = 0; i < n_mp; i++) best_idx[i] = -1, best_dist[i] = 256;
        *nfused = 0;
        return MSL_OK;
    }
    if (!mp_valid || !mp_world || !mp_normal || !mp_dist || !mp_desc || !kf_xy || !kf_octave || !kf_uright || !kf_desc)
        return fail(MSL_ERR_INVALID, "msl_fuse_search: null argument");
    for (int j = 0; j < n_kf; j++)
        if (kf_octave[j] < 0 || kf_octave[j] >= geom->nlevels) return fail(MSL_ERR_INVALID, "msl_fuse_search: octave out of range");
    MSL_CUDA(cudaSetDevice(m->device));
    FuseArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *geom, A.th = th, A.logScale = log_scale_factor, A.nq = n_mp, A.nc = n_kf;
    for (int l = 0; l < geom->nlevels; l++) A.invSigma2[l] = inv_level_sigma2[l];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) A.Rcw[r * 3 + c] = Tcw[r * 4 + c];
        A.tcw[r] = Tcw[r * 4 + 3];
    }
    for (int r = 0; r < 3; r++) {  // KeyFrame::SetPose (src/KeyFrame.cc:79-80): Rwc = Rcw.t(); Ow = -Rwc * tcw -- flags == 0, alpha = -1
        const float t0 = A.Rcw[0 * 3 + r] * A.tcw[0] + A.Rcw[1 * 3 + r] * A.tcw[1] + A.Rcw[2 * 3 + r] * A.tcw[2];
        A.Ow[r] = (float)((double)t0 * -1.0);
    }
    {
        const int rc_ = ensure_scratch(m, (size_t)(n_mp), (size_t)(n_kf));
        if (rc_) return rc_;
    }
    Arena ar{m->d_scr, m->batchOff, m->scrCap, m->stream, m->h_scr};
    A.q_valid = ar.put(mp_valid, n_mp), A.q_desc = ar.put(mp_desc, (size_t)n_mp * 32);
    A.q_world = ar.put(mp_world, (size_t)n_mp * 3), A.q_normal = ar.put(mp_normal, (size_t)n_mp * 3);
    A.q_dist = ar.put(mp_dist, (size_t)n_mp * 2);
    A.c_xy = ar.put(kf_xy, (size_t)n_kf * 2), A.c_uright = ar.put(kf_uright, n_kf), A.c_octave = ar.put(kf_octave, n_kf);
    A.c_desc = ar.put(kf_desc, (size_t)n_kf * 32);
    A.bestIdx = ar.get<int32_t>(n_mp), A.bestDist = ar.get<int32_t>(n_mp), A.nFused = ar.get<int32_t>(1);
    if (!ar.ok) return fail(MSL_ERR_CUDA, "msl_fuse_search: scratch arena / copy failure");
    m->bFuse.push_back(A);
    const int rc = matcher_record(m, ar, {best_idx, A.bestIdx, sizeof(int32_t) * n_mp}, {best_dist, A.bestDist, sizeof(int32_t) * n_mp},
                                  {nfused, A.nFused, sizeof(int32_t)});
    if (rc) m->bSearch.clear(), m->bNode.clear(), m->bFuse.clear(), m->pending.clear(), m->batchOff = 0;
    return rc;
}

int msl_distinctive_descriptors(msl_matcher *m, int n_points, const int32_t *offsets, const uint8_t *desc, int32_t *best_idx,
                                int32_t *best_median) {
    if (!m || n_points < 0 || (n_points > 0 && (!offsets || !best_idx))) return fail(MSL_ERR_INVALID, "msl_distinctive_descriptors: bad argument");
    if (n_points == 0) return MSL_OK;
    if (offsets[0] != 0) return fail(MSL_ERR_INVALID, "msl_distinctive_descriptors: offsets must start at 0");
    for (int k = 0; k < n_points; k++)
        if (offsets[k + 1] < offsets[k]) return fail(MSL_ERR_INVALID, "msl_distinctive_descriptors: offsets must be monotone");
    const size_t total = (size_t)offsets[n_points];
    if (total && !desc) return fail(MSL_ERR_INVALID, "msl_distinctive_descriptors: null descriptors");
    MSL_CUDA(cudaSetDevice(m->device));
    const size_t offBytes = align_up(sizeof(int32_t) * ((size_t)n_points + 1), 256), descBytes = align_up(total * 32 + 32, 256);
    const size_t outBytes = align_up(sizeof(int32_t) * (size_t)n_points, 256);
    const size_t need = offBytes + descBytes + 2 * outBytes;
    if (need > m->ddCap) {
        MSL_CUDA(cudaStreamSynchronize(m->stream));
        if (m->d_dd) cudaFree(m->d_dd);
        m->d_dd = nullptr, m->ddCap = 0;
        MSL_CUDA(cudaMalloc((void **)&m->d_dd, need));
        m->ddCap = need;
    }
    int32_t *d_off = (int32_t *)m->d_dd;
    uint8_t *d_desc = m->d_dd + offBytes;
    int32_t *d_bi = (int32_t *)(m->d_dd + offBytes + descBytes), *d_bm = (int32_t *)(m->d_dd + offBytes + descBytes + outBytes);
    MSL_CUDA(cudaMemcpyAsync(d_off, offsets, sizeof(int32_t) * ((size_t)n_points + 1), cudaMemcpyHostToDevice, m->stream));
    if (total) MSL_CUDA(cudaMemcpyAsync(d_desc, desc, total * 32, cudaMemcpyHostToDevice, m->stream));
    k_distinctive<<<n_points, 128, 0, m->stream>>>(d_off, d_desc, d_bi, d_bm);
    MSL_LAUNCH_CHECK();
    MSL_CUDA(cudaMemcpyAsync(best_idx, d_bi, sizeof(int32_t) * (size_t)n_points, cudaMemcpyDeviceToHost, m->stream));
    if (best_median) MSL_CUDA(cudaMemcpyAsync(best_median, d_bm, sizeof(int32_t) * (size_t)n_points, cudaMemcpyDeviceToHost, m->stream));
    MSL_CUDA(cudaStreamSynchronize(m->stream));
    return MSL_OK;
}

}  // extern "C"
