// orb_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, see msl_oracle.h) restating
// src/ORBextractor.cc of razayunus/ManhattanSLAM together with the OpenCV
// primitives it calls (cv::resize INTER_LINEAR 8U, cv::FAST TYPE_9_16 + NMS,
// cv::GaussianBlur 7x7 sigma 2 on 8U, cv::fastAtan2, cvRound).  The primitives are
// pinned bit-exactly against cv2 4.13.0 by tests/test_oracle_primitives.py.
//
// Build: g++ -O2 -ffp-contract=off (the reference is built "-Wall -O3" without
// -march=native, CMakeLists.txt:10-11, so no FMA contraction happens there either).
#include "msl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <stdexcept>
#include <utility>
#include <vector>

namespace {

static const int8_t kPattern[1024] = {
#include "rbrief_pattern_31.inc"
};

const int PATCH_SIZE = 31;       // src/ORBextractor.cc:70
const int HALF_PATCH_SIZE = 15;  // :71
const int EDGE_THRESHOLD = 19;   // :72

// cvRound(double/float): SSE2 cvtsd2si / cvtss2si under the default rounding mode = round-half-even.
inline int cv_round(double v) { return (int)std::nearbyint(v); }
inline int cv_round(float v) { return (int)std::nearbyintf(v); }
inline int cv_floor(double v) {
    int i = (int)v;
    return i - (i > v);
}
inline int cv_ceil(double v) {
    int i = (int)v;
    return i + (i < v);
}
inline short sat_short(float v) {
    int iv = cv_round(v);
    return (short)(iv < -32768 ? -32768 : iv > 32767 ? 32767 : iv);
}

// ----------------------------------------------------------------------------
// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 (OpenCV imgproc/resize.cpp:
// resizeGeneric_ with HResizeLinear<uchar,int,short,2048> and
// VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>,VResizeLinearVec_32s8u>).
// The IPP path is not taken for 8u linear ("Resize which doesn't match OpenCV exactly").
// Called at src/ORBextractor.cc:882.
void resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride) {
    const int COEF_BITS = 11, COEF_SCALE = 1 << COEF_BITS;
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(dw * 2), ibeta(dh * 2);
    int xmax = dw;
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) {
            fx = 0;
            sx = 0;
        }
        if (sx + 1 >= sw) {
            xmax = std::min(xmax, dx);
            if (sx >= sw - 1) {
                fx = 0;
                sx = sw - 1;
            }
        }
        xofs[dx] = sx;
        ialpha[dx * 2] = sat_short((1.f - fx) * COEF_SCALE);
        ialpha[dx * 2 + 1] = sat_short(fx * COEF_SCALE);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[dy * 2] = sat_short((1.f - fy) * COEF_SCALE);
        ibeta[dy * 2 + 1] = sat_short(fy * COEF_SCALE);
    }
    std::vector<int> row0(dw), row1(dw);
    auto hresize = [&](int sy, std::vector<int> &D) {
        const uint8_t *S = src + (size_t)sy * sstride;
        int dx = 0;
        for (; dx < xmax; dx++) {
            int sx = xofs[dx];
            D[dx] = S[sx] * ialpha[dx * 2] + S[sx + 1] * ialpha[dx * 2 + 1];
        }
        for (; dx < dw; dx++) D[dx] = S[xofs[dx]] * COEF_SCALE;
    };
    auto clip = [](int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; };
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = clip(yofs[dy], 0, sh), sy1 = clip(yofs[dy] + 1, 0, sh);
        hresize(sy0, row0);
        hresize(sy1, row1);
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        uint8_t *D = dst + (size_t)dy * dstride;
        for (int x = 0; x < dw; x++)
            D[x] = (uint8_t)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ----------------------------------------------------------------------------
// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1, OpenCV >= 3.4 / 4.x
// bit-exact fixed-point path (smooth.dispatch.cpp GaussianBlurFixedPoint, ufixedpoint16 8.8 taps from
// getGaussianKernelFixedPoint_ED: error-diffused rounding of getGaussianKernelBitExact(7,2)*256 with
// the centre tap taking the remainder so the taps sum to 256).  Called at src/ORBextractor.cc:852.
const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};

inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0)
            p = -p;
        else
            p = 2 * (len - 1) - p;
    }
    return p;
}

void gaussian_blur_7x7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *S = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            unsigned acc = 0;
            for (int k = -3; k <= 3; k++) acc += (unsigned)kGauss7[k + 3] * S[reflect101(x + k, w)];
            tmp[(size_t)y * w + x] = (uint16_t)acc;  // <= 255*256, no saturation possible
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = -3; k <= 3; k++) acc += (uint32_t)kGauss7[k + 3] * tmp[(size_t)reflect101(y + k, h) * w + x];
            D[x] = (uint8_t)((acc + (1u << 15)) >> 16);
        }
    }
}

// ----------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmax=true) = FAST_t<16> (features2d/fast.cpp) with
// cornerScore<16> (fast_score.cpp).  Called at src/ORBextractor.cc:763-768.
static const int kRing[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

inline void make_offsets(int pixel[25], int stride) {
    int k = 0;
    for (; k < 16; k++) pixel[k] = kRing[k][0] + kRing[k][1] * stride;
    for (; k < 25; k++) pixel[k] = pixel[k - 16];
}

// fast_score.cpp cornerScore<16>: largest threshold for which the pixel is still a corner.
inline int corner_score_16(const uint8_t *ptr, const int pixel[25], int threshold) {
    const int K = 8, N = K * 3 + 1;
    int k, v = ptr[0];
    short d[N];
    for (k = 0; k < N; k++) d[k] = (short)(v - ptr[pixel[k]]);
    int a0 = threshold;
    for (k = 0; k < 16; k += 2) {
        int a = std::min((int)d[k + 1], (int)d[k + 2]);
        a = std::min(a, (int)d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, (int)d[k + 4]);
        a = std::min(a, (int)d[k + 5]);
        a = std::min(a, (int)d[k + 6]);
        a = std::min(a, (int)d[k + 7]);
        a = std::min(a, (int)d[k + 8]);
        a0 = std::max(a0, std::min(a, (int)d[k]));
        a0 = std::max(a0, std::min(a, (int)d[k + 9]));
    }
    int b0 = -a0;
    for (k = 0; k < 16; k += 2) {
        int b = std::max((int)d[k + 1], (int)d[k + 2]);
        b = std::max(b, (int)d[k + 3]);
        b = std::max(b, (int)d[k + 4]);
        b = std::max(b, (int)d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, (int)d[k + 6]);
        b = std::max(b, (int)d[k + 7]);
        b = std::max(b, (int)d[k + 8]);
        b0 = std::min(b0, std::max(b, (int)d[k]));
        b0 = std::min(b0, std::max(b, (int)d[k + 9]));
    }
    return -b0 - 1;
}

// 9 contiguous ring pixels all > v+t or all < v-t (FAST_t<16> segment test, K = 8 => count > 8).
inline bool is_corner_9_16(const uint8_t *ptr, const int pixel[25], int t) {
    int v = ptr[0];
    int cb = 0, cd = 0;
    for (int k = 0; k < 25; k++) {
        int x = ptr[pixel[k]];
        if (x > v + t) {
            if (++cb > 8) return true;
        } else
            cb = 0;
        if (x < v - t) {
            if (++cd > 8) return true;
        } else
            cd = 0;
    }
    return false;
}

struct XYR {
    int x, y, r;
};

void fast_9_16(const uint8_t *img, int w, int h, int stride, int threshold, bool nms, std::vector<XYR> &out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int pixel[25];
    make_offsets(pixel, stride);
    threshold = std::min(std::max(threshold, 0), 255);
    std::vector<uint8_t> bufmem((size_t)w * 3, 0);
    uint8_t *buf[3] = {bufmem.data(), bufmem.data() + w, bufmem.data() + 2 * w};
    std::vector<int> cpmem((size_t)(w + 1) * 3, 0);
    int *cpbuf[3] = {cpmem.data(), cpmem.data() + (w + 1), cpmem.data() + 2 * (w + 1)};
    for (int i = 3; i < h - 2; i++) {
        const uint8_t *ptr = img + (size_t)i * stride + 3;
        uint8_t *curr = buf[(i - 3) % 3];
        int *cornerpos = cpbuf[(i - 3) % 3] + 1;
        std::memset(curr, 0, w);
        int ncorners = 0;
        if (i < h - 3) {
            for (int j = 3; j < w - 3; j++, ptr++) {
                if (is_corner_9_16(ptr, pixel, threshold)) {
                    cornerpos[ncorners++] = j;
                    if (nms) curr[j] = (uint8_t)corner_score_16(ptr, pixel, threshold);
                }
            }
        }
        cornerpos[-1] = ncorners;
        if (i == 3) continue;
        const uint8_t *prev = buf[(i - 4 + 3) % 3];
        const uint8_t *pprev = buf[(i - 5 + 3) % 3];
        cornerpos = cpbuf[(i - 4 + 3) % 3] + 1;
        ncorners = cornerpos[-1];
        for (int k = 0; k < ncorners; k++) {
            int j = cornerpos[k];
            int score = prev[j];
            if (!nms || (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] && score > pprev[j] &&
                         score > pprev[j + 1] && score > curr[j - 1] && score > curr[j] && score > curr[j + 1])) {
                out.push_back({j, i - 1, score});
            }
        }
    }
}

// Threshold-free restatement: S_max(p) = max over the 16 nine-pixel arcs and both polarities of
// min |I_k - I_c| with a consistent sign; cornerScore == max(t, S_max) - 1 and corner(t) <=> S_max > t.
inline int smax_9_16(const uint8_t *ptr, const int pixel[25]) {
    int v = ptr[0];
    int d[25];
    for (int k = 0; k < 25; k++) d[k] = v - ptr[pixel[k]];
    int best = 0;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) {
            mn = std::min(mn, d[s + k]);
            mx = std::max(mx, d[s + k]);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

// cv::fastAtan2 scalar (core/mathfuncs_core: atan_f32), degrees.
float fast_atan2(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    float ax = std::abs(x), ay = std::abs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    const uint8_t *at(int y, int x) const { return d.data() + (size_t)y * w + x; }
};

// ExtractorNode, include/ORBextractor.h:30-41 (UL/UR/BL/BR are always an axis-aligned rectangle).
struct Node {
    std::vector<XYR> vKeys;  // x,y relative to minBorder; r = response
    std::vector<int> vIdx;   // original candidate indices (diagnostics only)
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool bNoMore = false;
    long seq = 0;  // creation sequence number: the oracle's stand-in for the heap address used as
                   // the tie-break of sort(pair<int,ExtractorNode*>) at src/ORBextractor.cc:654
};

}  // namespace

struct orc_orb {
    int nfeatures;
    double scaleFactor;  // include/ORBextractor.h:97 (double member initialised from a float)
    int nlevels, iniThFAST, minThFAST;
    std::vector<int> mnFeaturesPerLevel, umax;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<Image> pyramid, blurred;
    std::vector<std::vector<XYR>> candidates;
    std::vector<std::vector<orc_keypoint>> levelKps;
    long seqCounter = 0;

    // src/ORBextractor.cc:477-529
    void DivideNode(Node &p, Node &n1, Node &n2, Node &n3, Node &n4) {
        const int halfX = (int)std::ceil(static_cast<float>(p.URx - p.ULx) / 2);
        const int halfY = (int)std::ceil(static_cast<float>(p.BRy - p.ULy) / 2);
        n1.ULx = p.ULx, n1.ULy = p.ULy;
        n1.URx = p.ULx + halfX, n1.URy = p.ULy;
        n1.BLx = p.ULx, n1.BLy = p.ULy + halfY;
        n1.BRx = p.ULx + halfX, n1.BRy = p.ULy + halfY;
        n2.ULx = n1.URx, n2.ULy = n1.URy;
        n2.URx = p.URx, n2.URy = p.URy;
        n2.BLx = n1.BRx, n2.BLy = n1.BRy;
        n2.BRx = p.URx, n2.BRy = p.ULy + halfY;
        n3.ULx = n1.BLx, n3.ULy = n1.BLy;
        n3.URx = n1.BRx, n3.URy = n1.BRy;
        n3.BLx = p.BLx, n3.BLy = p.BLy;
        n3.BRx = n1.BRx, n3.BRy = p.BLy;
        n4.ULx = n3.URx, n4.ULy = n3.URy;
        n4.URx = n2.BRx, n4.URy = n2.BRy;
        n4.BLx = n3.BRx, n4.BLy = n3.BRy;
        n4.BRx = p.BRx, n4.BRy = p.BRy;
        for (size_t i = 0; i < p.vKeys.size(); i++) {
            const XYR &kp = p.vKeys[i];
            if ((float)kp.x < n1.URx) {
                if ((float)kp.y < n1.BRy)
                    n1.vKeys.push_back(kp);
                else
                    n3.vKeys.push_back(kp);
            } else if ((float)kp.y < n1.BRy)
                n2.vKeys.push_back(kp);
            else
                n4.vKeys.push_back(kp);
        }
        if (n1.vKeys.size() == 1) n1.bNoMore = true;
        if (n2.vKeys.size() == 1) n2.bNoMore = true;
        if (n3.vKeys.size() == 1) n3.bNoMore = true;
        if (n4.vKeys.size() == 1) n4.bNoMore = true;
    }

    // src/ORBextractor.cc:531-721
    std::vector<XYR> DistributeOctTree(const std::vector<XYR> &vToDistributeKeys, int minX, int maxX, int minY,
                                       int maxY, int N) {
        const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
        // A level more than twice as tall as wide gives nIni == 0: the reference then divides by zero and indexes an empty
        // vector (undefined behaviour, :535-560).  The oracle reports it instead (orc_orb_extract returns -2); the CUDA path
        // refuses such geometry at msl_orb_create.
        if (nIni < 1) throw std::domain_error("DistributeOctTree: zero root nodes");
        const float hX = static_cast<float>(maxX - minX) / nIni;
        std::list<Node> lNodes;
        std::vector<Node *> vpIniNodes(nIni);
        for (int i = 0; i < nIni; i++) {
            Node ni;
            ni.ULx = (int)(hX * static_cast<float>(i)), ni.ULy = 0;
            ni.URx = (int)(hX * static_cast<float>(i + 1)), ni.URy = 0;
            ni.BLx = ni.ULx, ni.BLy = maxY - minY;
            ni.BRx = ni.URx, ni.BRy = maxY - minY;
            ni.seq = seqCounter++;
            lNodes.push_back(ni);
            vpIniNodes[i] = &lNodes.back();
        }
        for (size_t i = 0; i < vToDistributeKeys.size(); i++) {
            const XYR &kp = vToDistributeKeys[i];
            vpIniNodes[(int)((float)kp.x / hX)]->vKeys.push_back(kp);
        }
        auto lit = lNodes.begin();
        while (lit != lNodes.end()) {
            if (lit->vKeys.size() == 1) {
                lit->bNoMore = true;
                lit++;
            } else if (lit->vKeys.empty())
                lit = lNodes.erase(lit);
            else
                lit++;
        }
        bool bFinish = false;
        std::vector<std::pair<int, Node *>> vSizeAndPointerToNode;
        auto add_child = [&](Node &n, int *nToExpand) {
            if (n.vKeys.size() > 0) {
                n.seq = seqCounter++;
                lNodes.push_front(n);
                if (n.vKeys.size() > 1) {
                    if (nToExpand) (*nToExpand)++;
                    vSizeAndPointerToNode.push_back(std::make_pair((int)n.vKeys.size(), &lNodes.front()));
                    lNodes.front().lit = lNodes.begin();
                }
            }
        };
        // sort(pair<int,Node*>): first by size, ties by pointer value; the oracle substitutes the creation
        // sequence number for the pointer (monotone allocation model, SURVEY.md section 7 hard part 1).
        auto cmp = [](const std::pair<int, Node *> &a, const std::pair<int, Node *> &b) {
            if (a.first != b.first) return a.first < b.first;
            return a.second->seq < b.second->seq;
        };
        while (!bFinish) {
            int prevSize = (int)lNodes.size();
            lit = lNodes.begin();
            int nToExpand = 0;
            vSizeAndPointerToNode.clear();
            while (lit != lNodes.end()) {
                if (lit->bNoMore) {
                    lit++;
                    continue;
                } else {
                    Node n1, n2, n3, n4;
                    DivideNode(*lit, n1, n2, n3, n4);
                    add_child(n1, &nToExpand);
                    add_child(n2, &nToExpand);
                    add_child(n3, &nToExpand);
                    add_child(n4, &nToExpand);
                    lit = lNodes.erase(lit);
                    continue;
                }
            }
            if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
                bFinish = true;
            } else if (((int)lNodes.size() + nToExpand * 3) > N) {
                while (!bFinish) {
                    prevSize = (int)lNodes.size();
                    std::vector<std::pair<int, Node *>> vPrev = vSizeAndPointerToNode;
                    vSizeAndPointerToNode.clear();
                    std::sort(vPrev.begin(), vPrev.end(), cmp);
                    for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
                        Node n1, n2, n3, n4;
                        DivideNode(*vPrev[j].second, n1, n2, n3, n4);
                        add_child(n1, nullptr);
                        add_child(n2, nullptr);
                        add_child(n3, nullptr);
                        add_child(n4, nullptr);
                        lNodes.erase(vPrev[j].second->lit);
                        if ((int)lNodes.size() >= N) break;
                    }
                    if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
                }
            }
        }
        std::vector<XYR> vResultKeys;
        for (auto it = lNodes.begin(); it != lNodes.end(); it++) {
            std::vector<XYR> &vNodeKeys = it->vKeys;
            const XYR *pKP = &vNodeKeys[0];
            float maxResponse = (float)pKP->r;
            for (size_t k = 1; k < vNodeKeys.size(); k++) {
                if ((float)vNodeKeys[k].r > maxResponse) {
                    pKP = &vNodeKeys[k];
                    maxResponse = (float)vNodeKeys[k].r;
                }
            }
            vResultKeys.push_back(*pKP);
        }
        return vResultKeys;
    }

    // src/ORBextractor.cc:872-893 (the 19-px reflect border is never read downstream: FAST touches
    // x>=16, IC_Angle x>=4, BRIEF works on a clone -- so it is not materialised here)
    void ComputePyramid(const uint8_t *gray, int w, int h, int stride) {
        for (int level = 0; level < nlevels; ++level) {
            float scale = mvInvScaleFactor[level];
            int lw = cv_round((float)w * scale), lh = cv_round((float)h * scale);
            Image &im = pyramid[level];
            im.w = lw, im.h = lh;
            im.d.assign((size_t)lw * lh, 0);
            if (level != 0) {
                const Image &pv = pyramid[level - 1];
                resize_linear_u8(pv.d.data(), pv.w, pv.h, pv.w, im.d.data(), lw, lh, lw);
            } else {
                for (int y = 0; y < h; y++) std::memcpy(im.d.data() + (size_t)y * lw, gray + (size_t)y * stride, w);
            }
        }
    }

    // IC_Angle, src/ORBextractor.cc:75-99
    float IC_Angle(const Image &image, float ptx, float pty) {
        int m_01 = 0, m_10 = 0;
        const int step = image.w;
        const uint8_t *center = image.at(cv_round(pty), cv_round(ptx));
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0;
            int d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        return fast_atan2((float)m_01, (float)m_10);
    }

    // computeOrbDescriptor, src/ORBextractor.cc:104-149
    void computeOrbDescriptor(const orc_keypoint &kpt, const Image &img, uint8_t *desc) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float angle = (float)kpt.angle * factorPI;
        float a = (float)std::cos(angle), b = (float)std::sin(angle);  // float overloads: cosf/sinf
        const uint8_t *center = img.at(cv_round(kpt.y), cv_round(kpt.x));
        const int step = img.w;
        const int8_t *pattern = kPattern;
        auto GET_VALUE = [&](int idx) -> int {
            float px = (float)pattern[idx * 2], py = (float)pattern[idx * 2 + 1];
            return center[cv_round(px * b + py * a) * step + cv_round(px * a - py * b)];
        };
        for (int i = 0; i < 32; ++i, pattern += 32) {
            int val = 0;
            for (int k = 0; k < 8; k++) {
                int t0 = GET_VALUE(2 * k), t1 = GET_VALUE(2 * k + 1);
                val |= (t0 < t1) << k;
            }
            desc[i] = (uint8_t)val;
        }
    }

    // ComputeKeyPointsOctTree, src/ORBextractor.cc:723-803
    void ComputeKeyPointsOctTree() {
        const float W = 30;
        for (int level = 0; level < nlevels; ++level) {
            const Image &im = pyramid[level];
            const int minBorderX = EDGE_THRESHOLD - 3;
            const int minBorderY = minBorderX;
            const int maxBorderX = im.w - EDGE_THRESHOLD + 3;
            const int maxBorderY = im.h - EDGE_THRESHOLD + 3;
            std::vector<XYR> &vToDistributeKeys = candidates[level];
            vToDistributeKeys.clear();
            levelKps[level].clear();
            const float width = (float)(maxBorderX - minBorderX);
            const float height = (float)(maxBorderY - minBorderY);
            const int nCols = (int)(width / W);
            const int nRows = (int)(height / W);
            if (nCols < 1 || nRows < 1) continue;  // reference would divide by zero; levels this small are rejected
            const int wCell = (int)std::ceil(width / nCols);
            const int hCell = (int)std::ceil(height / nRows);
            std::vector<XYR> vKeysCell;
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBorderY - 3) continue;
                if (maxY > maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBorderX - 6) continue;
                    if (maxX > maxBorderX) maxX = (float)maxBorderX;
                    const int x0 = (int)iniX, y0 = (int)iniY, rw = (int)maxX - x0, rh = (int)maxY - y0;
                    fast_9_16(im.at(y0, x0), rw, rh, im.w, iniThFAST, true, vKeysCell);
                    if (vKeysCell.empty()) fast_9_16(im.at(y0, x0), rw, rh, im.w, minThFAST, true, vKeysCell);
                    for (auto &k : vKeysCell) vToDistributeKeys.push_back({k.x + j * wCell, k.y + i * hCell, k.r});
                }
            }
            std::vector<XYR> kept = DistributeOctTree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY,
                                                      maxBorderY, mnFeaturesPerLevel[level]);
            const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
            for (auto &k : kept) {
                orc_keypoint kp;
                kp.x = (float)k.x + minBorderX;
                kp.y = (float)k.y + minBorderY;
                kp.size = (float)scaledPatchSize;
                kp.angle = -1;
                kp.response = (float)k.r;
                kp.octave = level;
                kp.class_id = -1;
                levelKps[level].push_back(kp);
            }
        }
        for (int level = 0; level < nlevels; ++level)
            for (auto &kp : levelKps[level]) kp.angle = IC_Angle(pyramid[level], kp.x, kp.y);
    }
};

extern "C" {

orc_orb *orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    if (nlevels < 1 || nlevels > 16 || nfeatures < 1) return nullptr;
    orc_orb *o = new orc_orb();
    o->nfeatures = nfeatures;
    o->scaleFactor = scaleFactor;
    o->nlevels = nlevels;
    o->iniThFAST = iniThFAST;
    o->minThFAST = minThFAST;
    o->mvScaleFactor.resize(nlevels);
    o->mvLevelSigma2.resize(nlevels);
    o->mvScaleFactor[0] = 1.0f;
    o->mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        o->mvScaleFactor[i] = (float)(o->mvScaleFactor[i - 1] * o->scaleFactor);
        o->mvLevelSigma2[i] = o->mvScaleFactor[i] * o->mvScaleFactor[i];
    }
    o->mvInvScaleFactor.resize(nlevels);
    o->mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        o->mvInvScaleFactor[i] = 1.0f / o->mvScaleFactor[i];
        o->mvInvLevelSigma2[i] = 1.0f / o->mvLevelSigma2[i];
    }
    o->pyramid.resize(nlevels);
    o->blurred.resize(nlevels);
    o->candidates.resize(nlevels);
    o->levelKps.resize(nlevels);
    o->mnFeaturesPerLevel.resize(nlevels);
    float factor = (float)(1.0f / o->scaleFactor);
    float nDesiredFeaturesPerScale =
        nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sumFeatures = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        o->mnFeaturesPerLevel[level] = cv_round(nDesiredFeaturesPerScale);
        sumFeatures += o->mnFeaturesPerLevel[level];
        nDesiredFeaturesPerScale *= factor;
    }
    o->mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);
    // umax, src/ORBextractor.cc:453-467
    o->umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) o->umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
    return o;
}

void orc_orb_destroy(orc_orb *o) { delete o; }
int orc_orb_levels(const orc_orb *o) { return o->nlevels; }

void orc_orb_scale_factors(const orc_orb *o, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2) {
    for (int i = 0; i < o->nlevels; i++) {
        if (scale) scale[i] = o->mvScaleFactor[i];
        if (inv_scale) inv_scale[i] = o->mvInvScaleFactor[i];
        if (sigma2) sigma2[i] = o->mvLevelSigma2[i];
        if (inv_sigma2) inv_sigma2[i] = o->mvInvLevelSigma2[i];
    }
}

void orc_orb_features_per_level(const orc_orb *o, int32_t *n) {
    for (int i = 0; i < o->nlevels; i++) n[i] = o->mnFeaturesPerLevel[i];
}

void orc_orb_umax(const orc_orb *o, int32_t *umax16) {
    for (int i = 0; i < 16; i++) umax16[i] = o->umax[i];
}

int orc_orb_extract(orc_orb *o, const uint8_t *gray, int w, int h, int stride, orc_keypoint *kps, uint8_t *desc,
                    int cap) {
    if (!gray || w <= 0 || h <= 0) return 0;  // _image.empty() => silent return (:815-816)
    o->ComputePyramid(gray, w, h, stride);
    try {
        o->ComputeKeyPointsOctTree();
    } catch (const std::domain_error &) {
        return -2;  // input geometry for which the reference itself is undefined
    }
    int nkeypoints = 0;
    for (int level = 0; level < o->nlevels; ++level) nkeypoints += (int)o->levelKps[level].size();
    if (nkeypoints > cap) return -1;
    int offset = 0;
    for (int level = 0; level < o->nlevels; ++level) {
        std::vector<orc_keypoint> &keypoints = o->levelKps[level];
        Image &bl = o->blurred[level];
        const Image &im = o->pyramid[level];
        bl.w = im.w, bl.h = im.h;
        if (keypoints.empty()) {
            bl.d.clear();
            continue;
        }
        bl.d.resize(im.d.size());
        gaussian_blur_7x7(im.d.data(), im.w, im.h, im.w, bl.d.data(), bl.w);
        for (size_t i = 0; i < keypoints.size(); i++)
            o->computeOrbDescriptor(keypoints[i], bl, desc + (size_t)(offset + i) * 32);
        for (size_t i = 0; i < keypoints.size(); i++) {
            orc_keypoint kp = keypoints[i];
            if (level != 0) {
                float scale = o->mvScaleFactor[level];
                kp.x *= scale;
                kp.y *= scale;
            }
            kps[offset + i] = kp;
        }
        offset += (int)keypoints.size();
    }
    return nkeypoints;
}

int orc_orb_level_size(const orc_orb *o, int level, int *w, int *h) {
    if (level < 0 || level >= o->nlevels) return -1;
    *w = o->pyramid[level].w;
    *h = o->pyramid[level].h;
    return 0;
}
const uint8_t *orc_orb_level_image(const orc_orb *o, int level) { return o->pyramid[level].d.data(); }
const uint8_t *orc_orb_level_blurred(const orc_orb *o, int level) {
    return o->blurred[level].d.empty() ? nullptr : o->blurred[level].d.data();
}
int orc_orb_level_candidates(const orc_orb *o, int level, int32_t *xyr, int cap) {
    const auto &c = o->candidates[level];
    if (xyr) {
        int n = std::min((int)c.size(), cap);
        for (int i = 0; i < n; i++) xyr[3 * i] = c[i].x, xyr[3 * i + 1] = c[i].y, xyr[3 * i + 2] = c[i].r;
    }
    return (int)c.size();
}
int orc_orb_level_keypoints(const orc_orb *o, int level, orc_keypoint *kps, int cap) {
    const auto &c = o->levelKps[level];
    if (kps) {
        int n = std::min((int)c.size(), cap);
        for (int i = 0; i < n; i++) kps[i] = c[i];
    }
    return (int)c.size();
}

void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh,
                          int dstride) {
    resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_gaussian_blur_7x7_s2_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    gaussian_blur_7x7(src, w, h, sstride, dst, dstride);
}
int orc_fast_9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nms, int32_t *xyr, int cap) {
    std::vector<XYR> out;
    fast_9_16(img, w, h, stride, threshold, nms != 0, out);
    int n = std::min((int)out.size(), cap);
    for (int i = 0; i < n; i++) xyr[3 * i] = out[i].x, xyr[3 * i + 1] = out[i].y, xyr[3 * i + 2] = out[i].r;
    return (int)out.size();
}
void orc_fast_score_map(const uint8_t *img, int w, int h, int stride, uint8_t *smax, int ostride) {
    int pixel[25];
    make_offsets(pixel, stride);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            smax[(size_t)y * ostride + x] =
                (y < 3 || x < 3 || y >= h - 3 || x >= w - 3) ? 0 : (uint8_t)smax_9_16(img + (size_t)y * stride + x, pixel);
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_cv_round_f(float v) { return cv_round(v); }

}  // extern "C"
