// ref_plane_wrap.cpp -- C entry points around the REFERENCE's own src/PlaneExtractor.cpp + include/peac/ (TEST
// INFRASTRUCTURE).
//
// oracle/Makefile compiles /root/reference/src/PlaneExtractor.cpp where it lies, unmodified (and with it the header-only
// peac fitter the reference vendors under include/peac/), against the stand-in headers of oracle/ref_shim_cv/ into
// oracle/_ref/libplane_ref.so.  OpenCV is used by that code as a pixel container only; of Eigen it uses Vector3d as a
// record and one algorithm, SelfAdjointEigenSolver<Matrix3d>, for which the stand-in substitutes the oracle's Jacobi
// solver (the repository's stated assumption, "parity unpinned" for the solver).  Everything else is the reference's own
// code: readDepthImage, ImagePointCloud::get, the PlaneSeg constructor and Stats, ParamSet's thresholds, initGraph's
// node test and edge stepping, and the whole of ahCluster / refineDetails / floodFill behind run().
//
// peac keeps a node's neighbours in a std::set<PlaneSeg*> (AHCPlaneSeg.hpp:212): candidates for a merge are visited in
// HEAP-ADDRESS order, which decides exact-MSE ties (AHCPlaneFitter.hpp, ahCluster).  As for the ORB library, every call
// runs inside a bump arena (ref_arena.hpp) so that address order is creation order.
//
// tests/test_oracle_ref.py compares the oracle's plane pre-stage against ref_plane_prestage and uses ref_plane_run as the
// producer of real membership images; nothing else uses this library.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <vector>

#include "ref_arena.hpp"

#include <opencv2/opencv.hpp>

#define protected public
#include "PlaneExtractor.h"
#undef protected

struct ref_block_stat {  // = orc_block_stat
    double center[3], normal[3], mse, curvature;
    int32_t N, nouse;
};

typedef ahc::PlaneFitter<ImagePointCloud> Fitter;

// readColorImage + readDepthImage exactly as Frame::ExtractPlanes calls them (src/Frame.cc:607-608)
static bool read_frame(PlaneDetection &pd, const uint16_t *depth, int w, int h, int stride_px, float fx, float fy, float cx,
                       float cy, float factor, cv::Mat &color, cv::Mat &K) {
    color.create(h, w, CV_8UC3);
    color.setTo(cv::Vec3b(0, 0, 0));
    K.create(3, 3, CV_32FC1);
    K.setTo(0.0f);
    K.at<float>(0, 0) = fx, K.at<float>(1, 1) = fy, K.at<float>(0, 2) = cx, K.at<float>(1, 2) = cy, K.at<float>(2, 2) = 1.0f;
    cv::Mat dm(h, w, CV_16UC1, (void *)depth, sizeof(uint16_t) * (size_t)stride_px);
    return pd.readColorImage(color) && pd.readDepthImage(dm, K, factor);
}

extern "C" {

// P1-P5: cloud (h2*w2*3 doubles), one record per 10x10 block built by the reference's PlaneSeg
// constructor, seed[b] = 1 if initGraph makes block b a node, edges[b] bit0=left,1=right,2=up,3=down.
// center / normal of a block with N < 4 are indeterminate in the reference; they are returned as 0.
int ref_plane_prestage(const uint16_t *depth, int w, int h, int stride_px, float fx, float fy, float cx, float cy, float factor,
                       double *cloud_xyz, ref_block_stat *blocks, uint8_t *seed, uint8_t *edges) {
    if (ref_arena_begin() != 0) return -2;
    int rc = 0;
    {
        PlaneDetection pd;
        cv::Mat color, K;
        if (!read_frame(pd, depth, w, h, stride_px, fx, fy, cx, cy, factor, color, K)) rc = -1;
        if (rc == 0) {
            const int W2 = pd.cloud.w, H2 = pd.cloud.h;
            if (cloud_xyz)
                for (int i = 0; i < W2 * H2; i++)
                    for (int k = 0; k < 3; k++) cloud_xyz[3 * i + k] = pd.cloud.vertices[i][k];
            Fitter &f = pd.plane_filter;
            const int Nh = H2 / f.windowHeight, Nw = W2 / f.windowWidth;
            for (int i = 0; i < Nh; i++)
                for (int j = 0; j < Nw; j++) {
                    ahc::PlaneSeg p(pd.cloud, i * Nw + j, i * f.windowHeight, j * f.windowWidth, W2, H2, f.windowWidth, f.windowHeight,
                                    f.params);
                    ref_block_stat &b = blocks[i * Nw + j];
                    memset(&b, 0, sizeof(b));
                    if (p.N >= 4)
                        for (int k = 0; k < 3; k++) b.center[k] = p.center[k], b.normal[k] = p.normal[k];
                    b.mse = p.mse, b.curvature = p.curvature, b.N = p.N, b.nouse = p.nouse ? 1 : 0;
                    seed[i * Nw + j] = 0, edges[i * Nw + j] = 0;
                }
            // the first lines of PlaneFitter::run (AHCPlaneFitter.hpp:216-223), then initGraph alone
            f.clear();
            f.points = &pd.cloud;
            f.height = pd.cloud.height();
            f.width = pd.cloud.width();
            f.ds.reset(new DisjointSet((f.height / f.windowHeight) * (f.width / f.windowWidth)));
            Fitter::PlaneSegMinMSEQueue minQ;
            f.initGraph(minQ);
            std::vector<ahc::PlaneSeg::shared_ptr> keep;  // nodes must outlive the neighbour sets that point at them
            while (!minQ.empty()) {
                keep.push_back(minQ.top());
                minQ.pop();
            }
            for (size_t k = 0; k < keep.size(); k++) {
                const ahc::PlaneSeg &p = *keep[k];
                seed[p.rid] = 1;
                for (ahc::PlaneSeg::NbSet::const_iterator it = p.nbs.begin(); it != p.nbs.end(); ++it) {
                    const int d = (*it)->rid - p.rid;
                    edges[p.rid] |= d == -1 ? 1 : d == 1 ? 2 : d == -Nw ? 4 : d == Nw ? 8 : 16;
                }
            }
            for (size_t k = 0; k < keep.size(); k++) keep[k]->nbs.clear();
            f.clear();
        }
    }
    if (ref_arena_end() != 0) return -3;
    return rc;
}

// Frame::ExtractPlanes' three calls (src/Frame.cc:607-609).  membership: h2*w2 int32 = plane_filter.membershipImg
// (-1 = none, <= -2 = floodFill's trail counters, >= 0 = plane id); per plane (<= cap): normal, center, N of
// extractedPlanes[i] and the size of plane_vertices_[i].  Returns plane_num_, < 0 on error.
int ref_plane_run(const uint16_t *depth, int w, int h, int stride_px, float fx, float fy, float cx, float cy, float factor,
                  int32_t *membership, double *plane_normal, double *plane_center, int32_t *plane_N, int32_t *plane_vertices,
                  int cap) {
    if (ref_arena_begin() != 0) return -2;
    int rc = 0;
    {
        PlaneDetection pd;
        cv::Mat color, K;
        if (!read_frame(pd, depth, w, h, stride_px, fx, fy, cx, cy, factor, color, K)) rc = -1;
        if (rc == 0) {
            pd.runPlaneDetection();
            const cv::Mat &m = pd.plane_filter.membershipImg;
            for (int y = 0; y < m.rows; y++) memcpy(membership + (size_t)y * m.cols, m.ptr(y), sizeof(int32_t) * (size_t)m.cols);
            rc = pd.plane_num_;
            for (int i = 0; i < rc && i < cap; i++) {
                const ahc::PlaneSeg &p = *pd.plane_filter.extractedPlanes[i];
                for (int k = 0; k < 3; k++) plane_normal[3 * i + k] = p.normal[k], plane_center[3 * i + k] = p.center[k];
                plane_N[i] = p.N;
                plane_vertices[i] = (int32_t)pd.plane_vertices_[i].size();
            }
        }
    }
    if (ref_arena_end() != 0) return -3;
    return rc;
}

// Timing form for bench.py's reference arm (no arena: glibc's allocator, so it may run on many threads at once, one frame
// each, like Frame::ExtractPlanes' per-frame std::thread).  full = 0: readColorImage + readDepthImage and the pre-stage of
// PlaneFitter::run -- the PlaneSeg constructor of every block and initGraph (AHCPlaneFitter.hpp:216-223, 756-928), the part
// BASELINE.json's config 3 names; full = 1: runPlaneDetection as Frame::ExtractPlanes calls it (src/Frame.cc:607-609).
// membership (optional, full = 1): membershipImg out.  Returns the number of graph nodes (full = 0) or plane_num_ (full = 1),
// < 0 on error.  With glibc's allocator exact mse ties may resolve differently from the arena build (see ref_arena.hpp).
int ref_plane_timed(const uint16_t *depth, int w, int h, int stride_px, float fx, float fy, float cx, float cy, float factor,
                    int full, int32_t *membership) {
    PlaneDetection pd;
    cv::Mat color, K;
    if (!read_frame(pd, depth, w, h, stride_px, fx, fy, cx, cy, factor, color, K)) return -1;
    if (full) {
        pd.runPlaneDetection();
        if (membership) {  // what Tracking hands to SurfelMapping (src/Tracking.cc:227-229)
            const cv::Mat &m = pd.plane_filter.membershipImg;
            for (int y = 0; y < m.rows; y++) memcpy(membership + (size_t)y * m.cols, m.ptr(y), sizeof(int32_t) * (size_t)m.cols);
        }
        return pd.plane_num_;
    }
    Fitter &f = pd.plane_filter;
    f.clear();
    f.points = &pd.cloud;
    f.height = pd.cloud.height();
    f.width = pd.cloud.width();
    f.ds.reset(new DisjointSet((f.height / f.windowHeight) * (f.width / f.windowWidth)));
    Fitter::PlaneSegMinMSEQueue minQ;
    f.initGraph(minQ);
    const int nodes = (int)minQ.size();
    std::vector<ahc::PlaneSeg::shared_ptr> keep;
    while (!minQ.empty()) {
        keep.push_back(minQ.top());
        minQ.pop();
    }
    for (size_t k = 0; k < keep.size(); k++) keep[k]->nbs.clear();
    f.clear();
    return nodes;
}

}  // extern "C"
