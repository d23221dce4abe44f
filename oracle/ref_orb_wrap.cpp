// ref_orb_wrap.cpp -- C entry points around the REFERENCE's own src/ORBextractor.cc (TEST INFRASTRUCTURE).
//
// oracle/Makefile compiles /root/reference/src/ORBextractor.cc where it lies, unmodified, against the stand-in OpenCV
// header oracle/ref_shim_cv/cvshim.hpp into oracle/_ref/liborb_ref.so.  The five OpenCV algorithms that file calls are
// the oracle's cv2-pinned primitives (see the header); everything else -- the constructor tables, the FAST cell loop,
// DistributeOctTree / DivideNode, IC_Angle, computeOrbDescriptor, the level loop of operator() -- is the reference's own
// code.  tests/test_oracle_ref.py compares the oracle restatement against it; nothing else uses it.
//
// The one thing the reference leaves to the platform is the tie-break of DistributeOctTree's largest-first expansion:
// it sorts (key count, ExtractorNode*) pairs (src/ORBextractor.cc:654), so nodes with equal counts are ordered by HEAP
// ADDRESS.  The oracle substitutes the creation sequence of the nodes.  To make the reference itself deterministic and
// comparable, every allocation made while ref_orb_extract runs comes from a bump arena that never reuses memory:
// addresses then grow in allocation order, i.e. the address order of list nodes IS their creation order -- one legal
// allocator among many, the one whose outcome the oracle restates.  ref_orb_extract_malloc runs the same code on the
// process allocator (whatever order glibc's free lists produce), for the experiment in tools/ref_octree_tiebreak.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ref_arena.hpp"

#include <opencv2/core/core.hpp>

#define protected public
#include "ORBextractor.h"
#undef protected

struct ref_keypoint {  // = orc_keypoint
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

static int run(int use_arena, int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, const uint8_t *gray,
               int w, int h, int stride, ref_keypoint *kps, uint8_t *desc, int cap, int32_t *level_wh, uint8_t *pyramid) {
    if (use_arena && ref_arena_begin() != 0) return -2;
    int n = 0;
    {
        ORB_SLAM2::ORBextractor ext(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
        cv::Mat image(h, w, CV_8UC1, (void *)gray, (size_t)stride), mask, descriptors;
        std::vector<cv::KeyPoint> keypoints;
        ext(image, mask, keypoints, descriptors);
        n = (int)keypoints.size();
        for (int i = 0; i < n && i < cap; i++) {
            const cv::KeyPoint &k = keypoints[i];
            kps[i] = ref_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
            memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
        }
        size_t off = 0;
        for (int l = 0; l < nlevels && level_wh; l++) {
            const cv::Mat &m = ext.mvImagePyramid[l];
            level_wh[2 * l] = m.cols;
            level_wh[2 * l + 1] = m.rows;
            if (pyramid)
                for (int y = 0; y < m.rows; y++, off += m.cols) memcpy(pyramid + off, m.ptr(y), m.cols);
        }
    }
    if (use_arena && ref_arena_end() != 0) return -3;
    return n;
}

extern "C" {

// One ORBextractor built, called once on `gray` (8-bit, `stride` bytes per row) and destroyed.  Returns the number of
// keypoints (the first min(n, cap) are written), < 0 on error.  level_wh: 2 x nlevels ints; pyramid: the levels' bytes,
// dense, one after the other (either may be NULL).
int ref_orb_extract(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, const uint8_t *gray, int w,
                    int h, int stride, ref_keypoint *kps, uint8_t *desc, int cap, int32_t *level_wh, uint8_t *pyramid) {
    return run(1, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, gray, w, h, stride, kps, desc, cap, level_wh, pyramid);
}
int ref_orb_extract_malloc(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, const uint8_t *gray,
                           int w, int h, int stride, ref_keypoint *kps, uint8_t *desc, int cap, int32_t *level_wh,
                           uint8_t *pyramid) {
    return run(0, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, gray, w, h, stride, kps, desc, cap, level_wh, pyramid);
}

// constructor tables: scale[nlevels], inv_scale, sigma2, inv_sigma2, features per level[nlevels], umax[16]
void ref_orb_tables(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, float *scale, float *inv_scale,
                    float *sigma2, float *inv_sigma2, int32_t *per_level, int32_t *umax16) {
    ORB_SLAM2::ORBextractor ext(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    for (int i = 0; i < nlevels; i++) {
        scale[i] = ext.mvScaleFactor[i];
        inv_scale[i] = ext.mvInvScaleFactor[i];
        sigma2[i] = ext.mvLevelSigma2[i];
        inv_sigma2[i] = ext.mvInvLevelSigma2[i];
        per_level[i] = ext.mnFeaturesPerLevel[i];
    }
    for (int i = 0; i < 16; i++) umax16[i] = ext.umax[i];
}

}  // extern "C"
