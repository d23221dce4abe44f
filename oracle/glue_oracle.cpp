// glue_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, see msl_oracle.h) for the frame glue either side of the ORB
// extractor:
//   Tracking::GrabImage      src/Tracking.cc:184-211   cvtColor RGB/BGR(A) -> GRAY, depth convertTo(CV_32F, factor)
//   Frame::UndistortKeyPoints  src/Frame.cc:437-463    cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK)
//   Frame::ComputeStereoFromRGBD src/Frame.cc:495-513  depth lookup at the (truncated) keypoint, uRight = x_un - bf / d
// The OpenCV arithmetic (absent from /root/reference) is restated from its published algorithms and pinned bit-exactly
// against cv2 4.13 in tests/test_oracle_primitives.py (cvtColor, undistortPoints); Mat::convertTo has no Python
// binding to pin against ("parity unpinned": dst = (float)src * (float)alpha, one rounding).
#include "msl_oracle.h"

#include <cmath>

extern "C" {

// cv::cvtColor(..., CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) on CV_8U: 15-bit fixed point,
// gray = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (OpenCV >= 4.0; color_rgb.simd.hpp RGB2Gray<uchar>)
void orc_cvt_gray(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order, uint8_t *dst, int dstride) {
    const int ri = rgb_order ? 0 : 2, bi = rgb_order ? 2 : 0;
    for (int y = 0; y < h; y++) {
        const uint8_t *s = src + (size_t)y * stride;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++, s += channels)
            d[x] = (uint8_t)((s[ri] * 9798 + s[1] * 19235 + s[bi] * 3735 + (1 << 14)) >> 15);
    }
}

// imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) (src/Tracking.cc:205-207), CV_16U source
void orc_depth_to_float(const uint16_t *src, int64_t n, float factor, float *dst) {
    for (int64_t i = 0; i < n; i++) dst[i] = (float)src[i] * factor;
}

// cv::undistortPoints(src, dst, K, D(k1,k2,p1,p2,k3), noArray(), K): 5 fixed-point iterations in double
// (TermCriteria(MAX_ITER, 5, 0.01)), then x' = fx*x + cx (P = K, R = I), stored as float.
void orc_undistort_points(int n, const float *xy, const float K4[4], const float D5[5], float *out) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k[5] = {D5[0], D5[1], D5[2], D5[3], D5[4]};
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) {  // test: undistortPoints regression, opencv PR #11958
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0 * r2 + 0 * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}

// Frame::UndistortKeyPoints, src/Frame.cc:437-463: identity when mDistCoef[0] == 0
void orc_undistort_keypoints(int n, const float *xy, const float K4[4], const float D5[5], float *out) {
    if (D5[0] == 0.0f) {
        for (int i = 0; i < 2 * n; i++) out[i] = xy[i];
        return;
    }
    orc_undistort_points(n, xy, K4, D5, out);
}

// Frame::ComputeStereoFromRGBD, src/Frame.cc:495-513
void orc_stereo_from_rgbd(int n, const float *kp_xy, const float *kpun_xy, const float *depth, int w, float mbf,
                          float *uright, float *kdepth) {
    for (int i = 0; i < n; i++) {
        uright[i] = -1, kdepth[i] = -1;
        const float v = kp_xy[2 * i + 1], u = kp_xy[2 * i];
        const float d = depth[(size_t)(int)v * w + (int)u];  // imDepth.at<float>(v, u): float -> int truncation
        if (d > 0) {
            kdepth[i] = d;
            uright[i] = kpun_xy[2 * i] - mbf / d;
        }
    }
}

}  // extern "C"
