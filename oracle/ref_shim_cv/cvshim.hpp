// cvshim.hpp -- stand-in for the OpenCV headers src/ORBextractor.cc includes (TEST INFRASTRUCTURE, see
// oracle/ref_orb_wrap.cpp).  The build image has no OpenCV; this header supplies just enough of cv::Mat / cv::KeyPoint /
// cv::Point_ / cv::Size / cv::Rect / cv::InputArray to compile the reference's src/ORBextractor.cc UNMODIFIED, and routes
// the five OpenCV algorithms that file calls -- cv::resize (INTER_LINEAR, 8U), cv::FAST (TYPE_9_16), cv::GaussianBlur
// (7x7, sigma 2), cv::fastAtan2, cvRound -- to the oracle's primitives, which tests/test_oracle_primitives.py pins bit
// for bit against cv2 4.13.0.  What the resulting library validates is therefore everything ELSE in the oracle's
// 750-line ORB restatement: the constructor tables, the cell loop, DistributeOctTree / DivideNode, the orientation and
// descriptor arithmetic, the level bookkeeping -- against the reference's own source.
//
// Semantics kept where the reference depends on them:
//  * Mat headers share their buffer (shared_ptr) and ROIs alias the parent (operator()(Rect), rowRange, colRange);
//  * `m = Mat::zeros(r, c, type)` onto a header of the same size and type fills IN PLACE (cv::MatExpr assignment calls
//    Mat::create, which keeps a matching buffer) -- computeDescriptors relies on that to write into the caller's rows;
//  * resize / GaussianBlur / copyMakeBorder with a destination that already has the right size write into it.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include "msl_oracle.h"  // the cv2-pinned primitives

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) assert(expr)
#define CV_8U 0
#define CV_16U 2
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

// cvRound: round half to even (lrint / cvtsd2si); the float overload is the oracle's pinned one
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return orc_cv_round_f(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) {
    int i = (int)v;
    return i - (i > v);
}
static inline int cvCeil(double v) {
    int i = (int)v;
    return i + (i < v);
}

namespace cv {

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <typename T> class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U> &p) : x((T)p.x), y((T)p.y) {}
};
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, float b) {  // cv: saturate_cast<T>(a.x * b)
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T> class Size_ {
public:
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T, int n> class Vec {
public:
    T val[n];
    Vec() {
        for (int i = 0; i < n; i++) val[i] = T(0);
    }
    Vec(T a, T b) : Vec() { val[0] = a, val[1] = b; }
    Vec(T a, T b, T c) : Vec() { val[0] = a, val[1] = b, val[2] = c; }
    explicit Vec(const T *p) {
        for (int i = 0; i < n; i++) val[i] = p[i];
    }
    template <typename U> Vec(const Vec<U, n> &o) {
        for (int i = 0; i < n; i++) val[i] = (T)o.val[i];
    }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 2> Vec2d;
typedef Vec<double, 4> Scalar;

template <typename T> class Point3_ {
public:
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<double> Point3d;
typedef Point3_<float> Point3f;

class Range {
public:
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
};

class KeyPoint {
public:
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    MatStep(size_t s) : v(s) {}
    operator size_t() const { return v; }
};

class Mat;
struct MatExprT;    // alpha * A.t()
struct MatExprMul;  // alpha * op(A) * B

struct MatZeros {  // the MatExpr of Mat::zeros
    int rows, cols, type;
};

class Mat {
public:
    uchar *data;
    int rows, cols;
    MatStep step;  // bytes per row
    Mat() : data(nullptr), rows(0), cols(0), step(0), type_(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void *p, size_t stepBytes = 0) : data((uchar *)p), rows(r), cols(c), type_(type) {
        step = stepBytes ? stepBytes : (size_t)c * elemSize();
    }
    Mat(const MatZeros &z) : Mat() { *this = z; }
    Mat(const MatExprMul &e);  // evaluates the product (below)
    void copyTo(Mat &m) const { m = clone(); }
    Mat reshape(int cn) const {  // same bytes, other channel count (continuous matrices only)
        assert(isContinuous() && ((size_t)cols * channels()) % (size_t)cn == 0);
        Mat m(*this);
        m.cols = cols * channels() / cn;
        m.type_ = CV_MAKETYPE(depth(), cn);
        m.step = (size_t)m.cols * m.elemSize();
        return m;
    }
    void convertTo(Mat &m, int rtype) const {  // same-depth conversions only: a copy
        assert((rtype & 7) == depth());
        m = clone();
    }
    Mat row(int i) const { return (*this)(Rect(0, i, cols, 1)); }
    Mat col(int j) const { return (*this)(Rect(j, 0, 1, rows)); }
    MatExprT t() const;
    double dot(const Mat &o) const;
    static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
    Mat &operator=(const MatZeros &z) {
        create(z.rows, z.cols, z.type);  // keeps a matching buffer, like cv::Mat::create
        for (int i = 0; i < rows; i++) memset(data + (size_t)i * step, 0, (size_t)cols * elemSize());
        return *this;
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;
        type_ = type;
        rows = r;
        cols = c;
        step = (size_t)c * elemSize();
        size_t n = (size_t)r * step;
        hold_.reset(new uchar[n ? n : 1], std::default_delete<uchar[]>());
        data = hold_.get();
    }
    void release() {
        hold_.reset();
        data = nullptr;
        rows = cols = 0;
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const {
        static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return (size_t)sz[depth()];
    }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    size_t step1() const { return step / elemSize1(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * (size_t)cols; }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int i = 0; i < rows; i++) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
        return m;
    }
    Mat operator()(const Rect &r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat operator()(const Range &rr, const Range &cr) const { return (*this)(Rect(cr.start, rr.start, cr.end - cr.start, rr.end - rr.start)); }
    bool isContinuous() const { return (size_t)step == (size_t)cols * elemSize() || rows <= 1; }
    template <typename T> Mat &setTo(const T &v) {  // every element of this header's window (sizeof(T) == elemSize())
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) *(T *)(data + (size_t)i * step + (size_t)j * sizeof(T)) = v;
        return *this;
    }
    // cv::Mat::at(int i0) on a 2-D matrix: flat element index when continuous (or a single row), else (i0 / cols, i0 % cols)
    template <typename T> T &at(int i0) {
        if (isContinuous()) return ((T *)data)[i0];
        return at<T>(i0 / cols, i0 % cols);
    }
    template <typename T> const T &at(int i0) const { return const_cast<Mat *>(this)->at<T>(i0); }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }

private:
    int type_;
    std::shared_ptr<uchar> hold_;
};

// ---- cv::Mat arithmetic of src/ORBmatcher.cc: CV_32F pose products only (3x3 by 3x1) ----
// cv::MatExpr folds `A * B + C` into ONE cv::gemm(A, B, 1, C, 1) and `-A.t() * B` into cv::gemm(A, B, -1, noArray(), 0,
// GEMM_1_T); which arithmetic those take (small-matrix path with float row sums vs the general path with double
// accumulation) is pinned against cv2.gemm by tests/test_oracle_primitives.py::test_cv_gemm_semantics, and evaluated here
// by the same oracle primitives (orc_cv_rx_plus_t, orc_cv_neg_rt_times_t).
struct MatExprT {
    Mat a;
    double alpha;
};
struct MatExprMul {
    Mat a, b;
    double alpha;
    bool ta;
};
inline MatExprT Mat::t() const { return MatExprT{*this, 1.0}; }
static inline MatExprT operator-(const MatExprT &e) { return MatExprT{e.a, -e.alpha}; }
static inline MatExprMul operator*(const Mat &a, const Mat &b) { return MatExprMul{a, b, 1.0, false}; }
static inline MatExprMul operator*(const MatExprT &a, const Mat &b) { return MatExprMul{a.a, b, a.alpha, true}; }
static inline void mat33_31(const Mat &a, const Mat &b, float R[9], float x[3]) {
    assert(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows == 3 && a.cols == 3 && b.rows == 3 && b.cols == 1);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R[3 * i + j] = a.at<float>(i, j);
        x[i] = b.at<float>(i, 0);
    }
}
static inline Mat mat31(const float v[3]) {
    Mat m(3, 1, CV_32FC1);
    for (int i = 0; i < 3; i++) m.at<float>(i, 0) = v[i];
    return m;
}
static inline Mat operator+(const MatExprMul &e, const Mat &c) {  // A * B + C
    assert(!e.ta && e.alpha == 1.0 && c.rows == 3 && c.cols == 1);
    float R[9], x[3], t[3], o[3];
    mat33_31(e.a, e.b, R, x);
    for (int i = 0; i < 3; i++) t[i] = c.at<float>(i, 0);
    orc_cv_rx_plus_t(R, x, t, o);
    return mat31(o);
}
inline Mat::Mat(const MatExprMul &e) : Mat() {  // -A.t() * B (the only bare product the reference forms)
    assert(e.ta && e.alpha == -1.0);
    float R[9], x[3], o[3];
    mat33_31(e.a, e.b, R, x);
    orc_cv_neg_rt_times_t(R, x, o);
    *this = mat31(o);
}
static inline Mat operator-(const Mat &a, const Mat &b) {  // 3x1 float difference
    assert(a.rows == 3 && a.cols == 1 && b.rows == 3 && b.cols == 1);
    float o[3];
    for (int i = 0; i < 3; i++) o[i] = a.at<float>(i, 0) - b.at<float>(i, 0);
    return mat31(o);
}
// element-wise helpers that only src/MapPoint.cc's normal bookkeeping uses (compiled, not part of any comparison)
static inline Mat operator+(const Mat &a, const Mat &b) {
    float o[3];
    for (int i = 0; i < 3; i++) o[i] = a.at<float>(i, 0) + b.at<float>(i, 0);
    return mat31(o);
}
static inline Mat operator/(const Mat &a, double s) {
    float o[3];
    for (int i = 0; i < 3; i++) o[i] = (float)(a.at<float>(i, 0) * (1.0 / s));
    return mat31(o);
}
static inline double norm(const Mat &a) {  // NORM_L2 of a 3x1 float vector: squares accumulated in double
    float v[3];
    for (int i = 0; i < 3; i++) v[i] = a.at<float>(i, 0);
    return orc_cv_norm3(v);
}
// cv::Mat::dot on three floats: double products, double accumulation (no Python binding to pin it; the oracle's convention)
inline double Mat::dot(const Mat &o) const {
    double s = 0;
    for (int i = 0; i < 3; i++) s += (double)at<float>(i, 0) * (double)o.at<float>(i, 0);
    return s;
}

class _InputArray {
public:
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }

protected:
    Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat &m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

// ---- the OpenCV algorithms src/ORBextractor.cc calls, on the oracle's cv2-pinned primitives ----

static inline void resize(const Mat &src, Mat &dst, Size dsize, double, double, int interpolation) {
    assert(interpolation == INTER_LINEAR && src.type() == CV_8UC1);
    dst.create(dsize.height, dsize.width, src.type());
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
// BORDER_REFLECT_101 (+ BORDER_ISOLATED: the source ROI is treated as the whole image, which is what this stand-in always
// does -- for level 0 the reference passes the caller's full image).  src may be the centre ROI of dst (the reference
// calls it that way for levels >= 1); the centre is then left alone and only the frame is written.
static inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && src.elemSize() == 1);
    dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    const bool inplace = src.data == dst.data + (size_t)top * dst.step + (size_t)left;
    for (int y = 0; y < dst.rows; y++) {
        const uchar *s = src.data + (size_t)reflect101(y - top, src.rows) * src.step;
        uchar *d = dst.data + (size_t)y * dst.step;
        for (int x = 0; x < dst.cols; x++) {
            const bool centre = y >= top && y < top + src.rows && x >= left && x < left + src.cols;
            if (inplace && centre) continue;
            d[x] = s[reflect101(x - left, src.cols)];
        }
    }
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression) with the default TYPE_9_16: KeyPoint(x, y, 7.f, -1, score)
static inline void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true) {
    keypoints.clear();
    if (image.rows < 7 || image.cols < 7) return;
    std::vector<int32_t> xyr((size_t)image.rows * image.cols * 3 + 3);
    const int n = orc_fast_9_16(image.data, image.cols, image.rows, (int)image.step, threshold, nonmaxSuppression ? 1 : 0,
                                xyr.data(), image.rows * image.cols);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1, (float)xyr[3 * i + 2]));
}

static inline void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat in = src.clone();  // the reference blurs in place
    dst.create(src.rows, src.cols, src.type());
    orc_gaussian_blur_7x7_s2_u8(in.data, in.cols, in.rows, (int)in.step, dst.data, (int)dst.step);
}

static inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::undistortPoints(src, dst, K, D, R = noArray(), P = K) on N x 1 two-channel float points: the oracle's cv2-pinned
// restatement (tests/test_oracle_primitives.py, tests/golden/glue.npz)
static inline void undistortPoints(const Mat &src, Mat &dst, const Mat &K, const Mat &D, const Mat &, const Mat &) {
    const int n = src.rows * src.cols * src.channels() / 2;
    std::vector<float> in((size_t)2 * n), out((size_t)2 * n);
    memcpy(in.data(), src.data, sizeof(float) * in.size());
    const float K4[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
    float D5[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < (int)D.total() && i < 5; i++) D5[i] = D.at<float>(i);
    orc_undistort_points(n, in.data(), K4, D5, out.data());
    dst.create(src.rows, src.cols, src.type());
    memcpy(dst.data, out.data(), sizeof(float) * out.size());
}

static inline int64_t getTickCount() { return 0; }
static inline double getTickFrequency() { return 1.0; }

}  // namespace cv
