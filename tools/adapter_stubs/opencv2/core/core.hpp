// Minimal stand-in for the OpenCV types the bindings in adapters/ touch -- ONLY for tools/check_adapters.sh, which
// type-checks adapters/*.cc|*.cpp in a container without OpenCV.  Signatures follow opencv2/core/{mat,types,matx}.hpp;
// nothing here is ever linked or shipped.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <vector>
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32SC1 4
#define CV_Assert(expr) do { if (!(expr)) throw 1; } while (0)
typedef unsigned char uchar;
namespace cv {
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;
template <typename T, int n> struct Vec {
    T val[n];
    Vec() {}
    Vec(T a, T b, T c);
    template <typename T2> operator Vec<T2, n>() const;
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<double, 3> Vec3d;
struct KeyPoint {
    KeyPoint();
    KeyPoint(float x, float y, float size, float angle = -1, float response = 0, int octave = 0, int class_id = -1);
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
struct MatStep { operator size_t() const; };
class Mat {
public:
    Mat();
    Mat(int rows, int cols, int type);
    void create(int rows, int cols, int type);
    template <typename T> T &at(int i);
    template <typename T> const T &at(int i) const;
    template <typename T> T &at(int i, int j);
    template <typename T> const T &at(int i, int j) const;
    uchar *ptr(int row = 0);
    const uchar *ptr(int row = 0) const;
    template <typename T> T *ptr(int row = 0);
    template <typename T> const T *ptr(int row = 0) const;
    void convertTo(Mat &m, int rtype, double alpha = 1, double beta = 0) const;
    Mat clone() const;
    Mat row(int y) const;
    int type() const;
    int depth() const;
    bool empty() const;
    size_t total() const;
    bool isContinuous() const;
    int rows, cols;
    uchar *data;
    MatStep step;
};
class _InputArray { public: bool empty() const; Mat getMat(int idx = -1) const; };
class _OutputArray : public _InputArray { public: void release() const; void create(int rows, int cols, int type) const; };
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
}  // namespace cv
