// Minimal stand-in for the OpenCV types the bindings in adapters/ touch -- ONLY for tools/check_adapters.sh, which
// syntax-checks adapters/ORBmatcher_msl.cc and adapters/MapPoint_msl.cc in a container without OpenCV.  Signatures follow
// opencv2/core/mat.hpp; nothing here is ever linked or shipped.
#pragma once
#include <cstddef>
#include <vector>
#define CV_32F 5
typedef unsigned char uchar;
namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
class Mat {
public:
    Mat();
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
    int rows, cols;
};
}  // namespace cv
