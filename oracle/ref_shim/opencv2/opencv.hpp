// Stand-in for <opencv2/opencv.hpp> (TEST INFRASTRUCTURE, see oracle/ref_wrap.cpp): a non-owning cv::Mat header with
// at<T>(row, col) -- the only OpenCV facility src/SurfelFusion.cpp uses -- addressed exactly as cv::Mat does
// (data + row * step + col * sizeof(T)), which keeps the reference's at<cv::Vec3b> reads on the gray image meaningful.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
typedef unsigned char uchar;
namespace cv {
template <typename T, int n> struct Vec {
    T val[n];
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
class Mat {
public:
    uchar *data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;  // bytes per row
    Mat() {}
    Mat(int r, int c, size_t stepBytes, void *p) : data((uchar *)p), rows(r), cols(c), step(stepBytes) {}
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
};
}  // namespace cv
