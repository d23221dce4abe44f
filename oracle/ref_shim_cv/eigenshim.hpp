// eigenshim.hpp -- stand-in for the Eigen headers src/PlaneExtractor.cpp and include/peac/ include (TEST
// INFRASTRUCTURE, see oracle/ref_plane_wrap.cpp).  The build image has no Eigen.  What the plane path uses of it is
// Eigen::Vector3d as a 3-double record and ONE algorithm, SelfAdjointEigenSolver<Matrix3d> (include/peac/eig33sym.hpp:
// 71-75).  The record is trivial; the solver is the oracle's cyclic Jacobi solver (orc_eig33sym) -- the repository's
// stated substitute for Eigen's tridiagonal QL ("parity unpinned" for the solver itself, cross-checked against LAPACK).
// Building the reference against this header therefore validates everything of the plane restatement EXCEPT the
// eigen-solver's last bits: readDepthImage, ImagePointCloud::get, the PlaneSeg constructor, Stats, the thresholds of
// ParamSet, initGraph's node test and its idiosyncratic edge stepping -- and gives the real ahCluster / refineDetails.
#pragma once
#include "msl_oracle.h"
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };
template <typename T, int R, int C, int Opt = ColMajor> class Matrix {
public:
    T m[R * C];  // storage order per Opt
    Matrix() {}
    template <int Opt2> Matrix(const Matrix<T, R, C, Opt2> &o) {  // between storage orders
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C; j++) (*this)(i, j) = o(i, j);
    }
    const T *data() const { return m; }
    T *data() { return m; }
    Matrix(T x, T y, T z) {
        static_assert(R * C == 3, "3-vector constructor");
        m[0] = x, m[1] = y, m[2] = z;
    }
    T &operator[](int i) { return m[i]; }
    const T &operator[](int i) const { return m[i]; }
    T &operator()(int i) { return m[i]; }
    const T &operator()(int i) const { return m[i]; }
    T &operator()(int i, int j) { return Opt == RowMajor ? m[i * C + j] : m[j * R + i]; }
    const T &operator()(int i, int j) const { return Opt == RowMajor ? m[i * C + j] : m[j * R + i]; }
};
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 4, 4> Matrix4f;

template <typename M> class Map;
template <typename T, int R, int C, int Opt> class Map<Matrix<T, R, C, Opt>> {
public:
    T *p;
    Map(T *ptr, int, int) : p(ptr) {}
    Map(T *ptr) : p(ptr) {}
    T &operator()(int i, int j) { return Opt == RowMajor ? p[i * C + j] : p[j * R + i]; }
    const T &operator()(int i, int j) const { return Opt == RowMajor ? p[i * C + j] : p[j * R + i]; }
    template <int Opt2> Map &operator=(const Matrix<T, R, C, Opt2> &o) {
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C; j++) (*this)(i, j) = o(i, j);
        return *this;
    }
};

template <typename M> class SelfAdjointEigenSolver;
template <> class SelfAdjointEigenSolver<Matrix3d> {
public:
    template <typename Src> explicit SelfAdjointEigenSolver(const Src &a) {
        double K[9], s[3], V[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) K[3 * i + j] = a(i, j);
        orc_eig33sym(K, s, V);  // ascending eigenvalues; V[3 * k + i] = component k of eigenvector i
        for (int i = 0; i < 3; i++) {
            val_[i] = s[i];
            for (int k = 0; k < 3; k++) vec_(k, i) = V[3 * k + i];
        }
    }
    const Vector3d &eigenvalues() const { return val_; }
    const Matrix3d &eigenvectors() const { return vec_; }

private:
    Vector3d val_;
    Matrix3d vec_;
};
}  // namespace Eigen
