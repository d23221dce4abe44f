// eigenshim.hpp -- stand-in for the Eigen headers the reference's hot-path sources include (TEST INFRASTRUCTURE, see
// oracle/ref_plane_wrap.cpp and oracle/ref_wrap.cpp).  The build image has no Eigen.  src/SurfelFusion.cpp uses fixed-size
// matrices: element access, Zero(), block<>(), products and the 4x4 inverse -- evaluated here as this repository assumes
// Eigen evaluates them (left-to-right row-by-column products; cofactor inverse with one reciprocal of the determinant), the
// same the oracle restatement uses.  What the plane path uses is Eigen::Vector3d as a 3-double record and ONE algorithm,
// SelfAdjointEigenSolver<Matrix3d> (include/peac/eig33sym.hpp:71-75).  The record is trivial; the solver is the oracle's cyclic Jacobi solver (orc_eig33sym) -- the repository's
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
    static Matrix Zero() {
        Matrix z;
        for (int i = 0; i < R * C; i++) z.m[i] = 0;
        return z;
    }
    template <int BR, int BC> Matrix<T, BR, BC> block(int i0, int j0) const {
        Matrix<T, BR, BC> b;
        for (int i = 0; i < BR; i++)
            for (int j = 0; j < BC; j++) b(i, j) = (*this)(i0 + i, j0 + j);
        return b;
    }
    // 4x4 inverse by cofactors with one reciprocal of the determinant -- this repository's stated assumption about Eigen's
    // Matrix4f / Matrix4d::inverse() (src/SurfelFusion.cpp:59,153), the same the oracle restatement uses
    Matrix inverse() const {
        static_assert(R == 4 && C == 4, "only the 4x4 inverse is needed");
        T q[16], a[16];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) q[4 * i + j] = (*this)(i, j);
        inverse_rows(q, a);
        Matrix inv;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) inv(i, j) = a[4 * i + j];
        return inv;
    }

private:
    static void inverse_rows(const T *m, T *out) {  // row-major in, row-major out
        T a[16];
        a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
        a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
        a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
        a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
        a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
        a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
        a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
        a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
        a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
        a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
        a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
        a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
        a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
        a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
        a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
        a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
        T det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
        det = (T)1 / det;
        for (int i = 0; i < 16; i++) out[i] = a[i] * det;
    }
};
// fixed-size products: row by column, accumulated left to right (the stated assumption about Eigen's evaluation order)
template <typename T, int R, int K, int C, int O1, int O2> Matrix<T, R, C> operator*(const Matrix<T, R, K, O1> &a, const Matrix<T, K, C, O2> &b) {
    Matrix<T, R, C> o;
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < K; k++) s = s + a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 4, 1> Vector4d;

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
