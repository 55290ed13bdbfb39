// rs_eigen.h -- the handful of Eigen types that appear in the reference's hot-path signatures
// (minimal.h:41-160, nonlinearRefinement.h:77-111, camera.h, rsframe.h, scanline.h).
//
// Where a real Eigen is installed it is used unchanged.  This container has none, so a minimal
// stand-in with the same names, the same (column-major) layouts and the members the reference's
// callers actually touch (main.cc:398-523, errorMeasure.cpp:66-226) is provided: element access
// by (r,c) / (i) / [i], cols()/rows()/size(), Ones/Zero, conservativeResize, row(r) *= s,
// scalar *= and +=, dot/norm, transpose() for printing.  It is NOT a linear-algebra library:
// all numerics of the path live behind the C ABI (include/rsdsfm.h).
#pragma once

#if defined(RSDSFM_USE_REAL_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Dense>))
#include <Eigen/Dense>
#else

#include <cmath>
#include <cstddef>
#include <ostream>
#include <vector>

namespace Eigen {

// Fixed-size column vector (Vector2d, Vector3d).
template <int N>
struct FixedVector {
    double v[N];
    FixedVector() { for (int i = 0; i < N; ++i) v[i] = 0.0; }
    FixedVector(double a, double b) { static_assert(N == 2, "size"); v[0] = a; v[1] = b; }
    FixedVector(double a, double b, double c) { static_assert(N == 3, "size"); v[0] = a; v[1] = b; v[2] = c; }
    static FixedVector Zero() { return FixedVector(); }
    double &operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double &x() { return v[0]; }
    double &y() { return v[1]; }
    double &z() { static_assert(N >= 3, "size"); return v[2]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { static_assert(N >= 3, "size"); return v[2]; }
    double coeff(int i) const { return v[i]; }
    int size() const { return N; }
    int rows() const { return N; }
    int cols() const { return 1; }
    const double *data() const { return v; }
    double *data() { return v; }
    double dot(const FixedVector &o) const { double s = 0; for (int i = 0; i < N; ++i) s += v[i] * o.v[i]; return s; }
    double norm() const { return std::sqrt(dot(*this)); }
    FixedVector &operator*=(double s) { for (int i = 0; i < N; ++i) v[i] *= s; return *this; }
    FixedVector operator*(double s) const { FixedVector r = *this; r *= s; return r; }
    FixedVector operator+(const FixedVector &o) const { FixedVector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
    FixedVector operator-(const FixedVector &o) const { FixedVector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
    const FixedVector &transpose() const { return *this; }      // only ever used for printing
};
template <int N>
inline std::ostream &operator<<(std::ostream &os, const FixedVector<N> &a)
{
    for (int i = 0; i < N; ++i) os << (i ? " " : "") << a.v[i];
    return os;
}
typedef FixedVector<2> Vector2d;
typedef FixedVector<3> Vector3d;

// Fixed 3x3 (row access m(r,c); storage column-major like Eigen).
struct Matrix3d {
    double m[9];
    Matrix3d() { for (int i = 0; i < 9; ++i) m[i] = 0.0; }
    static Matrix3d Identity() { Matrix3d r; r(0, 0) = r(1, 1) = r(2, 2) = 1.0; return r; }
    static Matrix3d Zero() { return Matrix3d(); }
    double &operator()(int r, int c) { return m[r + 3 * c]; }
    double operator()(int r, int c) const { return m[r + 3 * c]; }
    Matrix3d transpose() const { Matrix3d t; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t(c, r) = (*this)(r, c); return t; }
    Vector3d operator*(const Vector3d &x) const
    {
        Vector3d y;
        for (int r = 0; r < 3; ++r) y(r) = (*this)(r, 0) * x(0) + (*this)(r, 1) * x(1) + (*this)(r, 2) * x(2);
        return y;
    }
};
inline std::ostream &operator<<(std::ostream &os, const Matrix3d &a)
{
    for (int r = 0; r < 3; ++r) os << a(r, 0) << " " << a(r, 1) << " " << a(r, 2) << (r < 2 ? "\n" : "");
    return os;
}

// Dynamic column vector / 1-D array (ArrayXd, VectorXd).
struct ArrayXd {
    std::vector<double> d;
    ArrayXd() {}
    explicit ArrayXd(std::ptrdiff_t n) : d((size_t)n, 0.0) {}
    static ArrayXd Zero(std::ptrdiff_t n) { return ArrayXd(n); }
    static ArrayXd Ones(std::ptrdiff_t n) { ArrayXd a(n); for (auto &x : a.d) x = 1.0; return a; }
    double &operator()(std::ptrdiff_t i) { return d[(size_t)i]; }
    double operator()(std::ptrdiff_t i) const { return d[(size_t)i]; }
    double &operator[](std::ptrdiff_t i) { return d[(size_t)i]; }
    double operator[](std::ptrdiff_t i) const { return d[(size_t)i]; }
    std::ptrdiff_t size() const { return (std::ptrdiff_t)d.size(); }
    std::ptrdiff_t rows() const { return size(); }
    std::ptrdiff_t cols() const { return 1; }
    const double *data() const { return d.data(); }
    double *data() { return d.data(); }
    ArrayXd &operator*=(double s) { for (auto &x : d) x *= s; return *this; }
    ArrayXd &operator+=(double s) { for (auto &x : d) x += s; return *this; }
    double mean() const { double s = 0; for (double x : d) s += x; return d.empty() ? 0.0 : s / (double)d.size(); }
    void conservativeResize(std::ptrdiff_t n) { d.resize((size_t)n, 0.0); }
    ArrayXd head(std::ptrdiff_t n) const { ArrayXd a(n); for (std::ptrdiff_t i = 0; i < n; ++i) a(i) = d[(size_t)i]; return a; }
};
typedef ArrayXd VectorXd;

// R x n column-major array with a compile-time number of rows (Array2Xd, Matrix2Xd, Array3Xd).
template <int R>
struct FixedRowsArray {
    std::vector<double> d;
    std::ptrdiff_t n = 0;
    FixedRowsArray() {}
    FixedRowsArray(std::ptrdiff_t rows_, std::ptrdiff_t cols_) : d((size_t)(R * cols_), 0.0), n(cols_) { (void)rows_; }
    static FixedRowsArray Zero(std::ptrdiff_t rows_, std::ptrdiff_t cols_) { return FixedRowsArray(rows_, cols_); }
    static FixedRowsArray Ones(std::ptrdiff_t rows_, std::ptrdiff_t cols_) { FixedRowsArray a(rows_, cols_); for (auto &x : a.d) x = 1.0; return a; }
    double &operator()(int r, std::ptrdiff_t c) { return d[(size_t)(r + R * c)]; }
    double operator()(int r, std::ptrdiff_t c) const { return d[(size_t)(r + R * c)]; }
    std::ptrdiff_t cols() const { return n; }
    std::ptrdiff_t rows() const { return R; }
    const double *data() const { return d.data(); }
    double *data() { return d.data(); }
    void conservativeResize(std::ptrdiff_t rows_, std::ptrdiff_t cols_) { (void)rows_; d.resize((size_t)(R * cols_), 0.0); n = cols_; }
    struct RowRef {
        FixedRowsArray *a; int r;
        RowRef &operator*=(double s) { for (std::ptrdiff_t c = 0; c < a->n; ++c) (*a)(r, c) *= s; return *this; }
    };
    RowRef row(int r) { return RowRef{this, r}; }
    FixedVector<R> col(std::ptrdiff_t c) const { FixedVector<R> v; for (int r = 0; r < R; ++r) v(r) = (*this)(r, c); return v; }
};
typedef FixedRowsArray<2> Array2Xd;
typedef FixedRowsArray<2> Matrix2Xd;
typedef FixedRowsArray<3> Array3Xd;

// Dynamic column-major matrix (MatrixXd): the depth map, (y, x) at y + x*rows.
struct MatrixXd {
    std::vector<double> d;
    std::ptrdiff_t r = 0, c = 0;
    MatrixXd() {}
    MatrixXd(std::ptrdiff_t rows_, std::ptrdiff_t cols_) : d((size_t)(rows_ * cols_), 0.0), r(rows_), c(cols_) {}
    static MatrixXd Zero(std::ptrdiff_t rows_, std::ptrdiff_t cols_) { return MatrixXd(rows_, cols_); }
    double &operator()(std::ptrdiff_t y, std::ptrdiff_t x) { return d[(size_t)(y + x * r)]; }
    double operator()(std::ptrdiff_t y, std::ptrdiff_t x) const { return d[(size_t)(y + x * r)]; }
    double coeff(std::ptrdiff_t y, std::ptrdiff_t x) const { return (*this)(y, x); }
    std::ptrdiff_t rows() const { return r; }
    std::ptrdiff_t cols() const { return c; }
    const double *data() const { return d.data(); }
    double *data() { return d.data(); }
};

}  // namespace Eigen
#endif
