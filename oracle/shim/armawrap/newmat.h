// TEST INFRASTRUCTURE ONLY — header shim so the UNMODIFIED reference sources under
// /root/reference compile here without FSL (FSL's armawrap/NEWMAT is not installed and is
// not part of /root/reference; no version is pinned by the reference).
// This is our own minimal dense-matrix stand-in: 1-based element access, row-major `<<`
// fill, plain left-to-right triple-loop products (no FMA, no blocking) so that results
// are reproducible against the restated oracle (oracle/msm_oracle.cpp).
// Only the NEWMAT surface the reference's hot-path files touch is provided.
#ifndef ORACLE_SHIM_NEWMAT_H
#define ORACLE_SHIM_NEWMAT_H

// FSL's real header drags these in transitively; the reference relies on that.
#include <algorithm>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <stdexcept>
#include <vector>

namespace NEWMAT {

class Matrix;
class ColumnVector;
class RowVector;
class DiagonalMatrix;

class Matrix {
protected:
    int nr = 0, nc = 0;
    std::vector<double> d;
    // state for `M << a << b << ...`
    mutable std::size_t fillpos = 0;

public:
    Matrix() = default;
    Matrix(int r, int c) : nr(r), nc(c), d((std::size_t)r * c, 0.0) {}
    virtual ~Matrix() = default;

    int Nrows() const { return nr; }
    int Ncols() const { return nc; }
    int Storage() const { return nr * nc; }

    double& operator()(int i, int j) { return d[(std::size_t)(i - 1) * nc + (j - 1)]; }
    double operator()(int i, int j) const { return d[(std::size_t)(i - 1) * nc + (j - 1)]; }
    double& element(int i, int j) { return d[(std::size_t)i * nc + j]; }
    double element(int i, int j) const { return d[(std::size_t)i * nc + j]; }

    void ReSize(int r, int c) { nr = r; nc = c; d.assign((std::size_t)r * c, 0.0); }
    void resize(int r, int c) { ReSize(r, c); }
    void Release() {}
    void CleanUp() { nr = nc = 0; d.clear(); }

    Matrix& operator=(double v) { std::fill(d.begin(), d.end(), v); return *this; }

    // `M << v0 << v1 ...` fills row-major from the start.
    struct Filler {
        Matrix* m; std::size_t pos;
        Filler& operator<<(double v) { m->d.at(pos++) = v; return *this; }
    };
    Filler operator<<(double v) { d.at(0) = v; return Filler{this, 1}; }

    Matrix t() const {
        Matrix r(nc, nr);
        for (int i = 1; i <= nr; ++i)
            for (int j = 1; j <= nc; ++j) r(j, i) = (*this)(i, j);
        return r;
    }

    double Trace() const {
        double s = 0.0;
        for (int i = 1; i <= std::min(nr, nc); ++i) s += (*this)(i, i);
        return s;
    }

    double Sum() const { double s = 0.0; for (double v : d) s += v; return s; }

    // Laplace / closed forms for the tiny sizes the reference uses (2,3,4).
    double Determinant() const {
        if (nr != nc) throw std::runtime_error("shim NEWMAT: Determinant of non-square");
        const Matrix& a = *this;
        if (nr == 1) return a(1, 1);
        if (nr == 2) return a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
        if (nr == 3)
            return a(1, 1) * (a(2, 2) * a(3, 3) - a(2, 3) * a(3, 2))
                 - a(1, 2) * (a(2, 1) * a(3, 3) - a(2, 3) * a(3, 1))
                 + a(1, 3) * (a(2, 1) * a(3, 2) - a(2, 2) * a(3, 1));
        double det = 0.0;
        for (int c = 1; c <= nc; ++c) {
            Matrix m(nr - 1, nc - 1);
            for (int i = 2; i <= nr; ++i) {
                int cc = 1;
                for (int j = 1; j <= nc; ++j) { if (j == c) continue; m(i - 1, cc++) = a(i, j); }
            }
            det += ((c % 2) ? 1.0 : -1.0) * a(1, c) * m.Determinant();
        }
        return det;
    }

    // inverse by adjugate (2x2, 3x3) — the only sizes reg_tools.cpp inverts.
    Matrix i() const {
        if (nr != nc || nr > 3 || nr < 1) throw std::runtime_error("shim NEWMAT: i() supports 1..3 square");
        const Matrix& a = *this;
        Matrix r(nr, nc);
        double det = Determinant();
        if (nr == 1) { r(1, 1) = 1.0 / a(1, 1); return r; }
        if (nr == 2) {
            r(1, 1) = a(2, 2) / det; r(1, 2) = -a(1, 2) / det;
            r(2, 1) = -a(2, 1) / det; r(2, 2) = a(1, 1) / det;
            return r;
        }
        r(1, 1) = (a(2, 2) * a(3, 3) - a(2, 3) * a(3, 2)) / det;
        r(1, 2) = (a(1, 3) * a(3, 2) - a(1, 2) * a(3, 3)) / det;
        r(1, 3) = (a(1, 2) * a(2, 3) - a(1, 3) * a(2, 2)) / det;
        r(2, 1) = (a(2, 3) * a(3, 1) - a(2, 1) * a(3, 3)) / det;
        r(2, 2) = (a(1, 1) * a(3, 3) - a(1, 3) * a(3, 1)) / det;
        r(2, 3) = (a(1, 3) * a(2, 1) - a(1, 1) * a(2, 3)) / det;
        r(3, 1) = (a(2, 1) * a(3, 2) - a(2, 2) * a(3, 1)) / det;
        r(3, 2) = (a(1, 2) * a(3, 1) - a(1, 1) * a(3, 2)) / det;
        r(3, 3) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) / det;
        return r;
    }

    Matrix SubMatrix(int r0, int r1, int c0, int c1) const {
        Matrix r(r1 - r0 + 1, c1 - c0 + 1);
        for (int i = r0; i <= r1; ++i)
            for (int j = c0; j <= c1; ++j) r(i - r0 + 1, j - c0 + 1) = (*this)(i, j);
        return r;
    }
    Matrix Row(int i) const { return SubMatrix(i, i, 1, nc); }
    Matrix Column(int j) const { return SubMatrix(1, nr, j, j); }

    Matrix& operator+=(const Matrix& b) { for (std::size_t k = 0; k < d.size(); ++k) d[k] += b.d[k]; return *this; }
    Matrix& operator-=(const Matrix& b) { for (std::size_t k = 0; k < d.size(); ++k) d[k] -= b.d[k]; return *this; }
    Matrix& operator*=(double s) { for (double& v : d) v *= s; return *this; }
    Matrix& operator/=(double s) { for (double& v : d) v /= s; return *this; }

    const std::vector<double>& raw() const { return d; }
    std::vector<double>& raw() { return d; }
};

typedef Matrix ReturnMatrix;

inline Matrix operator*(const Matrix& a, const Matrix& b) {
    if (a.Ncols() != b.Nrows()) throw std::runtime_error("shim NEWMAT: product dimension mismatch");
    Matrix r(a.Nrows(), b.Ncols());
    for (int i = 1; i <= a.Nrows(); ++i)
        for (int j = 1; j <= b.Ncols(); ++j) {
            double s = 0.0;
            for (int k = 1; k <= a.Ncols(); ++k) s += a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
inline Matrix operator+(const Matrix& a, const Matrix& b) { Matrix r = a; r += b; return r; }
inline Matrix operator-(const Matrix& a, const Matrix& b) { Matrix r = a; r -= b; return r; }
inline Matrix operator*(const Matrix& a, double s) { Matrix r = a; r *= s; return r; }
inline Matrix operator*(double s, const Matrix& a) { Matrix r = a; r *= s; return r; }
inline Matrix operator/(const Matrix& a, double s) { Matrix r = a; r /= s; return r; }
inline Matrix operator-(const Matrix& a) { Matrix r = a; r *= -1.0; return r; }

class ColumnVector : public Matrix {
public:
    ColumnVector() = default;
    explicit ColumnVector(int n) : Matrix(n, 1) {}
    ColumnVector(const Matrix& m) : Matrix(m) {
        if (nc != 1 && nr == 1) { std::swap(nr, nc); }
    }
    ColumnVector& operator=(const Matrix& m) { Matrix::operator=(m); if (nc != 1 && nr == 1) std::swap(nr, nc); return *this; }
    ColumnVector& operator=(double v) { Matrix::operator=(v); return *this; }
    using Matrix::operator();
    double& operator()(int i) { return d[(std::size_t)(i - 1)]; }
    double operator()(int i) const { return d[(std::size_t)(i - 1)]; }
    void ReSize(int n) { Matrix::ReSize(n, 1); }
    void resize(int n) { ReSize(n); }
    double Maximum() const { return *std::max_element(d.begin(), d.end()); }
    double Minimum() const { return *std::min_element(d.begin(), d.end()); }
};

class RowVector : public Matrix {
public:
    RowVector() = default;
    explicit RowVector(int n) : Matrix(1, n) {}
    RowVector(const Matrix& m) : Matrix(m) {
        if (nr != 1 && nc == 1) { std::swap(nr, nc); }
    }
    RowVector& operator=(const Matrix& m) { Matrix::operator=(m); if (nr != 1 && nc == 1) std::swap(nr, nc); return *this; }
    RowVector& operator=(double v) { Matrix::operator=(v); return *this; }
    using Matrix::operator();
    double& operator()(int i) { return d[(std::size_t)(i - 1)]; }
    double operator()(int i) const { return d[(std::size_t)(i - 1)]; }
    void ReSize(int n) { Matrix::ReSize(1, n); }
    void resize(int n) { ReSize(n); }
};

// Square matrix that only ever holds a diagonal (enough for the SVD call sites, which are
// off the hot path and never executed by the oracle driver).
class DiagonalMatrix : public Matrix {
public:
    DiagonalMatrix() = default;
    explicit DiagonalMatrix(int n) : Matrix(n, n) {}
    using Matrix::operator();
    double& operator()(int i) { return Matrix::operator()(i, i); }
    double operator()(int i) const { return Matrix::operator()(i, i); }
    void ReSize(int n) { Matrix::ReSize(n, n); }
};

class IdentityMatrix : public Matrix {
public:
    explicit IdentityMatrix(int n) : Matrix(n, n) { for (int i = 1; i <= n; ++i) (*this)(i, i) = 1.0; }
};

// Off the hot path (post-hoc strain maps, reg_tools.cpp:406,443). One-sided Jacobi SVD, singular values
// in descending order like NEWMAT's: A (m x n, m >= n) = U * D * V.t(), U m x n, D n x n, V n x n.
inline void SVD(const Matrix& A, DiagonalMatrix& D, Matrix& U, Matrix& V) {
    const int m = A.Nrows(), n = A.Ncols();
    if (m < n) throw std::runtime_error("shim NEWMAT: SVD needs rows >= cols");
    Matrix W = A;
    V = Matrix(n, n);
    for (int i = 1; i <= n; ++i) V(i, i) = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 1; p < n; ++p)
            for (int q = p + 1; q <= n; ++q) {
                double a = 0, b = 0, c = 0;
                for (int i = 1; i <= m; ++i) { a += W(i, p) * W(i, p); b += W(i, q) * W(i, q); c += W(i, p) * W(i, q); }
                if (std::fabs(c) <= 1e-300 || std::fabs(c) <= 1e-15 * std::sqrt(a * b)) continue;
                off += std::fabs(c);
                double zeta = (b - a) / (2.0 * c);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 1; i <= m; ++i) { double x = W(i, p), y = W(i, q); W(i, p) = cs * x - sn * y; W(i, q) = sn * x + cs * y; }
                for (int i = 1; i <= n; ++i) { double x = V(i, p), y = V(i, q); V(i, p) = cs * x - sn * y; V(i, q) = sn * x + cs * y; }
            }
        if (off == 0.0) break;
    }
    std::vector<double> sv(n);
    for (int j = 1; j <= n; ++j) { double s = 0; for (int i = 1; i <= m; ++i) s += W(i, j) * W(i, j); sv[j - 1] = std::sqrt(s); }
    std::vector<int> ord(n);
    for (int j = 0; j < n; ++j) ord[j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return sv[x] > sv[y]; });
    D.ReSize(n);
    U = Matrix(m, n);
    Matrix V2(n, n);
    for (int j = 1; j <= n; ++j) {
        int o = ord[j - 1] + 1;
        D(j) = sv[o - 1];
        for (int i = 1; i <= m; ++i) U(i, j) = sv[o - 1] > 0 ? W(i, o) / sv[o - 1] : 0.0;
        for (int i = 1; i <= n; ++i) V2(i, j) = V(i, o);
    }
    V = V2;
}
inline void SVD(const Matrix& A, DiagonalMatrix& D, Matrix& U) { Matrix V; SVD(A, D, U, V); }
inline void SVD(const Matrix& A, DiagonalMatrix& D) { Matrix U, V; SVD(A, D, U, V); }

inline std::ostream& operator<<(std::ostream& os, const Matrix& m) {
    for (int i = 1; i <= m.Nrows(); ++i) {
        for (int j = 1; j <= m.Ncols(); ++j) os << m(i, j) << ' ';
        os << '\n';
    }
    return os;
}

} // namespace NEWMAT

#endif
