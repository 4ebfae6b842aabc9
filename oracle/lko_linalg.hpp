// oracle/lko_linalg.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal fixed-size dense algebra used by the CPU restatement so that its statements read like
// the reference's Eigen expressions. Products are evaluated coefficient-wise, left to right, the
// way Eigen's lazy small-matrix product does (sum over k ascending); compile with
// -ffp-contract=off so no FMA is formed (the reference is built -O3 for generic x86-64,
// legkilo/CMakeLists.txt:15, i.e. without FMA contraction).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

namespace lko {

template <int R, int C>
struct Mat {
    double a[R * C];
    Mat() { for (int i = 0; i < R * C; ++i) a[i] = 0.0; }
    static Mat Zero() { return Mat(); }
    static Mat Identity() {
        Mat m;
        for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
        return m;
    }
    double& operator()(int i, int j) { return a[i * C + j]; }
    double operator()(int i, int j) const { return a[i * C + j]; }
    double& operator[](int i) { return a[i]; }  // vectors
    double operator[](int i) const { return a[i]; }
    Mat<C, R> transpose() const {
        Mat<C, R> t;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    Mat operator+(const Mat& o) const {
        Mat r;
        for (int i = 0; i < R * C; ++i) r.a[i] = a[i] + o.a[i];
        return r;
    }
    Mat operator-(const Mat& o) const {
        Mat r;
        for (int i = 0; i < R * C; ++i) r.a[i] = a[i] - o.a[i];
        return r;
    }
    Mat operator-() const {
        Mat r;
        for (int i = 0; i < R * C; ++i) r.a[i] = -a[i];
        return r;
    }
    Mat operator*(double s) const {
        Mat r;
        for (int i = 0; i < R * C; ++i) r.a[i] = a[i] * s;
        return r;
    }
    Mat operator/(double s) const {
        Mat r;
        for (int i = 0; i < R * C; ++i) r.a[i] = a[i] / s;
        return r;
    }
    Mat& operator+=(const Mat& o) {
        for (int i = 0; i < R * C; ++i) a[i] += o.a[i];
        return *this;
    }
    double norm() const {
        double s = 0;
        for (int i = 0; i < R * C; ++i) s += a[i] * a[i];
        return std::sqrt(s);
    }
    double trace() const {
        double s = 0;
        for (int i = 0; i < (R < C ? R : C); ++i) s += (*this)(i, i);
        return s;
    }
    template <int BR, int BC>
    Mat<BR, BC> block(int r0, int c0) const {
        Mat<BR, BC> b;
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) b(i, j) = (*this)(r0 + i, c0 + j);
        return b;
    }
    template <int BR, int BC>
    void setBlock(int r0, int c0, const Mat<BR, BC>& b) {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) (*this)(r0 + i, c0 + j) = b(i, j);
    }
};

template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C>& m) {
    return m * s;
}

template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& x, const Mat<K, C>& y) {
    Mat<R, C> r;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            double s = x(i, 0) * y(0, j);
            for (int k = 1; k < K; ++k) s += x(i, k) * y(k, j);
            r(i, j) = s;
        }
    return r;
}

using V3 = Mat<3, 1>;
using M3 = Mat<3, 3>;
using V6 = Mat<6, 1>;
using M6 = Mat<6, 6>;

inline V3 vec3(double x, double y, double z) {
    V3 v;
    v[0] = x; v[1] = y; v[2] = z;
    return v;
}
inline double dot(const V3& x, const V3& y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
inline V3 cross(const V3& x, const V3& y) {
    return vec3(x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0]);
}
inline V3 normalized(const V3& v) { return v / v.norm(); }

// Cyclic Jacobi eigen-decomposition of a real symmetric 3x3 (stands in for
// Eigen::EigenSolver<Matrix3d>, legkilo/src/core/slam/voxel_map.cc:55-58; any convergent
// symmetric solver agrees to O(eps*||C||); eigenvector sign is free and the path is invariant
// to it). Returns eigenvalues w[3] (unordered) and unit eigenvectors as COLUMNS of V.
inline void eig_sym3(const M3& Ain, double w[3], M3& V) {
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = 0.5 * (Ain(i, j) + Ain(j, i));
    double Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        double diag = std::fabs(A[0][0]) + std::fabs(A[1][1]) + std::fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A * J
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T * A
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double qkp = Q[k][p], qkq = Q[k][q];
                    Q[k][p] = c * qkp - s * qkq;
                    Q[k][q] = s * qkp + c * qkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) {
        w[i] = A[i][i];
        double n = std::sqrt(Q[0][i] * Q[0][i] + Q[1][i] * Q[1][i] + Q[2][i] * Q[2][i]);
        for (int k = 0; k < 3; ++k) V(k, i) = Q[k][i] / n;
    }
}

// Dense row-major dynamic matrix helpers for the literal (measurement-space) Kalman gain.
struct DMat {
    int r = 0, c = 0;
    std::vector<double> a;
    DMat() {}
    DMat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

// In-place LU with partial pivoting (the algorithm of Eigen::PartialPivLU, which
// MatrixXd::inverse() uses, legkilo/src/core/slam/eskf.cc:109). perm[i] = source row of row i.
// Returns false when a pivot is exactly zero.
inline bool lu_factor(DMat& A, std::vector<int>& perm) {
    const int n = A.r;
    perm.resize(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::fabs(A(k, k));
        for (int i = k + 1; i < n; ++i) {
            double v = std::fabs(A(i, k));
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) return false;
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        const double inv = 1.0 / A(k, k);
        for (int i = k + 1; i < n; ++i) {
            double l = A(i, k) * inv;
            A(i, k) = l;
            if (l != 0.0) {
                double* ri = &A.a[(size_t)i * n];
                const double* rk = &A.a[(size_t)k * n];
                for (int j = k + 1; j < n; ++j) ri[j] -= l * rk[j];
            }
        }
    }
    return true;
}

// Solve A X = B for X (B is n x m, overwritten) given the factorisation above.
inline void lu_solve(const DMat& LU, const std::vector<int>& perm, DMat& B) {
    const int n = LU.r, m = B.c;
    DMat X(n, m);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) X(i, j) = B(perm[i], j);
    for (int i = 0; i < n; ++i) {  // forward, unit lower
        double* xi = &X.a[(size_t)i * m];
        for (int k = 0; k < i; ++k) {
            double l = LU(i, k);
            if (l == 0.0) continue;
            const double* xk = &X.a[(size_t)k * m];
            for (int j = 0; j < m; ++j) xi[j] -= l * xk[j];
        }
    }
    for (int i = n - 1; i >= 0; --i) {  // backward
        double* xi = &X.a[(size_t)i * m];
        for (int k = i + 1; k < n; ++k) {
            double u = LU(i, k);
            if (u == 0.0) continue;
            const double* xk = &X.a[(size_t)k * m];
            for (int j = 0; j < m; ++j) xi[j] -= u * xk[j];
        }
        double inv = 1.0 / LU(i, i);
        for (int j = 0; j < m; ++j) xi[j] *= inv;
    }
    B = X;
}

}  // namespace lko
