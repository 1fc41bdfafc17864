// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// Minimal dense dynamic matrix type standing in for the Eigen types the reference uses
// (reference: include/grbda/Utils/cppTypes.h:17-83 — DMat, DVec, Mat3, Mat6, SVec ...).
// Row-major storage; a vector is an n x 1 matrix.
#pragma once
#include <cassert>
#include <cmath>
#include <stdexcept>
#include <vector>
#include "scalar.h"

namespace grbda_oracle
{
    using std::cos;
    using std::fabs;
    using std::sin;
    using std::sqrt;

    template <typename T>
    struct Mat
    {
        int r = 0, c = 0;
        std::vector<T> a;

        Mat() {}
        Mat(int rows, int cols) : r(rows), c(cols), a((size_t)rows * cols, T(0.0)) {}
        static Mat Zero(int rows, int cols = 1) { return Mat(rows, cols); }
        static Mat Identity(int n)
        {
            Mat m(n, n);
            for (int i = 0; i < n; i++)
                m(i, i) = T(1.0);
            return m;
        }
        int rows() const { return r; }
        int cols() const { return c; }
        int size() const { return r * c; }
        T &operator()(int i, int j) { return a[(size_t)i * c + j]; }
        const T &operator()(int i, int j) const { return a[(size_t)i * c + j]; }
        T &operator[](int i) { return a[i]; }
        const T &operator[](int i) const { return a[i]; }

        Mat block(int i0, int j0, int nr, int nc) const
        {
            Mat m(nr, nc);
            for (int i = 0; i < nr; i++)
                for (int j = 0; j < nc; j++)
                    m(i, j) = (*this)(i0 + i, j0 + j);
            return m;
        }
        void setBlock(int i0, int j0, const Mat &m)
        {
            for (int i = 0; i < m.r; i++)
                for (int j = 0; j < m.c; j++)
                    (*this)(i0 + i, j0 + j) = m(i, j);
        }
        void addBlock(int i0, int j0, const Mat &m)
        {
            for (int i = 0; i < m.r; i++)
                for (int j = 0; j < m.c; j++)
                    (*this)(i0 + i, j0 + j) = (*this)(i0 + i, j0 + j) + m(i, j);
        }
        Mat segment(int i0, int n) const { return block(i0, 0, n, 1); }
        void setSegment(int i0, const Mat &v) { setBlock(i0, 0, v); }
        void addSegment(int i0, const Mat &v) { addBlock(i0, 0, v); }
        Mat col(int j) const { return block(0, j, r, 1); }
        Mat transpose() const
        {
            Mat m(c, r);
            for (int i = 0; i < r; i++)
                for (int j = 0; j < c; j++)
                    m(j, i) = (*this)(i, j);
            return m;
        }
        void setZero()
        {
            for (auto &x : a)
                x = T(0.0);
        }
        double norm() const
        {
            double s = 0;
            for (auto &x : a)
                s += to_double(x) * to_double(x);
            return std::sqrt(s);
        }
    };

    template <typename T>
    using Vec = Mat<T>;

    template <typename T>
    Mat<T> operator+(const Mat<T> &A, const Mat<T> &B)
    {
        assert(A.r == B.r && A.c == B.c);
        Mat<T> m(A.r, A.c);
        for (size_t i = 0; i < A.a.size(); i++)
            m.a[i] = A.a[i] + B.a[i];
        return m;
    }
    template <typename T>
    Mat<T> operator-(const Mat<T> &A, const Mat<T> &B)
    {
        assert(A.r == B.r && A.c == B.c);
        Mat<T> m(A.r, A.c);
        for (size_t i = 0; i < A.a.size(); i++)
            m.a[i] = A.a[i] - B.a[i];
        return m;
    }
    template <typename T>
    Mat<T> operator-(const Mat<T> &A)
    {
        Mat<T> m(A.r, A.c);
        for (size_t i = 0; i < A.a.size(); i++)
            m.a[i] = -A.a[i];
        return m;
    }
    template <typename T>
    Mat<T> operator*(const Mat<T> &A, const Mat<T> &B)
    {
        if (A.c != B.r)
            throw std::runtime_error("oracle Mat product: dimension mismatch");
        Mat<T> m(A.r, B.c);
        for (int i = 0; i < A.r; i++)
            for (int j = 0; j < B.c; j++)
            {
                T s = T(0.0);
                for (int k = 0; k < A.c; k++)
                    s = s + A(i, k) * B(k, j);
                m(i, j) = s;
            }
        return m;
    }
    template <typename T>
    Mat<T> operator*(const T &s, const Mat<T> &A)
    {
        Mat<T> m(A.r, A.c);
        for (size_t i = 0; i < A.a.size(); i++)
            m.a[i] = s * A.a[i];
        return m;
    }
    template <typename T>
    Mat<T> operator*(const Mat<T> &A, const T &s) { return s * A; }

    template <typename T>
    Mat<T> vec3(const T &x, const T &y, const T &z)
    {
        Mat<T> v(3, 1);
        v[0] = x;
        v[1] = y;
        v[2] = z;
        return v;
    }

    template <typename T>
    Mat<T> mat3(std::initializer_list<double> l)
    {
        Mat<T> m(3, 3);
        int i = 0;
        for (double x : l)
            m.a[i++] = T(x);
        return m;
    }

    // reference: include/grbda/Utils/OrientationTools.h:135-143 (vectorToSkewMat)
    template <typename T>
    Mat<T> skew(const Mat<T> &v)
    {
        Mat<T> m(3, 3);
        m(0, 1) = -v[2];
        m(0, 2) = v[1];
        m(1, 0) = v[2];
        m(1, 2) = -v[0];
        m(2, 0) = -v[1];
        m(2, 1) = v[0];
        return m;
    }

    // reference: include/grbda/Utils/OrientationTools.h:148-155 (matToSkewVec)
    template <typename T>
    Mat<T> matToSkewVec(const Mat<T> &m)
    {
        return T(0.5) * vec3<T>(m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1));
    }

    // Solve A X = B for square regular A. The reference applies D^-1 and K_d^-1 through
    // Eigen::ColPivHouseholderQR (include/grbda/Utils/Utilities.h:325-329,
    // src/Dynamics/Nodes/ClusterTreeNode.cpp:34-37). Eigen is not vendored under /root/reference;
    // the published algorithm is an exact solve of a small regular system, restated here as
    // Gaussian elimination with partial pivoting (agrees with any backward-stable solver to ~1e-13
    // on the well-conditioned systems the parity tests use).
    template <typename T>
    Mat<T> solve(Mat<T> A, Mat<T> B)
    {
        const int n = A.r;
        if (A.c != n || B.r != n)
            throw std::runtime_error("oracle solve: dimension mismatch");
        for (int k = 0; k < n; k++)
        {
            int p = k;
            double best = std::fabs(to_double(A(k, k)));
            for (int i = k + 1; i < n; i++)
                if (std::fabs(to_double(A(i, k))) > best)
                {
                    best = std::fabs(to_double(A(i, k)));
                    p = i;
                }
            if (best == 0.0)
                throw std::runtime_error("oracle solve: singular matrix");
            if (p != k)
            {
                for (int j = 0; j < n; j++)
                    std::swap(A(k, j), A(p, j));
                for (int j = 0; j < B.c; j++)
                    std::swap(B(k, j), B(p, j));
            }
            for (int i = k + 1; i < n; i++)
            {
                T f = A(i, k) / A(k, k);
                for (int j = k + 1; j < n; j++)
                    A(i, j) = A(i, j) - f * A(k, j);
                A(i, k) = T(0.0);
                for (int j = 0; j < B.c; j++)
                    B(i, j) = B(i, j) - f * B(k, j);
            }
        }
        for (int j = 0; j < B.c; j++)
            for (int i = n - 1; i >= 0; i--)
            {
                T s = B(i, j);
                for (int k = i + 1; k < n; k++)
                    s = s - A(i, k) * B(k, j);
                B(i, j) = s / A(i, i);
            }
        return B;
    }

    // 2-norm condition estimate of a small matrix through the Frobenius norms of A and A^-1
    // (an upper bound within a factor n; only used to reject near-singular K_d in state generation).
    inline double cond_estimate(const Mat<double> &A)
    {
        Mat<double> Ainv = solve(A, Mat<double>::Identity(A.r));
        return A.norm() * Ainv.norm();
    }

} // namespace grbda_oracle
