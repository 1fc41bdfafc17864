// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// Second-order truncated Taylor scalar x(t) = x0 + x1 t + x2 t^2 used to differentiate the implicit
// loop-constraint functions phi(q). The reference obtains K = dphi/dq and k = -Kdot*qd from CasADi
// symbolic autodiff (src/Dynamics/ClusterJoints/GenericJoint.cpp:58-68: jacobian(), jtimes()).
// CasADi 3.6.3 is not vendored under /root/reference; its published semantics (exact derivatives
// of the expression graph) are restated with forward-mode Taylor arithmetic:
//   phi(q + t d) = phi + (K d) t + (1/2 d^T H d) t^2
// so column j of K is the t-coefficient along d = e_j, and k = -(qd^T H qd) = -2 * (t^2-coefficient
// along d = qd).
#pragma once
#include "linalg.h"

namespace grbda_oracle
{
    template <typename T>
    struct Taylor2
    {
        T c0, c1, c2;
        Taylor2() : c0(0.0), c1(0.0), c2(0.0) {}
        Taylor2(double x) : c0(x), c1(0.0), c2(0.0) {}
        Taylor2(const T &a, const T &b, const T &c) : c0(a), c1(b), c2(c) {}
    };

    template <typename T>
    Taylor2<T> operator+(const Taylor2<T> &a, const Taylor2<T> &b)
    {
        return Taylor2<T>(a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2);
    }
    template <typename T>
    Taylor2<T> operator-(const Taylor2<T> &a, const Taylor2<T> &b)
    {
        return Taylor2<T>(a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2);
    }
    template <typename T>
    Taylor2<T> operator-(const Taylor2<T> &a) { return Taylor2<T>(-a.c0, -a.c1, -a.c2); }
    template <typename T>
    Taylor2<T> operator*(const Taylor2<T> &a, const Taylor2<T> &b)
    {
        return Taylor2<T>(a.c0 * b.c0, a.c0 * b.c1 + a.c1 * b.c0,
                          a.c0 * b.c2 + a.c1 * b.c1 + a.c2 * b.c0);
    }
    template <typename T>
    Taylor2<T> operator*(double s, const Taylor2<T> &a)
    {
        return Taylor2<T>(T(s) * a.c0, T(s) * a.c1, T(s) * a.c2);
    }
    template <typename T>
    Taylor2<T> operator*(const Taylor2<T> &a, double s) { return s * a; }
    template <typename T>
    Taylor2<T> operator/(const Taylor2<T> &a, double s)
    {
        return Taylor2<T>(a.c0 / T(s), a.c1 / T(s), a.c2 / T(s));
    }
    template <typename T>
    Taylor2<T> operator+(const Taylor2<T> &a, double s) { return Taylor2<T>(a.c0 + T(s), a.c1, a.c2); }
    template <typename T>
    Taylor2<T> operator-(const Taylor2<T> &a, double s) { return Taylor2<T>(a.c0 - T(s), a.c1, a.c2); }
    template <typename T>
    Taylor2<T> operator+(double s, const Taylor2<T> &a) { return a + s; }
    template <typename T>
    Taylor2<T> sin(const Taylor2<T> &a)
    {
        T s = sin(a.c0), c = cos(a.c0);
        return Taylor2<T>(s, c * a.c1, c * a.c2 - T(0.5) * s * a.c1 * a.c1);
    }
    template <typename T>
    Taylor2<T> cos(const Taylor2<T> &a)
    {
        T s = sin(a.c0), c = cos(a.c0);
        return Taylor2<T>(c, -(s * a.c1), -(s * a.c2) - T(0.5) * c * a.c1 * a.c1);
    }
    template <typename T>
    double to_double(const Taylor2<T> &a) { return to_double(a.c0); }

} // namespace grbda_oracle
