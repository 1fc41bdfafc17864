// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// First-order dual number x + eps dx: the CPU restatement instantiated on it evaluates exact directional
// derivatives of inverse / forward dynamics. The reference obtains the same Jacobians from CasADi's symbolic
// jacobian() of the model instantiated on casadi::SX
// (reference: UnitTests/testRigidBodyDynamicsAlgosDerivatives.cpp:126-155); CasADi 3.6.3 is not vendored under
// /root/reference, its published semantics (exact derivatives of the expression graph) are restated as
// forward-mode arithmetic, one sweep per input direction.
#pragma once
#include <cmath>
#include "scalar.h"

namespace grbda_oracle
{
    struct Dual
    {
        double v = 0.0, d = 0.0;
        Dual() {}
        Dual(double x) : v(x) {}
        Dual(int x) : v((double)x) {}
        Dual(double x, double dx) : v(x), d(dx) {}
    };
    inline Dual operator+(const Dual &a, const Dual &b) { return Dual(a.v + b.v, a.d + b.d); }
    inline Dual operator-(const Dual &a, const Dual &b) { return Dual(a.v - b.v, a.d - b.d); }
    inline Dual operator-(const Dual &a) { return Dual(-a.v, -a.d); }
    inline Dual operator*(const Dual &a, const Dual &b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
    inline Dual operator/(const Dual &a, const Dual &b)
    {
        const double x = a.v / b.v;
        return Dual(x, (a.d - x * b.d) / b.v);
    }
    inline Dual &operator+=(Dual &a, const Dual &b) { return a = a + b; }
    inline Dual &operator-=(Dual &a, const Dual &b) { return a = a - b; }
    inline Dual &operator*=(Dual &a, const Dual &b) { return a = a * b; }
    inline Dual &operator/=(Dual &a, const Dual &b) { return a = a / b; }
    inline bool operator<(const Dual &a, const Dual &b) { return a.v < b.v; }
    inline bool operator>(const Dual &a, const Dual &b) { return a.v > b.v; }
    inline bool operator<=(const Dual &a, const Dual &b) { return a.v <= b.v; }
    inline bool operator>=(const Dual &a, const Dual &b) { return a.v >= b.v; }
    inline Dual sqrt(const Dual &a)
    {
        const double s = std::sqrt(a.v);
        return Dual(s, a.d / (2.0 * s));
    }
    inline Dual sin(const Dual &a) { return Dual(std::sin(a.v), std::cos(a.v) * a.d); }
    inline Dual cos(const Dual &a) { return Dual(std::cos(a.v), -std::sin(a.v) * a.d); }
    inline Dual fabs(const Dual &a) { return a.v < 0.0 ? -a : a; }
    inline double to_double(const Dual &x) { return x.v; }
} // namespace grbda_oracle
