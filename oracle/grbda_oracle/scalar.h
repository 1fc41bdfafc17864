// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call anything under oracle/.
//
// Scalar helpers for the CPU restatement of generalized_rbda's ClusterTreeModel hot path.
// The restatement is templated on the scalar so that it can run with
//   * double                     — the parity oracle,
//   * grbda_oracle::Counter      — the same arithmetic, counting floating-point operations; this
//                                  is the definition of F_alg used by bench.py's roofline.
// The reference instantiates the same code for double / float / casadi::SX
// (reference: src/Dynamics/TreeModel.cpp:265-267).
#pragma once
#include <cmath>
#include <cstdint>

namespace grbda_oracle
{

    // Global operation counters filled by Counter arithmetic.
    struct OpCounts
    {
        // "all": every add/sub/mul/div/sqrt executed by the dense restatement
        // "alg": the same minus operations whose operand is a structural constant 0 or 1
        //        (a value that is 0/1 for every state: zeros of S, G, block-diagonal I, identity
        //        rotations ...). This is F_alg.
        uint64_t add_all = 0, mul_all = 0, div_all = 0, sqrt_all = 0, trig_all = 0;
        uint64_t add_alg = 0, mul_alg = 0, div_alg = 0, sqrt_alg = 0, trig_alg = 0;
        void reset() { *this = OpCounts(); }
        uint64_t flops_all() const { return add_all + mul_all + div_all + sqrt_all; }
        uint64_t flops_alg() const { return add_alg + mul_alg + div_alg + sqrt_alg; }
    };

    inline OpCounts &op_counts()
    {
        static thread_local OpCounts c;
        return c;
    }

    // Counting scalar. `k` tags structural constants: 0 = state dependent, 1 = structural zero,
    // 2 = structural one, 3 = other structural constant (model parameter).
    struct Counter
    {
        double v = 0.0;
        unsigned char k = 1;
        Counter() {}
        Counter(double x) : v(x), k(x == 0.0 ? 1 : (x == 1.0 ? 2 : 3)) {}
        Counter(int x) : Counter((double)x) {}
        static Counter variable(double x)
        {
            Counter c;
            c.v = x;
            c.k = 0;
            return c;
        }
        bool is_zero() const { return k == 1; }
        bool is_one() const { return k == 2; }
        bool is_const() const { return k != 0; }
    };

    inline Counter mk(double v, bool is_const)
    {
        if (is_const)
            return Counter(v);
        return Counter::variable(v);
    }

    inline Counter operator+(const Counter &a, const Counter &b)
    {
        auto &c = op_counts();
        c.add_all++;
        if (!(a.is_zero() || b.is_zero() || (a.is_const() && b.is_const())))
            c.add_alg++;
        return mk(a.v + b.v, a.is_const() && b.is_const());
    }
    inline Counter operator-(const Counter &a, const Counter &b)
    {
        auto &c = op_counts();
        c.add_all++;
        if (!(a.is_zero() || b.is_zero() || (a.is_const() && b.is_const())))
            c.add_alg++;
        return mk(a.v - b.v, a.is_const() && b.is_const());
    }
    inline Counter operator-(const Counter &a) { return mk(-a.v, a.is_const()); }
    inline Counter operator*(const Counter &a, const Counter &b)
    {
        auto &c = op_counts();
        c.mul_all++;
        if (a.is_zero() || b.is_zero())
            return Counter(0.0);
        bool cc = a.is_const() && b.is_const();
        if (!(cc || a.is_one() || b.is_one() || (a.is_const() && a.v == -1.0) ||
              (b.is_const() && b.v == -1.0)))
            c.mul_alg++;
        return mk(a.v * b.v, cc);
    }
    inline Counter operator/(const Counter &a, const Counter &b)
    {
        auto &c = op_counts();
        c.div_all++;
        if (a.is_zero())
            return Counter(0.0);
        bool cc = a.is_const() && b.is_const();
        if (!(cc || b.is_one()))
            c.div_alg++;
        return mk(a.v / b.v, cc);
    }
    inline Counter &operator+=(Counter &a, const Counter &b) { return a = a + b; }
    inline Counter &operator-=(Counter &a, const Counter &b) { return a = a - b; }
    inline Counter &operator*=(Counter &a, const Counter &b) { return a = a * b; }
    inline Counter &operator/=(Counter &a, const Counter &b) { return a = a / b; }
    inline bool operator<(const Counter &a, const Counter &b) { return a.v < b.v; }
    inline bool operator>(const Counter &a, const Counter &b) { return a.v > b.v; }
    inline bool operator<=(const Counter &a, const Counter &b) { return a.v <= b.v; }
    inline bool operator>=(const Counter &a, const Counter &b) { return a.v >= b.v; }
    inline Counter sqrt(const Counter &a)
    {
        auto &c = op_counts();
        c.sqrt_all++;
        if (!a.is_const())
            c.sqrt_alg++;
        return mk(std::sqrt(a.v), a.is_const());
    }
    inline Counter sin(const Counter &a)
    {
        auto &c = op_counts();
        c.trig_all++;
        if (!a.is_const())
            c.trig_alg++;
        return mk(std::sin(a.v), a.is_const());
    }
    inline Counter cos(const Counter &a)
    {
        auto &c = op_counts();
        c.trig_all++;
        if (!a.is_const())
            c.trig_alg++;
        return mk(std::cos(a.v), a.is_const());
    }
    inline Counter fabs(const Counter &a) { return mk(std::fabs(a.v), a.is_const()); }

    inline double to_double(double x) { return x; }
    inline double to_double(float x) { return x; }
    inline double to_double(const Counter &x) { return x.v; }

    // Mark a value as state dependent (inputs q, yd, ydd/tau).
    template <typename T>
    inline T as_variable(double x) { return T(x); }
    template <>
    inline Counter as_variable<Counter>(double x) { return Counter::variable(x); }

} // namespace grbda_oracle
