// Test of include/grbda_cuda_casadi.hpp without CasADi: a stand-in with casadi::Function's instruction
// interface (work-vector form, slots reused, casadi's opcode numbering) holds the hip-differential constraint
// of Tello (reference: src/Robots/Tello.cpp:139-154) written the way CasADi's SX would lower it; the walker's
// grbda_phi_op program must evaluate to the same phi, and a model created from a schedule that carries it
// must hash like the product's own TelloWithArms hip cluster would evaluate (checked numerically).
#include <cmath>
#include <cstdio>
#include <functional>
#include <vector>
#include "grbda_cuda_casadi.hpp"

namespace mock
{
    // casadi::Operation values (casadi/core/calculus.hpp, 3.6.3) for the operations used here
    enum
    {
        OP_ASSIGN = 0,
        OP_ADD = 1,
        OP_SUB = 2,
        OP_MUL = 3,
        OP_DIV = 4,
        OP_NEG = 5,
        OP_SQ = 11,
        OP_TWICE = 12,
        OP_SIN = 13,
        OP_COS = 14,
        OP_CONST = 51,
        OP_INPUT = 52,
        OP_OUTPUT = 53
    };
    struct Instr
    {
        int id;
        std::vector<long long> in, out;
        double c;
    };
    // records a straight-line program while an expression is evaluated; work slots are recycled like
    // CasADi's live-variable allocation does
    struct Function
    {
        std::vector<Instr> ins;
        int n_work = 0;
        size_t sz_w() const { return (size_t)n_work; }
        long long n_instructions() const { return (long long)ins.size(); }
        int instruction_id(long long k) const { return ins[(size_t)k].id; }
        std::vector<long long> instruction_input(long long k) const { return ins[(size_t)k].in; }
        std::vector<long long> instruction_output(long long k) const { return ins[(size_t)k].out; }
        double instruction_constant(long long k) const { return ins[(size_t)k].c; }
    };
    struct Builder
    {
        Function f;
        std::vector<int> free_slots;
        int alloc()
        {
            if (!free_slots.empty())
            {
                const int s = free_slots.back();
                free_slots.pop_back();
                return s;
            }
            return f.n_work++;
        }
        void release(int s) { free_slots.push_back(s); }
        int constant(double v)
        {
            const int s = alloc();
            f.ins.push_back({OP_CONST, {}, {s}, v});
            return s;
        }
        int input(int i)
        {
            const int s = alloc();
            f.ins.push_back({OP_INPUT, {0, i}, {s}, 0});
            return s;
        }
        int unary(int op, int a)
        {
            const int s = alloc();
            f.ins.push_back({op, {a}, {s}, 0});
            return s;
        }
        int binary(int op, int a, int b)
        {
            const int s = alloc();
            f.ins.push_back({op, {a, b}, {s}, 0});
            return s;
        }
        void output(int row, int a) { f.ins.push_back({OP_OUTPUT, {a}, {0, row}, 0}); }
    };
} // namespace mock

// Tello hip differential (Tello.cpp:139-154; the integer-division terms are zero): rows for (y1, +) and (y2, -)
static double hipRow(const double *q, int row)
{
    const double ql1 = q[0], ql2 = q[1], y = q[2 + row] / 6.0, sgn = row == 0 ? 1.0 : -1.0;
    return 57. * std::sin(y) / 2500. - 49. * std::cos(ql1) / 5000. - sgn * 399. * std::sin(ql1) / 20000. -
           8. * std::cos(y) * std::cos(ql2) / 625. - 57. * std::cos(ql1) * std::sin(ql2) / 2500. -
           sgn * 7. * std::sin(y) * std::sin(ql1) / 625. + sgn * 7. * std::sin(ql1) * std::sin(ql2) / 625. -
           8. * std::cos(ql1) * std::sin(y) * std::sin(ql2) / 625.;
}

static double evalProgram(const grbda_bridge::PhiProgram &p, const double *q, int row)
{
    std::vector<double> v;
    for (const grbda_phi_op &o : p.ops)
        switch (o.op)
        {
        case 0: v.push_back(o.val); break;
        case 1: v.push_back(q[o.b]); break;
        case 2: v.push_back(v[o.a] + v[o.b]); break;
        case 3: v.push_back(v[o.a] - v[o.b]); break;
        case 4: v.push_back(v[o.a] * v[o.b]); break;
        case 5: v.push_back(v[o.a] / v[o.b]); break;
        case 6: v.push_back(-v[o.a]); break;
        case 7: v.push_back(std::sin(v[o.a])); break;
        case 8: v.push_back(std::cos(v[o.a])); break;
        }
    return v[p.outputs[row]];
}

int main()
{
    using namespace mock;
    Builder b;
    // the expression of hipRow(), lowered term by term with slot recycling; uses OP_NEG, OP_TWICE and OP_SQ too
    for (int row = 0; row < 2; row++)
    {
        const double sgn = row == 0 ? 1.0 : -1.0;
        const int ql1 = b.input(0), ql2 = b.input(1), yq = b.input(2 + row), six = b.constant(6.0);
        const int y = b.binary(OP_DIV, yq, six);
        b.release(yq), b.release(six);
        const int sy = b.unary(OP_SIN, y), cy = b.unary(OP_COS, y), s1 = b.unary(OP_SIN, ql1), c1 = b.unary(OP_COS, ql1),
                  s2 = b.unary(OP_SIN, ql2), c2 = b.unary(OP_COS, ql2);
        auto scaled = [&](int x, double num, double den) {
            const int k = b.constant(num), t = b.binary(OP_MUL, k, x), d = b.constant(den), r = b.binary(OP_DIV, t, d);
            b.release(k), b.release(t), b.release(d);
            return r;
        };
        int acc = scaled(sy, 57., 2500.);
        auto sub = [&](int term) {
            const int r = b.binary(OP_SUB, acc, term);
            b.release(acc), b.release(term);
            acc = r;
        };
        auto add = [&](int term) {
            const int r = b.binary(OP_ADD, acc, term);
            b.release(acc), b.release(term);
            acc = r;
        };
        sub(scaled(c1, 49., 5000.));
        sub(scaled(s1, sgn * 399., 20000.));
        {
            const int t = b.binary(OP_MUL, cy, c2);
            sub(scaled(t, 8., 625.));
            b.release(t);
        }
        {
            const int t = b.binary(OP_MUL, c1, s2);
            sub(scaled(t, 57., 2500.));
            b.release(t);
        }
        {
            const int t = b.binary(OP_MUL, sy, s1);
            sub(scaled(t, sgn * 7., 625.));
            b.release(t);
        }
        {
            const int t = b.binary(OP_MUL, s1, s2);
            add(scaled(t, sgn * 7., 625.));
            b.release(t);
        }
        {
            const int t = b.binary(OP_MUL, c1, sy), u = b.binary(OP_MUL, t, s2);
            sub(scaled(u, 8., 625.));
            b.release(t), b.release(u);
        }
        // + (x^2 - x*x) + (2x - (x + x)) - (-0): exercises OP_SQ / OP_TWICE / OP_NEG without changing the value
        {
            const int sq = b.unary(OP_SQ, sy), mm = b.binary(OP_MUL, sy, sy), d = b.binary(OP_SUB, sq, mm);
            add(d);
            const int tw = b.unary(OP_TWICE, c2), pp = b.binary(OP_ADD, c2, c2), e = b.binary(OP_SUB, tw, pp);
            add(e);
            const int z = b.constant(0.0), nz = b.unary(OP_NEG, z);
            sub(nz);
        }
        b.output(row, acc);
        for (int s : {ql1, ql2, y, sy, cy, s1, c1, s2, c2, acc})
            b.release(s);
    }
    const grbda_bridge::Opcodes oc{OP_CONST, OP_INPUT, OP_OUTPUT, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SIN, OP_COS, OP_SQ, OP_TWICE};
    const grbda_bridge::PhiProgram p = grbda_bridge::walkSXFunction(b.f, oc);
    if (p.outputs.size() != 2)
    {
        std::printf("expected two constraint rows\n");
        return 1;
    }
    double worst = 0;
    for (int t = 0; t < 100; t++)
    {
        double q[4];
        for (int i = 0; i < 4; i++)
            q[i] = std::sin(1.0 + 7.3 * t + 2.1 * i) * 1.5;
        for (int row = 0; row < 2; row++)
            worst = std::fmax(worst, std::fabs(evalProgram(p, q, row) - hipRow(q, row)));
    }
    std::printf("work slots %zu, instructions %lld, phi ops %zu, max |phi_program - phi| = %.2e\n", b.f.sz_w(),
                b.f.n_instructions(), p.ops.size(), worst);
    // hand the program to the library exactly as a reference-side binding would: a one-cluster schedule
    // (four revolute bodies on the ground, independent = {1, 1, 0, 0}) must be accepted as a valid model
    const int32_t parent[4] = {-1, -1, -1, -1}, axis[4] = {2, 2, 2, 2};
    double E[36] = {0}, r[12] = {0}, I[144] = {0};
    for (int bdy = 0; bdy < 4; bdy++)
    {
        E[9 * bdy] = E[9 * bdy + 4] = E[9 * bdy + 8] = 1.0;
        for (int d = 0; d < 6; d++)
            I[36 * bdy + 7 * d] = 1.0;
    }
    const uint8_t independent[4] = {1, 1, 0, 0};
    const int32_t type = GRBDA_CLUSTER_IMPLICIT, nbodies = 4, nind = 2, zero = 0, ncnstr = 2;
    const int32_t phi_count = (int32_t)p.ops.size();
    grbda_schedule s{};
    s.num_bodies = 4, s.num_clusters = 1;
    s.gravity[2] = -9.81;
    s.body_parent = parent, s.body_joint_axis = axis, s.body_xtree_E = E, s.body_xtree_r = r, s.body_inertia = I;
    s.body_independent = independent;
    s.cluster_type = &type, s.cluster_num_bodies = &nbodies, s.cluster_num_independent = &nind;
    s.cluster_G_offset = &zero, s.G_values = r, s.cluster_phi_offset = &zero, s.cluster_phi_count = &phi_count;
    s.cluster_phi_out_offset = &zero, s.cluster_num_constraints = &ncnstr;
    s.phi_ops = p.ops.data(), s.phi_outputs = p.outputs.data();
    grbda_model *m = nullptr;
    if (grbda_cuda_model_create(&s, -1, &m) != GRBDA_OK)
    {
        std::printf("model_create: %s\n", grbda_cuda_last_error_string());
        return 1;
    }
    const bool sizes_ok = grbda_cuda_num_positions(m) == 4 && grbda_cuda_num_degrees_of_freedom(m) == 2;
    int32_t sizes2[2] = {0, 0};
    grbda_cuda_cluster_phi(m, 0, nullptr, nullptr, nullptr, sizes2);
    grbda_cuda_model_destroy(m);
    const bool ok = worst < 1e-15 && sizes_ok && sizes2[0] == phi_count && sizes2[1] == 2;
    std::printf(ok ? "OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
