// Device-side model compiler: forward-mode differentiation of an expression DAG.
//
// The reference has no hand-written derivative algorithms: its "analytical derivatives" are CasADi's
// jacobian() of the symbolic inverse dynamics with respect to a tangent-space perturbation dq, the velocities
// and the third argument (reference: UnitTests/testRigidBodyDynamicsAlgosDerivatives.cpp:126-155, perturbation
// q (+) dq: UnitTests/testHelpers.hpp:50-112). The compiler already holds the same symbolic program, so the
// derivative programs are produced the same way: one tangent sweep per input direction over the DAG, through
// the Sym operators - constant folding and hash-consing remove every term whose tangent is structurally zero
// (a joint only moves its own subtree; a force only travels to its ancestors), which is the sparsity the
// published O(N d) derivative algorithms exploit by hand.
#pragma once
#include <unordered_map>
#include <vector>
#include "sym.h"

namespace grbda
{
    namespace compiler
    {
        // Tangents of `roots` along ONE direction. `seed`: tangent of input nodes (node id -> expression);
        // inputs that are not listed have tangent zero. Nodes created by the sweep are appended to the
        // current graph; only nodes that existed before the call are differentiated.
        inline std::vector<sym::Sym> tangentSweep(const std::vector<sym::Sym> &roots,
                                                  const std::unordered_map<int32_t, sym::Sym> &seed)
        {
            using sym::Sym;
            sym::Graph &g = Sym::G();
            const int32_t n0 = (int32_t)g.nodes.size();
            // nodes the roots depend on (ids are topologically ordered: operands precede their users)
            std::vector<char> live(n0, 0);
            for (const Sym &r : roots)
                live[r.id] = 1;
            for (int32_t i = n0 - 1; i >= 0; i--)
            {
                if (!live[i])
                    continue;
                const sym::Node n = g.nodes[i];
                if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                    continue;
                for (int32_t o : {n.a, n.b, n.c, n.e})
                    if (o >= 0)
                        live[o] = 1;
            }
            const int32_t zero = g.constant(0.0);
            std::vector<int32_t> tan(n0, zero);
            auto T = [&](int32_t id) { return Sym::fromId(tan[id]); };
            auto V = [&](int32_t id) { return Sym::fromId(id); };
            for (int32_t i = 0; i < n0; i++)
            {
                if (!live[i])
                    continue;
                const sym::Node n = g.nodes[i]; // copy: the vector grows below
                Sym t(0.0);
                switch (n.op)
                {
                case sym::OP_CONST:
                    break;
                case sym::OP_INPUT:
                {
                    auto it = seed.find(i);
                    if (it != seed.end())
                        t = it->second;
                    break;
                }
                case sym::OP_ADD:
                    t = T(n.a) + T(n.b);
                    break;
                case sym::OP_SUB:
                    t = T(n.a) - T(n.b);
                    break;
                case sym::OP_NEG:
                    t = -T(n.a);
                    break;
                case sym::OP_MUL:
                    t = T(n.a) * V(n.b) + V(n.a) * T(n.b);
                    break;
                case sym::OP_DIV: // x = a / b: dx = (da - x db) / b
                    if (!(T(n.a).isZero() && T(n.b).isZero()))
                        t = (T(n.a) - V(i) * T(n.b)) / V(n.b);
                    break;
                case sym::OP_SIN:
                    if (!T(n.a).isZero())
                        t = sym::cos(V(n.a)) * T(n.a);
                    break;
                case sym::OP_COS:
                    if (!T(n.a).isZero())
                        t = -(sym::sin(V(n.a)) * T(n.a));
                    break;
                case sym::OP_SQRT: // x = sqrt(a): dx = da / (2 x)
                    if (!T(n.a).isZero())
                        t = T(n.a) / (Sym(2.0) * V(i));
                    break;
                case sym::OP_SELECT_GT: // the condition is piecewise constant
                    t = sym::selectGt(V(n.a), V(n.b), T(n.c), T(n.e));
                    break;
                }
                tan[i] = t.id;
            }
            std::vector<Sym> out;
            out.reserve(roots.size());
            for (const Sym &r : roots)
                out.push_back(T(r.id));
            return out;
        }

        // `roots` with some input nodes replaced by expressions (node id -> expression): the partial derivatives of
        // a program are taken with placeholder inputs held fixed, which are then given their values.
        inline std::vector<sym::Sym> substituteInputs(const std::vector<sym::Sym> &roots,
                                                      const std::unordered_map<int32_t, sym::Sym> &values)
        {
            using sym::Sym;
            sym::Graph &g = Sym::G();
            const int32_t n0 = (int32_t)g.nodes.size();
            std::vector<char> live(n0, 0);
            for (const Sym &r : roots)
                live[r.id] = 1;
            for (int32_t i = n0 - 1; i >= 0; i--)
            {
                if (!live[i])
                    continue;
                const sym::Node n = g.nodes[i];
                if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                    continue;
                for (int32_t o : {n.a, n.b, n.c, n.e})
                    if (o >= 0)
                        live[o] = 1;
            }
            std::vector<int32_t> sub(n0, -1);
            auto S = [&](int32_t id) { return Sym::fromId(sub[id]); };
            for (int32_t i = 0; i < n0; i++)
            {
                if (!live[i])
                    continue;
                const sym::Node n = g.nodes[i];
                Sym x = Sym::fromId(i);
                switch (n.op)
                {
                case sym::OP_CONST:
                    break;
                case sym::OP_INPUT:
                {
                    auto it = values.find(i);
                    if (it != values.end())
                        x = it->second;
                    break;
                }
                case sym::OP_ADD:
                    x = S(n.a) + S(n.b);
                    break;
                case sym::OP_SUB:
                    x = S(n.a) - S(n.b);
                    break;
                case sym::OP_NEG:
                    x = -S(n.a);
                    break;
                case sym::OP_MUL:
                    x = S(n.a) * S(n.b);
                    break;
                case sym::OP_DIV:
                    x = S(n.a) / S(n.b);
                    break;
                case sym::OP_SIN:
                    x = sym::sin(S(n.a));
                    break;
                case sym::OP_COS:
                    x = sym::cos(S(n.a));
                    break;
                case sym::OP_SQRT:
                    x = sym::sqrt(S(n.a));
                    break;
                case sym::OP_SELECT_GT:
                    x = sym::selectGt(S(n.a), S(n.b), S(n.c), S(n.e));
                    break;
                }
                sub[i] = x.id;
            }
            std::vector<Sym> out;
            out.reserve(roots.size());
            for (const Sym &r : roots)
                out.push_back(S(r.id));
            return out;
        }
    } // namespace compiler
} // namespace grbda
