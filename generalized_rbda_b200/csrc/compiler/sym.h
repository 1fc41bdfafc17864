// Symbolic scalar for the device-side model compiler.
//
// The batched kernels are generated per model: the cluster algorithms (compiler/algorithms.h) are
// run once, on the host, over `Sym` scalars. Every arithmetic operation appends a node to an
// expression DAG with constant folding, algebraic identities (x*0, x*1, x+0, double negation ...)
// and hash-consing (common sub-expression elimination), so model constants (Xtree rotations that
// are identity or signed permutations, rotor inertias with zero COM, zeros of S and G, block
// diagonal I ...) disappear at compile time and the cluster-joint dispatch is resolved per model.
// The surviving straight-line program is emitted as a CUDA kernel body (compiler/emit.h).
//
// The reference does the same thing offline with CasADi SX to produce src/Codegen/*.cpp
// (reference: scripts/matlab/derive_revolute_with_rotor.m, include/grbda/Codegen/CasadiGen.h) and
// at run time for GenericImplicit constraints (src/Dynamics/ClusterJoints/GenericJoint.cpp:10-109).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace grbda
{
    namespace sym
    {
        enum Op : uint8_t
        {
            OP_CONST = 0,
            OP_INPUT, // a = input array id, b = element index
            OP_ADD,
            OP_SUB,
            OP_MUL,
            OP_DIV,
            OP_NEG,
            OP_SIN,
            OP_COS,
            OP_SQRT,
            OP_SELECT_GT, // (a > b) ? c : d   (d stored in `e`)
        };

        struct Node
        {
            Op op;
            int32_t a = -1, b = -1, c = -1, e = -1;
            double val = 0.0;
        };

        struct Graph
        {
            std::vector<Node> nodes;
            std::unordered_map<uint64_t, std::vector<int32_t>> index;

            static uint64_t hashNode(const Node &n)
            {
                uint64_t h = 1469598103934665603ull;
                auto mix = [&h](uint64_t x)
                {
                    h ^= x;
                    h *= 1099511628211ull;
                    h ^= h >> 29;
                };
                mix(n.op);
                mix((uint64_t)(uint32_t)n.a);
                mix((uint64_t)(uint32_t)n.b);
                mix((uint64_t)(uint32_t)n.c);
                mix((uint64_t)(uint32_t)n.e);
                uint64_t bits;
                std::memcpy(&bits, &n.val, 8);
                mix(bits);
                return h;
            }
            static bool same(const Node &x, const Node &y)
            {
                return x.op == y.op && x.a == y.a && x.b == y.b && x.c == y.c && x.e == y.e &&
                       std::memcmp(&x.val, &y.val, 8) == 0;
            }
            int32_t intern(const Node &n)
            {
                const uint64_t h = hashNode(n);
                auto &bucket = index[h];
                for (int32_t id : bucket)
                    if (same(nodes[id], n))
                        return id;
                nodes.push_back(n);
                bucket.push_back((int32_t)nodes.size() - 1);
                return (int32_t)nodes.size() - 1;
            }
            int32_t constant(double v)
            {
                if (v == 0.0)
                    v = 0.0; // -0.0 -> +0.0
                Node n;
                n.op = OP_CONST;
                n.val = v;
                return intern(n);
            }
            int32_t input(int array, int element)
            {
                Node n;
                n.op = OP_INPUT;
                n.a = array;
                n.b = element;
                return intern(n);
            }
        };

        // The graph currently being built (one model compilation at a time per thread).
        inline Graph *&currentGraph()
        {
            static thread_local Graph *g = nullptr;
            return g;
        }
        struct GraphScope
        {
            Graph *prev;
            explicit GraphScope(Graph &g) : prev(currentGraph()) { currentGraph() = &g; }
            ~GraphScope() { currentGraph() = prev; }
        };

        struct Sym
        {
            int32_t id;
            Sym() : id(G().constant(0.0)) {}
            Sym(double v) : id(G().constant(v)) {}
            Sym(int v) : id(G().constant((double)v)) {}
            static Sym fromId(int32_t i)
            {
                Sym s(Raw{});
                s.id = i;
                return s;
            }
            static Sym input(int array, int element) { return fromId(G().input(array, element)); }

            static Graph &G()
            {
                Graph *g = currentGraph();
                if (!g)
                    throw std::runtime_error("grbda::sym: no active Graph");
                return *g;
            }
            const Node &node() const { return G().nodes[id]; }
            bool isConst() const { return node().op == OP_CONST; }
            double constValue() const { return node().val; }
            bool isZero() const { return isConst() && node().val == 0.0; }
            bool isOne() const { return isConst() && node().val == 1.0; }
            bool isMinusOne() const { return isConst() && node().val == -1.0; }
            bool isNeg() const { return node().op == OP_NEG; }

        private:
            struct Raw
            {
            };
            explicit Sym(Raw) : id(-1) {}
        };

        inline Sym mk(Op op, int32_t a, int32_t b = -1, int32_t c = -1, int32_t e = -1)
        {
            Node n;
            n.op = op;
            n.a = a;
            n.b = b;
            n.c = c;
            n.e = e;
            return Sym::fromId(Sym::G().intern(n));
        }

        inline Sym operator-(const Sym &x);
        inline Sym operator+(const Sym &x, const Sym &y);
        inline Sym operator-(const Sym &x, const Sym &y);
        inline Sym operator*(const Sym &x, const Sym &y);

        inline Sym operator-(const Sym &x)
        {
            if (x.isConst())
                return Sym(-x.constValue());
            const Node &n = x.node();
            if (n.op == OP_NEG)
                return Sym::fromId(n.a);
            if (n.op == OP_SUB)
                return mk(OP_SUB, n.b, n.a);
            return mk(OP_NEG, x.id);
        }
        inline Sym operator+(const Sym &x, const Sym &y)
        {
            if (x.isConst() && y.isConst())
                return Sym(x.constValue() + y.constValue());
            if (x.isZero())
                return y;
            if (y.isZero())
                return x;
            if (y.isNeg())
                return x - Sym::fromId(y.node().a);
            if (x.isNeg())
                return y - Sym::fromId(x.node().a);
            if (x.id <= y.id)
                return mk(OP_ADD, x.id, y.id);
            return mk(OP_ADD, y.id, x.id);
        }
        inline Sym operator-(const Sym &x, const Sym &y)
        {
            if (x.isConst() && y.isConst())
                return Sym(x.constValue() - y.constValue());
            if (y.isZero())
                return x;
            if (x.isZero())
                return -y;
            if (x.id == y.id)
                return Sym(0.0);
            if (y.isNeg())
                return x + Sym::fromId(y.node().a);
            return mk(OP_SUB, x.id, y.id);
        }
        inline Sym operator*(const Sym &x, const Sym &y)
        {
            if (x.isConst() && y.isConst())
                return Sym(x.constValue() * y.constValue());
            if (x.isZero() || y.isZero())
                return Sym(0.0);
            if (x.isOne())
                return y;
            if (y.isOne())
                return x;
            if (x.isMinusOne())
                return -y;
            if (y.isMinusOne())
                return -x;
            // pull negations outward so that products are shared
            if (x.isNeg() && y.isNeg())
                return Sym::fromId(x.node().a) * Sym::fromId(y.node().a);
            if (x.isNeg())
                return -(Sym::fromId(x.node().a) * y);
            if (y.isNeg())
                return -(x * Sym::fromId(y.node().a));
            if (x.isConst() && x.constValue() < 0.0)
                return -(Sym(-x.constValue()) * y);
            if (y.isConst() && y.constValue() < 0.0)
                return -(x * Sym(-y.constValue()));
            if (x.id <= y.id)
                return mk(OP_MUL, x.id, y.id);
            return mk(OP_MUL, y.id, x.id);
        }
        inline Sym operator/(const Sym &x, const Sym &y)
        {
            if (x.isConst() && y.isConst())
                return Sym(x.constValue() / y.constValue());
            if (x.isZero())
                return Sym(0.0);
            if (y.isOne())
                return x;
            if (y.isConst())
                return x * Sym(1.0 / y.constValue());
            if (y.isNeg())
                return -(x / Sym::fromId(y.node().a));
            if (x.isNeg())
                return -(Sym::fromId(x.node().a) / y);
            return mk(OP_DIV, x.id, y.id);
        }
        inline Sym &operator+=(Sym &x, const Sym &y) { return x = x + y; }
        inline Sym &operator-=(Sym &x, const Sym &y) { return x = x - y; }
        inline Sym &operator*=(Sym &x, const Sym &y) { return x = x * y; }
        inline Sym &operator/=(Sym &x, const Sym &y) { return x = x / y; }

        inline Sym sin(const Sym &x)
        {
            if (x.isConst())
                return Sym(std::sin(x.constValue()));
            if (x.isNeg())
                return -mk(OP_SIN, x.node().a);
            return mk(OP_SIN, x.id);
        }
        inline Sym cos(const Sym &x)
        {
            if (x.isConst())
                return Sym(std::cos(x.constValue()));
            if (x.isNeg())
                return mk(OP_COS, x.node().a);
            return mk(OP_COS, x.id);
        }
        inline Sym sqrt(const Sym &x)
        {
            if (x.isConst())
                return Sym(std::sqrt(x.constValue()));
            return mk(OP_SQRT, x.id);
        }
        // (a > b) ? c : d
        inline Sym selectGt(const Sym &a, const Sym &b, const Sym &c, const Sym &d)
        {
            if (a.isConst() && b.isConst())
                return a.constValue() > b.constValue() ? c : d;
            if (c.id == d.id)
                return c;
            return mk(OP_SELECT_GT, a.id, b.id, c.id, d.id);
        }

    } // namespace sym
} // namespace grbda
