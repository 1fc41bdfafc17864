// Device-side model compiler, part 4: limb-parallel partition of a per-state program.
//
// One state per thread keeps the whole robot's working set in one thread (ABA on Tello: ~3 KB that
// must survive between the upward and the downward sweep), which forces 255 registers, heavy local
// memory traffic and 2 warps per scheduler. The partition below spreads ONE STATE OVER W WARPS:
// the cluster tree is cut below its trunk (the root cluster) into W limbs (the sub-trees hanging
// off the trunk); warp r of a CTA evaluates limb r for 32 states (lane = state), so every warp
// still runs divergence-free straight-line code, needs a fraction of the registers, and the CTA
// holds W x fewer bytes per thread. The limbs meet only at the trunk: each warp publishes the few
// values the trunk needs from its limb (ABA: the limb's articulated inertia and bias force seen
// from the trunk, 27 values; RNEA: 6; CRBA: 10) in shared memory, ONE named barrier, and every
// warp then finishes the (cheap) trunk redundantly and continues down its own limb.
//
// The cut is derived from the expression DAG, not hand-written per algorithm:
//   touch(n)  = set of limbs whose inputs node n depends on (trunk inputs / constants: empty set)
//   a node with touch = {r} is computed by warp r only; a node that mixes limbs is recomputed by
//   every warp that needs it; whenever warp r needs a node with touch = {r'} (r' != r) that node is
//   communicated through shared memory (a "frontier" value of limb r').
#pragma once
#include <algorithm>
#include <set>
#include <unordered_map>
#include "../host/model.h"
#include "emit.h"

namespace grbda
{
    namespace compiler
    {
        struct RolePlan
        {
            int W = 1;
            std::vector<int> cluster_role; // per cluster: limb index, -1 = trunk
        };

        // trunk = root cluster (when it has at least two children); limbs = its child sub-trees
        inline RolePlan planRoles(const ClusterTreeModel &m, int max_roles = 8)
        {
            RolePlan plan;
            const int Nc = m.getNumClusters();
            plan.cluster_role.assign(Nc, 0);
            std::vector<int> roots, kids;
            for (const ClusterTreeNode &c : m.clusters())
                if (c.parent_index_ < 0)
                    roots.push_back(c.index_);
            if (roots.size() != 1)
                return plan;
            for (const ClusterTreeNode &c : m.clusters())
                if (c.parent_index_ == roots[0])
                    kids.push_back(c.index_);
            if (kids.size() < 2 || (int)kids.size() > max_roles)
                return plan;
            plan.W = (int)kids.size();
            plan.cluster_role[roots[0]] = -1;
            for (const ClusterTreeNode &c : m.clusters())
            {
                if (c.index_ == roots[0])
                    continue;
                int top = c.index_;
                while (m.clusters()[top].parent_index_ != roots[0])
                    top = m.clusters()[top].parent_index_;
                plan.cluster_role[c.index_] = (int)(std::find(kids.begin(), kids.end(), top) - kids.begin());
            }
            return plan;
        }

        // One instruction of a role program (the same op set as the tape, plus communication)
        enum RoleOpKind
        {
            ROLE_NODE = 0,       // evaluate graph node `id`
            ROLE_COMM_STORE = 1, // publish node `id` in slot `slot`
            ROLE_BARRIER = 2,
            ROLE_COMM_LOAD = 3, // node `id` := slot `slot`
        };
        struct RoleOp
        {
            int kind;
            int32_t id;
            int slot;
        };

        struct RolePrograms
        {
            int W = 1;
            int num_slots = 0;
            bool has_barrier = false;
            std::vector<std::vector<RoleOp>> ops;                            // per role
            std::vector<std::vector<std::vector<std::pair<int, int>>>> outs; // per role: (array, element) per ...
            std::vector<ProgramStats> stats;                                 // per role
        };

        class RolePartitioner
        {
        public:
            // input_role(array, element) -> limb or -1 (trunk); output_role(array, element) -> role
            RolePartitioner(const sym::Graph &g, const Program &p, int W,
                            const std::function<int(int, int)> &input_role,
                            const std::function<int(int, int)> &output_role)
                : g_(g), p_(p), W_(W)
            {
                const size_t N = g.nodes.size();
                // touch masks
                touch_.assign(N, 0);
                for (size_t i = 0; i < N; i++)
                {
                    const sym::Node &n = g.nodes[i];
                    if (n.op == sym::OP_CONST)
                        continue;
                    if (n.op == sym::OP_INPUT)
                    {
                        const int r = input_role(n.a, n.b);
                        touch_[i] = r >= 0 ? (1u << r) : 0u;
                        continue;
                    }
                    uint32_t t = 0;
                    for (int32_t c : {n.a, n.b, n.c, n.e})
                        if (c >= 0)
                            t |= touch_[c];
                    touch_[i] = t;
                }
                // outputs per role
                role_outputs_.assign(W, {});
                for (size_t arr = 0; arr < p.outputs.size(); arr++)
                    for (size_t e = 0; e < p.outputs[arr].size(); e++)
                    {
                        int r = output_role((int)arr, (int)e);
                        if (r < 0)
                            r = 0;
                        role_outputs_[r].push_back({p.outputs[arr][e].id, (int)arr, (int)e});
                    }
                // needed sets; discover frontier nodes
                needed_.assign(W, std::vector<char>(N, 0));
                frontier_role_.assign(N, -1);
                for (int r = 0; r < W; r++)
                {
                    std::vector<int32_t> stack;
                    for (auto &o : role_outputs_[r])
                        stack.push_back(o.id);
                    traverse(r, stack);
                }
                // every role computes its own frontier values (their closure stays inside the limb)
                phase1_.assign(W, std::vector<char>(N, 0));
                for (int r = 0; r < W; r++)
                {
                    std::vector<int32_t> stack;
                    for (size_t i = 0; i < N; i++)
                        if (frontier_role_[i] == r)
                            stack.push_back((int32_t)i);
                    closure(r, stack, phase1_[r]);
                    for (size_t i = 0; i < N; i++)
                        if (phase1_[r][i])
                            needed_[r][i] = 1;
                }
                // slots
                slot_.assign(N, -1);
                int slots = 0;
                for (size_t i = 0; i < N; i++)
                    if (frontier_role_[i] >= 0)
                        slot_[i] = slots++;
                num_slots_ = slots;
            }

            RolePrograms build() const
            {
                RolePrograms rp;
                rp.W = W_;
                rp.num_slots = num_slots_;
                rp.has_barrier = num_slots_ > 0;
                rp.ops.assign(W_, {});
                const size_t N = g_.nodes.size();
                for (int r = 0; r < W_; r++)
                {
                    auto &ops = rp.ops[r];
                    // phase 1: everything the limb's published values depend on
                    for (size_t i = 0; i < N; i++)
                        if (phase1_[r][i] && frontier_role_[i] != r)
                            ops.push_back({ROLE_NODE, (int32_t)i, -1});
                        else if (phase1_[r][i])
                        {
                            ops.push_back({ROLE_NODE, (int32_t)i, -1});
                            ops.push_back({ROLE_COMM_STORE, (int32_t)i, slot_[i]});
                        }
                    if (rp.has_barrier)
                        ops.push_back({ROLE_BARRIER, -1, -1});
                    // phase 2: the rest, foreign frontier values come from shared memory
                    for (size_t i = 0; i < N; i++)
                    {
                        if (!needed_[r][i] || phase1_[r][i])
                            continue;
                        if (frontier_role_[i] >= 0 && frontier_role_[i] != r)
                            ops.push_back({ROLE_COMM_LOAD, (int32_t)i, slot_[i]});
                        else
                            ops.push_back({ROLE_NODE, (int32_t)i, -1});
                    }
                }
                return rp;
            }

            struct Out
            {
                int32_t id;
                int array, element;
            };
            const std::vector<std::vector<Out>> &roleOutputs() const { return role_outputs_; }

        private:
            static int singleBit(uint32_t m)
            {
                if (m == 0 || (m & (m - 1)) != 0)
                    return -1;
                int b = 0;
                while (!(m & 1u))
                {
                    m >>= 1;
                    b++;
                }
                return b;
            }
            // backward traversal for role r, stopping at nodes that belong to exactly one OTHER limb
            void traverse(int r, std::vector<int32_t> &stack)
            {
                while (!stack.empty())
                {
                    const int32_t i = stack.back();
                    stack.pop_back();
                    if (needed_[r][i])
                        continue;
                    needed_[r][i] = 1;
                    const sym::Node &n = g_.nodes[i];
                    if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                    {
                        // an input of another limb read directly (e.g. an output that copies it)
                        const int owner = singleBit(touch_[i]);
                        if (n.op == sym::OP_INPUT && owner >= 0 && owner != r)
                            frontier_role_[i] = owner;
                        continue;
                    }
                    const int owner = singleBit(touch_[i]);
                    if (owner >= 0 && owner != r)
                    {
                        frontier_role_[i] = owner; // communicated, not recomputed
                        continue;
                    }
                    for (int32_t c : {n.a, n.b, n.c, n.e})
                        if (c >= 0)
                            stack.push_back(c);
                }
            }
            void closure(int r, std::vector<int32_t> &stack, std::vector<char> &mark) const
            {
                while (!stack.empty())
                {
                    const int32_t i = stack.back();
                    stack.pop_back();
                    if (mark[i])
                        continue;
                    mark[i] = 1;
                    const sym::Node &n = g_.nodes[i];
                    if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                        continue;
                    for (int32_t c : {n.a, n.b, n.c, n.e})
                        if (c >= 0)
                            stack.push_back(c);
                }
                (void)r;
            }

            const sym::Graph &g_;
            const Program &p_;
            int W_;
            std::vector<uint32_t> touch_;
            std::vector<std::vector<Out>> role_outputs_;
            std::vector<std::vector<char>> needed_, phase1_;
            std::vector<int> frontier_role_;
            std::vector<int> slot_;
            int num_slots_ = 0;
        };

        // CUDA text of every role program. Macros used by the text (defined by the kernel shell):
        //   IN0/IN1/IN2(i), OUT0/OUT1/OUT2(i, x), KC(x), COMM_ST(slot, x), COMM_LD(slot), ROLE_BARRIER()
        class RoleEmitter
        {
        public:
            RoleEmitter(const sym::Graph &g, const RolePartitioner &part, const RolePrograms &rp,
                        ConstTable *consts = nullptr)
                : g_(g), part_(part), rp_(rp), consts_(consts) {}

            ProgramStats roleStats(int r) const
            {
                ProgramStats st;
                for (const RoleOp &op : rp_.ops[r])
                {
                    if (op.kind != ROLE_NODE)
                        continue;
                    st.n_nodes++;
                    switch (g_.nodes[op.id].op)
                    {
                    case sym::OP_ADD:
                    case sym::OP_SUB: st.n_add++; break;
                    case sym::OP_MUL: st.n_mul++; break;
                    case sym::OP_DIV: st.n_div++; break;
                    case sym::OP_SQRT: st.n_sqrt++; break;
                    case sym::OP_SIN: st.n_sin++; break;
                    case sym::OP_COS: st.n_cos++; break;
                    case sym::OP_NEG: st.n_neg++; break;
                    case sym::OP_INPUT: st.n_inputs++; break;
                    default: break;
                    }
                }
                return st;
            }

            std::string roleBody(int r) const
            {
                std::ostringstream os;
                const size_t N = g_.nodes.size();
                std::vector<char> defined(N, 0), loaded(N, 0);
                std::vector<std::vector<std::pair<int, int>>> stores(N);
                std::vector<std::pair<int32_t, std::pair<int, int>>> late; // CONST / NEG outputs
                for (auto &o : part_.roleOutputs()[r])
                {
                    const sym::Op op = g_.nodes[o.id].op;
                    if (op == sym::OP_CONST || op == sym::OP_NEG)
                        late.push_back({o.id, {o.array, o.element}});
                    else
                        stores[o.id].push_back({o.array, o.element});
                }
                // sin/cos partners within this role
                std::unordered_map<int32_t, int32_t> sin_of, cos_of;
                for (const RoleOp &op : rp_.ops[r])
                    if (op.kind == ROLE_NODE)
                    {
                        const sym::Node &n = g_.nodes[op.id];
                        if (n.op == sym::OP_SIN)
                            sin_of[n.a] = op.id;
                        else if (n.op == sym::OP_COS)
                            cos_of[n.a] = op.id;
                    }
                auto ref = [&](int32_t id) { return refOf(id, loaded); };
                auto emitStores = [&](int32_t id) {
                    for (auto &st : stores[id])
                        os << "OUT" << st.first << "(" << st.second << ", " << ref(id) << ");\n";
                };
                for (const RoleOp &op : rp_.ops[r])
                {
                    if (op.kind == ROLE_BARRIER)
                    {
                        os << "ROLE_BARRIER();\n";
                        continue;
                    }
                    if (op.kind == ROLE_COMM_STORE)
                    {
                        os << "COMM_ST(" << op.slot << ", " << ref(op.id) << ");\n";
                        continue;
                    }
                    if (op.kind == ROLE_COMM_LOAD)
                    {
                        os << "const real t" << op.id << " = COMM_LD(" << op.slot << ");\n";
                        loaded[op.id] = 1;
                        emitStores(op.id);
                        continue;
                    }
                    const int32_t i = op.id;
                    const sym::Node &n = g_.nodes[i];
                    if (n.op == sym::OP_CONST || n.op == sym::OP_NEG)
                        continue;
                    if (defined[i])
                    {
                        emitStores(i);
                        continue;
                    }
                    switch (n.op)
                    {
                    case sym::OP_INPUT:
                        os << "const real t" << i << " = IN" << n.a << "(" << n.b << ");\n";
                        break;
                    case sym::OP_ADD:
                        os << "const real t" << i << " = " << ref(n.a) << " + " << ref(n.b) << ";\n";
                        break;
                    case sym::OP_SUB:
                        os << "const real t" << i << " = " << ref(n.a) << " - " << ref(n.b) << ";\n";
                        break;
                    case sym::OP_MUL:
                        os << "const real t" << i << " = " << ref(n.a) << " * " << ref(n.b) << ";\n";
                        break;
                    case sym::OP_DIV:
                        os << "const real t" << i << " = GRBDA_DIV(" << ref(n.a) << ", " << ref(n.b) << ");\n";
                        break;
                    case sym::OP_SQRT:
                        os << "const real t" << i << " = sqrt(" << ref(n.a) << ");\n";
                        break;
                    case sym::OP_SELECT_GT:
                        os << "const real t" << i << " = (" << ref(n.a) << " > " << ref(n.b) << ") ? " << ref(n.c)
                           << " : " << ref(n.e) << ";\n";
                        break;
                    case sym::OP_SIN:
                    case sym::OP_COS:
                    {
                        auto &other_map = n.op == sym::OP_SIN ? cos_of : sin_of;
                        auto it = other_map.find(n.a);
                        if (it != other_map.end() && !defined[it->second])
                        {
                            const int32_t sN = n.op == sym::OP_SIN ? i : it->second;
                            const int32_t cN = n.op == sym::OP_SIN ? it->second : i;
                            os << "real t" << sN << ", t" << cN << "; grbda_sincos<false>(" << ref(n.a) << ", &t" << sN
                               << ", &t" << cN << ");\n";
                            defined[it->second] = 1;
                        }
                        else
                            os << "const real t" << i << " = " << (n.op == sym::OP_SIN ? "grbda_sin<false>(" : "grbda_cos<false>(") << ref(n.a)
                               << ");\n";
                        break;
                    }
                    default:
                        throw std::runtime_error("role emit: unknown op");
                    }
                    defined[i] = 1;
                    emitStores(i);
                }
                for (auto &l : late)
                    os << "OUT" << l.second.first << "(" << l.second.second << ", " << ref(l.first) << ");\n";
                return os.str();
            }

        private:
            std::string refOf(int32_t id, const std::vector<char> &loaded) const
            {
                const sym::Node &n = g_.nodes[id];
                if (loaded[id])
                    return "t" + std::to_string(id);
                if (n.op == sym::OP_CONST)
                {
                    if (consts_)
                        return consts_->ref(n.val);
                    char buf[64];
                    std::snprintf(buf, sizeof(buf), "KC(%.17g)", n.val);
                    return buf;
                }
                if (n.op == sym::OP_NEG)
                    return "(-" + refOf(n.a, loaded) + ")";
                return "t" + std::to_string(id);
            }
            const sym::Graph &g_;
            const RolePartitioner &part_;
            const RolePrograms &rp_;
            ConstTable *consts_;
        };

    } // namespace compiler
} // namespace grbda
