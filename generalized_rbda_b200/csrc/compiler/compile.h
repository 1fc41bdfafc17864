// Device-side model compiler, part 3: driver. Model + algorithm -> emitted CUDA body, tape, counts.
#pragma once
#include <fstream>
#include <ostream>
#include <sstream>
#include "algorithms.h"
#include "emit.h"
#include "partition.h"

namespace grbda
{
    namespace compiler
    {
        enum Algo
        {
            ALGO_ID = 0,      // in: q, yd, ydd   out: tau[nv]
            ALGO_FD = 1,      // in: q, yd, tau   out: ydd[nv]
            ALGO_FK = 2,      // in: q, yd        out: p[3 Nb], R[9 Nb], v[6 Nb]
            ALGO_H = 3,       // in: q            out: H[nv nv]
            ALGO_PHI = 4,     // in: q            out: phi[sum nc], Kd (row-major per implicit cluster)
            // generalized force of external forces on the model's terminal links (ModelCompiler::
            // externalForceBodies): in: q, f_ext[6 nf] (world frame, TreeModel::setExternalForces), tau_in
            ALGO_GFA = 5,     // out: tau_in + J^T f_ext   (forward dynamics with external forces: FD(.., tau + J^T f))
            ALGO_GFS = 6,     // out: tau_in - J^T f_ext   (inverse dynamics with external forces: ID(..) - J^T f)
            ALGO_COUNT = 7,   // entry points / registry slots
            // alternative programs of an entry point (selected per kernel variant, same I/O as the entry)
            PROGRAM_FD_LTL = 7, // forward dynamics as H^-1 (tau - C): CRBA + RNEA bias + sparse L^T D L
            PROGRAM_COUNT = 8
        };
        // registry slot a program belongs to
        inline int algoOfProgram(int program) { return program == PROGRAM_FD_LTL ? ALGO_FD : program; }
        inline const char *algoName(int a)
        {
            static const char *names[] = {"id", "fd", "fk", "h", "phi", "gfa", "gfs", "fd_ltl"};
            return names[a];
        }

        struct CompiledAlgo
        {
            std::string name;
            int n_in[3] = {0, 0, 0};
            int n_out[3] = {0, 0, 0};
            std::string body;
            std::string range_check; // expression: may the fast sin/cos forms be used for this state
            bool parked = false;     // body parks long-lived values in the thread's shared-memory row
            int num_parked = 0;
            int stage_buffers = 1;   // chunked outputs: staging buffers per warp (1, or one per output array)
            ProgramStats stats;
            Tape tape;
        };

        // Build the symbolic program of one algorithm in the CURRENT graph scope.
        inline Program buildProgram(const ClusterTreeModel &model, int algo)
        {
            ModelCompiler mc(model);
            Program p;
            p.name = algoName(algo);
            const int nq = model.getNumPositions(), nv = model.getNumDegreesOfFreedom();
            switch (algo)
            {
            case ALGO_ID:
                p.n_in[0] = nq, p.n_in[1] = nv, p.n_in[2] = nv;
                p.outputs.push_back(mc.inverseDynamics());
                break;
            case ALGO_FD:
                p.n_in[0] = nq, p.n_in[1] = nv, p.n_in[2] = nv;
                p.outputs.push_back(mc.forwardDynamics());
                break;
            case PROGRAM_FD_LTL:
                p.n_in[0] = nq, p.n_in[1] = nv, p.n_in[2] = nv;
                p.outputs.push_back(mc.forwardDynamicsLTL());
                break;
            case ALGO_GFA:
            case ALGO_GFS:
                p.n_in[0] = nq, p.n_in[1] = 6 * (int)ModelCompiler::externalForceBodies(model).size(), p.n_in[2] = nv;
                p.outputs.push_back(mc.generalizedExternalForce(algo == ALGO_GFA ? 1 : -1));
                break;
            case ALGO_FK:
            {
                p.n_in[0] = nq, p.n_in[1] = nv;
                std::vector<sym::Sym> pp, R, v;
                mc.setFreezeAxisymmetricLeaves(false); // the outputs are the poses themselves
                mc.forwardKinematics(pp, R, v);
                p.outputs = {pp, R, v};
                break;
            }
            case ALGO_H:
                p.n_in[0] = nq;
                p.outputs.push_back(mc.massMatrix());
                break;
            case ALGO_PHI:
            {
                // constraint violation and K_d of every implicit cluster (state generation / validity)
                p.n_in[0] = nq;
                std::vector<sym::Sym> phi_all, Kd_all;
                for (const ClusterTreeNode &c : model.clusters())
                {
                    if (c.joint_.type != ClusterType::Implicit)
                        continue;
                    const ClusterDesc &d = c.joint_;
                    std::vector<sym::Sym> q(d.num_bodies), phi, K;
                    for (int i = 0; i < d.num_bodies; i++)
                        q[i] = sym::Sym::input(IN_Q, c.position_index_ + i);
                    mc.implicitJacobian(d, q, nullptr, phi, K, nullptr);
                    for (auto &x : phi)
                        phi_all.push_back(x);
                    for (int i = 0; i < d.num_constraints; i++)
                        for (int j = 0; j < d.num_bodies; j++)
                            if (!d.independent[j])
                                Kd_all.push_back(K[i * d.num_bodies + j]);
                }
                p.outputs = {phi_all, Kd_all};
                break;
            }
            default:
                throw std::runtime_error("compileAlgo: unknown algorithm");
            }
            return p;
        }

        inline CompiledAlgo compileAlgo(const ClusterTreeModel &model, int algo, bool want_body = true,
                                        int sync_every = 0, ConstTable *consts = nullptr, int out_chunk = 0,
                                        bool park = false)
        {
            sym::Graph graph;
            sym::GraphScope scope(graph);
            const Program p = buildProgram(model, algo);
            CompiledAlgo out;
            out.name = p.name;
            for (int i = 0; i < 3; i++)
                out.n_in[i] = p.n_in[i];
            for (size_t i = 0; i < p.outputs.size(); i++)
                out.n_out[i] = (int)p.outputs[i].size();
            Emitter em(graph, p, consts);
            out.stats = em.stats();
            out.tape = em.tape();
            if (want_body)
            {
                ParkConfig pc;
                for (int i = 0; i < 3; i++)
                    pc.n_slots[i] = p.n_in[i];
                pc.n_slots[3] = out.n_out[0] <= 64 ? out.n_out[0] : 0; // the shells stage output 0 when it is small
                if (const char *gap = std::getenv("GRBDA_PARK_GAP")) // tuning experiments
                    pc.min_gap = std::atoi(gap);
                out.parked = park && sync_every == 0;
                out.body = em.cudaBody(sync_every, out_chunk, out.parked ? &pc : nullptr);
                out.num_parked = em.numParked();
                out.stage_buffers = em.stageBuffers();
                out.range_check = em.cudaRangeCheck();
            }
            return out;
        }

        // The generated `struct Body` of one program: sizes, range check, run<real, FAST>(). Shared by the
        // build-time tool (tools/modelc.cpp) and grbda_cuda_emit_source (host-compiled emitter self test).
            inline void emitBodyStruct(std::ostream &os, const std::string &struct_name, const CompiledAlgo &c)
        {
            os << "struct " << struct_name << "\n{\n";
            os << "    static constexpr int N_IN0 = " << c.n_in[0] << ", N_IN1 = " << c.n_in[1]
               << ", N_IN2 = " << c.n_in[2] << ";\n";
            os << "    static constexpr int N_OUT0 = " << c.n_out[0] << ", N_OUT1 = " << c.n_out[1]
               << ", N_OUT2 = " << c.n_out[2] << ";\n";
            os << "    static constexpr bool RANGE_CHECKED = " << (c.range_check == "true" ? "false" : "true") << ";\n";
            os << "    static constexpr int STAGE_BUFFERS = " << c.stage_buffers << ";\n";
            os << "    static constexpr bool PARKED = " << (c.parked ? "true" : "false") << "; // " << c.num_parked
               << " values parked in the thread's tile row\n";
            os << "    template <typename real>\n    static __device__ __forceinline__ bool inRange(const real "
                  "*__restrict__ in0, const real *__restrict__ in1,\n        const real *__restrict__ in2)\n    {\n"
                  "#define KC(x) ((real)(x))\n#define IN0(i) in0[i]\n#define IN1(i) in1[i]\n#define IN2(i) in2[i]\n"
                  "        return " << c.range_check << ";\n#undef KC\n#undef IN0\n#undef IN1\n#undef IN2\n    }\n";
            os << "    template <typename real, bool FAST>\n";
            if (c.parked)
                // the thread's rows are read AND written (parking): no __restrict__, one pointer per row
                os << "    static __device__ __forceinline__ void run(const real *in0, const real *in1, const real *in2,\n"
                      "        real *out0, real *out1, real *out2, const OutStage<real> &stage)\n    {\n"
                      "        real *const row0 = const_cast<real *>(in0), *const row1 = const_cast<real *>(in1),\n"
                      "                   *const row2 = const_cast<real *>(in2), *const row3 = out0;\n"
                      "#define IN0(i) row0[i]\n#define IN1(i) row1[i]\n#define IN2(i) row2[i]\n#define OUT0(i, x) row3[i] = (x)\n"
                      // volatile: otherwise the compiler forwards the stored value to the load and keeps (spills) it itself
                      "#define PARK_ST(t, i, x) *(volatile real *)(row##t + (i)) = (x)\n"
                      "#define PARK_LD(t, i) (*(volatile real *)(row##t + (i)))\n";
            else
                os << "    static __device__ __forceinline__ void run(const real *__restrict__ in0, const real "
                      "*__restrict__ in1,\n"
                      "        const real *__restrict__ in2, real *__restrict__ out0, real *__restrict__ out1, "
                      "real *__restrict__ out2,\n        const OutStage<real> &stage)\n    {\n"
                      "#define IN0(i) in0[i]\n#define IN1(i) in1[i]\n#define IN2(i) in2[i]\n#define OUT0(i, x) out0[i] = (x)\n"
                      "#define PARK_ST(t, i, x)\n#define PARK_LD(t, i) KC(0.0)\n";
            os << "#define KC(x) ((real)(x))\n#define KT(i) kc_table<real>(i)\n"
                  "#define OUT1(i, x) out1[i] = (x)\n#define OUT2(i, x) out2[i] = (x)\n"
                  "#define GRBDA_ALIGN() __syncwarp()\n#define GRBDA_PIN(x, late) GRBDA_PIN_IMPL(x, late, stage.zero)\n"
                  "#define STG_PUT(j, x) stage.lane[j] = (x)\n"
                  "#define STG_PUTK(k, j, x) stage.lane[(k) * stage.buf_stride + (j)] = (x)\n"
                  "#define STG_FLUSHI0(base, count) flushChunk<real, N_OUT0, count>(stage.g[0], base, stage.warp, stage.valid)\n"
                  "#define STG_FLUSHI1(base, count) flushChunk<real, N_OUT1, count>(stage.g[1], base, stage.warp + stage.buf_stride, stage.valid)\n"
                  "#define STG_FLUSHI2(base, count) flushChunk<real, N_OUT2, count>(stage.g[2], base, stage.warp + 2 * stage.buf_stride, stage.valid)\n"
                  "#define STG_FLUSH0(base, count) flushChunk<real, N_OUT0, count>(stage.g[0], base, stage.warp, stage.valid)\n"
                  "#define STG_FLUSH1(base, count) flushChunk<real, N_OUT1, count>(stage.g[1], base, stage.warp, stage.valid)\n"
                  "#define STG_FLUSH2(base, count) flushChunk<real, N_OUT2, count>(stage.g[2], base, stage.warp, stage.valid)\n";
            os << c.body;
            os << "#undef KC\n#undef KT\n#undef IN0\n#undef IN1\n#undef IN2\n#undef OUT0\n#undef OUT1\n#undef OUT2\n#undef GRBDA_ALIGN\n#undef GRBDA_PIN\n#undef PARK_ST\n#undef PARK_LD\n#undef STG_PUT\n#undef STG_PUTK\n#undef STG_FLUSHI0\n#undef STG_FLUSHI1\n#undef STG_FLUSHI2\n#undef STG_FLUSH0\n#undef STG_FLUSH1\n#undef STG_FLUSH2\n";
            os << "    }\n};\n";
        }


        // ---- limb-parallel (one warp per limb) version of the same program ---------------------------
        struct CompiledRoles
        {
            std::string name;
            int W = 1, num_slots = 0;
            bool has_barrier = false;
            int n_in[3] = {0, 0, 0};
            int n_out[3] = {0, 0, 0};
            std::vector<std::string> bodies;  // per role
            std::vector<ProgramStats> stats;  // per role
            // for the CPU-side self test (tests/tape.py): the whole graph + the role instruction lists
            sym::Graph graph;
            RolePrograms programs;
            std::vector<std::vector<RolePartitioner::Out>> role_outputs;
        };

        inline CompiledRoles compileAlgoRoles(const ClusterTreeModel &model, int algo, bool want_body = true,
                                              ConstTable *consts = nullptr)
        {
            CompiledRoles out;
            const RolePlan plan = planRoles(model);
            sym::GraphScope scope(out.graph);
            const Program p = buildProgram(model, algo);
            out.name = p.name;
            out.W = algo == ALGO_PHI ? 1 : plan.W;
            for (int i = 0; i < 3; i++)
                out.n_in[i] = p.n_in[i];
            for (size_t i = 0; i < p.outputs.size(); i++)
                out.n_out[i] = (int)p.outputs[i].size();
            const int nv = model.getNumDegreesOfFreedom();

            // cluster of a position / velocity / body index
            std::vector<int> pos_cluster(model.getNumPositions()), vel_cluster(nv), body_cluster(model.getNumBodies());
            std::vector<int> weight(std::max(1, plan.W), 0);
            for (const ClusterTreeNode &c : model.clusters())
            {
                for (int i = 0; i < c.num_positions_; i++)
                    pos_cluster[c.position_index_ + i] = c.index_;
                for (int i = 0; i < c.num_velocities_; i++)
                    vel_cluster[c.velocity_index_ + i] = c.index_;
                for (int i = 0; i < c.joint_.num_bodies; i++)
                    body_cluster[c.first_body_ + i] = c.index_;
                if (plan.cluster_role[c.index_] >= 0)
                    weight[plan.cluster_role[c.index_]] += c.joint_.num_bodies * (c.joint_.type == ClusterType::Implicit ? 3 : 1);
            }
            // trunk outputs go to the lightest limb
            const int trunk_role = (int)(std::min_element(weight.begin(), weight.end()) - weight.begin());
            auto roleOfCluster = [&](int c) { return out.W > 1 ? plan.cluster_role[c] : 0; };
            auto input_role = [&](int array, int element) {
                return roleOfCluster(array == IN_Q ? pos_cluster[element] : vel_cluster[element]);
            };
            auto output_role = [&](int array, int element) {
                int r = -1;
                switch (algoOfProgram(algo))
                {
                case ALGO_ID:
                case ALGO_FD: r = roleOfCluster(vel_cluster[element]); break;
                case ALGO_FK:
                    r = roleOfCluster(body_cluster[element / (array == 0 ? 3 : (array == 1 ? 9 : 6))]);
                    break;
                case ALGO_H:
                {
                    const int ra = roleOfCluster(vel_cluster[element / nv]), rb = roleOfCluster(vel_cluster[element % nv]);
                    r = ra >= 0 ? ra : rb;
                    break;
                }
                default: r = 0;
                }
                return r >= 0 ? r : trunk_role;
            };
            RolePartitioner part(out.graph, p, out.W, input_role, output_role);
            out.programs = part.build();
            out.num_slots = out.programs.num_slots;
            out.has_barrier = out.programs.has_barrier;
            out.role_outputs = part.roleOutputs();
            RoleEmitter em(out.graph, part, out.programs, consts);
            for (int r = 0; r < out.W; r++)
            {
                out.stats.push_back(em.roleStats(r));
                if (want_body)
                    out.bodies.push_back(em.roleBody(r));
            }
            return out;
        }

        // Role tape (tests/tape.py run_role_tape): int32 header {magic 0x47524245, n_nodes, W, num_slots,
        // n_in0, n_in1, n_in2, n_out0, n_out1, n_out2}; node table op,a,b,c,e (int32) + val (float64);
        // per role: int32 n_ops, then (kind, id, slot) int32 triples; int32 n_outputs, then
        // (id, array, element) int32 triples.
        inline void writeRoleTape(const CompiledRoles &c, const std::string &path)
        {
            std::ofstream f(path, std::ios::binary);
            if (!f)
                throw std::runtime_error("cannot write " + path);
            const int32_t n = (int32_t)c.graph.nodes.size();
            const int32_t hdr[10] = {0x47524245, n, c.W, c.num_slots, c.n_in[0], c.n_in[1], c.n_in[2],
                                     c.n_out[0], c.n_out[1], c.n_out[2]};
            f.write((const char *)hdr, sizeof(hdr));
            std::vector<int32_t> col(n);
            auto dump = [&](auto get) {
                for (int32_t i = 0; i < n; i++)
                    col[i] = get(c.graph.nodes[i]);
                f.write((const char *)col.data(), (size_t)n * 4);
            };
            dump([](const sym::Node &x) { return (int32_t)x.op; });
            dump([](const sym::Node &x) { return x.a; });
            dump([](const sym::Node &x) { return x.b; });
            dump([](const sym::Node &x) { return x.c; });
            dump([](const sym::Node &x) { return x.e; });
            for (int32_t i = 0; i < n; i++)
                f.write((const char *)&c.graph.nodes[i].val, 8);
            for (int r = 0; r < c.W; r++)
            {
                const int32_t no = (int32_t)c.programs.ops[r].size();
                f.write((const char *)&no, 4);
                for (const RoleOp &op : c.programs.ops[r])
                {
                    const int32_t t[3] = {op.kind, op.id, op.slot};
                    f.write((const char *)t, 12);
                }
                const int32_t nout = (int32_t)c.role_outputs[r].size();
                f.write((const char *)&nout, 4);
                for (auto &o : c.role_outputs[r])
                {
                    const int32_t t[3] = {o.id, o.array, o.element};
                    f.write((const char *)t, 12);
                }
            }
        }

        // Binary tape: int32 header {magic, n_ops, n_arrays, n_in0, n_in1, n_in2}, then op,a,b,c,e
        // (int32 each, n_ops), val (float64 n_ops), then per output array: int32 count + indices.
        inline void writeTape(const CompiledAlgo &c, const std::string &path)
        {
            std::ofstream f(path, std::ios::binary);
            if (!f)
                throw std::runtime_error("cannot write " + path);
            const Tape &t = c.tape;
            const int32_t hdr[6] = {0x47524244, (int32_t)t.op.size(), (int32_t)t.outputs.size(), c.n_in[0],
                                    c.n_in[1], c.n_in[2]};
            f.write((const char *)hdr, sizeof(hdr));
            auto wi = [&](const std::vector<int32_t> &v) { f.write((const char *)v.data(), v.size() * 4); };
            wi(t.op), wi(t.a), wi(t.b), wi(t.c), wi(t.e);
            f.write((const char *)t.val.data(), t.val.size() * 8);
            for (auto &o : t.outputs)
            {
                const int32_t n = (int32_t)o.size();
                f.write((const char *)&n, 4);
                wi(o);
            }
        }

    } // namespace compiler
} // namespace grbda
