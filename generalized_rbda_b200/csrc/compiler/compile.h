// Device-side model compiler, part 3: driver. Model + algorithm -> emitted CUDA body, tape, counts.
#pragma once
#include <fstream>
#include "algorithms.h"
#include "emit.h"

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
            ALGO_COUNT = 5
        };
        inline const char *algoName(int a)
        {
            static const char *names[] = {"id", "fd", "fk", "h", "phi"};
            return names[a];
        }

        struct CompiledAlgo
        {
            std::string name;
            int n_in[3] = {0, 0, 0};
            int n_out[3] = {0, 0, 0};
            std::string body;
            ProgramStats stats;
            Tape tape;
        };

        inline CompiledAlgo compileAlgo(const ClusterTreeModel &model, int algo, bool want_body = true)
        {
            sym::Graph graph;
            sym::GraphScope scope(graph);
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
            case ALGO_FK:
            {
                p.n_in[0] = nq, p.n_in[1] = nv;
                std::vector<sym::Sym> pp, R, v;
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
            CompiledAlgo out;
            out.name = p.name;
            for (int i = 0; i < 3; i++)
                out.n_in[i] = p.n_in[i];
            for (size_t i = 0; i < p.outputs.size(); i++)
                out.n_out[i] = (int)p.outputs[i].size();
            Emitter em(graph, p);
            out.stats = em.stats();
            out.tape = em.tape();
            if (want_body)
                out.body = em.cudaBody();
            return out;
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
