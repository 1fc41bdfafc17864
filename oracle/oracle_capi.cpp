// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points (ctypes) around the CPU restatement in
// grbda_oracle/*.h. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (generalized_rbda_b200) never does.
//
// A model is either one of the reference's hand-coded robots (grbda_oracle/robots.h) or is
// assembled body by body / cluster by cluster through oracle_builder_* (the same calls the
// reference's registerBody / appendRegisteredBodiesAsCluster make), which is how the tests hand
// URDF-derived models to the oracle.
#include <atomic>
#include <cstring>
#include <ctime>
#include <sstream>
#include <thread>
#include "grbda_oracle/robots.h"
#include "grbda_oracle/rng.h"
#include "grbda_oracle/dual.h"

using namespace grbda_oracle;

namespace
{
    // Recorded builder commands so that a model can be instantiated for any scalar type and
    // once per worker thread (the model caches per-state data in its nodes, like the reference).
    struct BodyCmd
    {
        std::string name, parent;
        double inertia[36], E[9], r[3];
    };
    struct LoopCapture // ClusterTreeParsing.cpp:131-177 (LoopConstraintCapture)
    {
        std::vector<int> nca_to_predecessor, nca_to_successor; // sub-indices within the cluster
        double pred_E[9], pred_r[3], succ_E[9], succ_r[3];
    };
    struct ClusterCmd
    {
        std::string name;
        int kind; // 0 Free(quat) 1 Free(rpy) 2 Revolute 3 RevoluteWithRotor 4 Generic+Static
                  // 5 Generic+loops 6 RevolutePairWithRotor 7 RevolutePair
        std::vector<BodyCmd> bodies;
        std::vector<int> axes;
        std::vector<int> independent;
        std::vector<double> K, G; // Static: row-major K (nc x N), G (N x n)
        int num_constraints = 0;
        std::vector<LoopCapture> loops;
        std::vector<int> loop_rows; // rows (0..2) of each loop that are kept, flattened (loop, axis)
        double gear[3] = {0, 0, 0};
        std::vector<double> belt1, belt2, belt3;
        // kind 8: phi given as a straight-line op list (op, a, b, val), see include/grbda_cuda.h
        std::vector<int> phi_op, phi_a, phi_b, phi_out;
        std::vector<double> phi_val;
    };
    struct ModelSpec
    {
        std::string named;
        bool generic = false;
        double gravity[3] = {0, 0, -9.81};
        std::vector<ClusterCmd> clusters;
        ClusterCmd pending;

        template <typename T>
        static Mat<T> m3(const double *p)
        {
            Mat<T> m(3, 3);
            for (int i = 0; i < 9; i++)
                m.a[i] = T(p[i]);
            return m;
        }
        template <typename T>
        static Mat<T> v3(const double *p)
        {
            return vec3<T>(T(p[0]), T(p[1]), T(p[2]));
        }

        template <typename T>
        ClusterTreeModel<T> instantiate() const
        {
            ClusterTreeModel<T> base = instantiateBase<T>();
            if (generic)
                return extractGenericJointModel(base);
            return base;
        }

        template <typename T>
        ClusterTreeModel<T> instantiateBase() const
        {
            if (!named.empty())
            {
                if (named == "tello")
                    return buildTello<T>();
                if (named == "tello_with_arms")
                    return buildTelloWithArms<T>();
                if (named == "mini_cheetah")
                    return buildMiniCheetah<T>(true);
                if (named == "mini_cheetah_rpy")
                    return buildMiniCheetah<T>(false);
                if (named == "mit_humanoid")
                    return buildMitHumanoid<T>(true);
                if (named == "mit_humanoid_rpy")
                    return buildMitHumanoid<T>(false);
                if (named == "mit_humanoid_leg")
                    return buildMitHumanoidLeg<T>();
                const std::string a = "revolute_chain_with_rotor_", b = "revolute_pair_chain_with_rotor_";
                if (named.compare(0, a.size(), a) == 0)
                    return buildRevoluteChainWithRotor<T>(std::stoi(named.substr(a.size())));
                if (named.compare(0, b.size(), b) == 0)
                    return buildRevolutePairChainWithRotor<T>(std::stoi(named.substr(b.size())));
                throw std::runtime_error("oracle: unknown robot '" + named + "'");
            }
            ClusterTreeModel<T> model;
            model.setGravity(gravity[0], gravity[1], gravity[2]);
            for (const ClusterCmd &c : clusters)
            {
                std::vector<Body<T>> bodies;
                for (const BodyCmd &b : c.bodies)
                {
                    Mat<T> I(6, 6);
                    for (int i = 0; i < 36; i++)
                        I.a[i] = T(b.inertia[i]);
                    bodies.push_back(model.registerBody(b.name, I, b.parent,
                                                        Transform<T>(m3<T>(b.E), v3<T>(b.r))));
                }
                std::shared_ptr<ClusterJointBase<T>> joint;
                const int N = (int)bodies.size();
                switch (c.kind)
                {
                case 0:
                case 1:
                    joint = std::make_shared<FreeCluster<T>>(bodies[0], c.kind == 0);
                    break;
                case 2:
                    joint = std::make_shared<RevoluteCluster<T>>(bodies[0], (Axis)c.axes[0]);
                    break;
                case 3:
                {
                    GearedTransmissionModule<T> m{bodies[0], bodies[1], (Axis)c.axes[0],
                                                  (Axis)c.axes[1], T(c.gear[0])};
                    joint = std::make_shared<RevoluteWithRotorCluster<T>>(m);
                    break;
                }
                case 4:
                {
                    JointVec<T> joints;
                    for (int i = 0; i < N; i++)
                        joints.push_back(std::make_shared<SingleRevolute<T>>((Axis)c.axes[i]));
                    const int n = (int)c.G.size() / N;
                    Mat<T> G(N, n), K(c.num_constraints, N);
                    for (size_t i = 0; i < c.G.size(); i++)
                        G.a[i] = T(c.G[i]);
                    for (size_t i = 0; i < c.K.size(); i++)
                        K.a[i] = T(c.K[i]);
                    joint = std::make_shared<GenericCluster<T>>(
                        bodies, joints, std::make_shared<StaticConstraint<T>>(G, K));
                    break;
                }
                case 5:
                {
                    JointVec<T> joints;
                    for (int i = 0; i < N; i++)
                        joints.push_back(std::make_shared<SingleRevolute<T>>((Axis)c.axes[i]));
                    std::vector<bool> ind;
                    for (int v : c.independent)
                        ind.push_back(v != 0);
                    // ClusterTreeParsing.cpp:310-376 (implicitPositionConstraint)
                    using S = Taylor2<T>;
                    std::vector<Transform<S>> xtree;
                    std::vector<Axis> axes;
                    for (int i = 0; i < N; i++)
                    {
                        xtree.push_back(Transform<S>(m3<S>(c.bodies[i].E), v3<S>(c.bodies[i].r)));
                        axes.push_back((Axis)c.axes[i]);
                    }
                    std::vector<LoopCapture> loops = c.loops;
                    std::vector<int> rows = c.loop_rows;
                    auto phi = [xtree, axes, loops, rows](const std::vector<S> &q)
                    {
                        std::vector<S> out;
                        for (size_t l = 0; l < loops.size(); l++)
                        {
                            auto through = [&](const std::vector<int> &chain, const double *E,
                                               const double *r)
                            {
                                Transform<S> X;
                                for (int sub : chain)
                                {
                                    Transform<S> XJ(coordinateRotation<S>(axes[sub], q[sub]));
                                    X = XJ * xtree[sub] * X;
                                }
                                X = Transform<S>(m3<S>(E), v3<S>(r)) * X;
                                return X.r;
                            };
                            Mat<S> rp = through(loops[l].nca_to_predecessor, loops[l].pred_E,
                                                loops[l].pred_r);
                            Mat<S> rs = through(loops[l].nca_to_successor, loops[l].succ_E,
                                                loops[l].succ_r);
                            for (int j = 0; j < 3; j++)
                                if (rows[3 * l + j])
                                    out.push_back(rp[j] - rs[j]);
                        }
                        return out;
                    };
                    joint = std::make_shared<GenericCluster<T>>(
                        bodies, joints, std::make_shared<GenericImplicitConstraint<T>>(ind, phi));
                    break;
                }
                case 10:
                {
                    // ClusterJoints::FourBar (FourBarJoint.h): three revolute joints + LoopConstraint::FourBar;
                    // belt1 = path-1 link lengths, belt2 = path-2 link lengths, belt3 = offset, gear[0] = independent coordinate
                    JointVec<T> joints;
                    for (int i = 0; i < N; i++)
                        joints.push_back(std::make_shared<SingleRevolute<T>>((Axis)c.axes[i]));
                    std::vector<T> p1, p2;
                    for (double x : c.belt1) p1.push_back(T(x));
                    for (double x : c.belt2) p2.push_back(T(x));
                    joint = std::make_shared<GenericCluster<T>>(
                        bodies, joints,
                        std::make_shared<FourBarConstraint<T>>(p1, p2, T(c.belt3[0]), T(c.belt3[1]), (int)c.gear[0]));
                    break;
                }
                case 8:
                {
                    JointVec<T> joints;
                    for (int i = 0; i < N; i++)
                        joints.push_back(std::make_shared<SingleRevolute<T>>((Axis)c.axes[i]));
                    std::vector<bool> ind;
                    for (int v : c.independent)
                        ind.push_back(v != 0);
                    using S = Taylor2<T>;
                    const ClusterCmd cc = c;
                    auto phi = [cc](const std::vector<S> &q)
                    {
                        std::vector<S> v;
                        for (size_t i = 0; i < cc.phi_op.size(); i++)
                        {
                            const int a = cc.phi_a[i], b = cc.phi_b[i];
                            switch (cc.phi_op[i])
                            {
                            case 0: v.push_back(S(cc.phi_val[i])); break;
                            case 1: v.push_back(q[b]); break;
                            case 2: v.push_back(v[a] + v[b]); break;
                            case 3: v.push_back(v[a] - v[b]); break;
                            case 4: v.push_back(v[a] * v[b]); break;
                            case 5:
                                if (cc.phi_op[b] != 0)
                                    throw std::runtime_error("oracle: phi division by a non-constant");
                                v.push_back(v[a] / cc.phi_val[b]);
                                break;
                            case 6: v.push_back(-v[a]); break;
                            case 7: v.push_back(sin(v[a])); break;
                            case 8: v.push_back(cos(v[a])); break;
                            default: throw std::runtime_error("oracle: unsupported phi op");
                            }
                        }
                        std::vector<S> out;
                        for (int o : cc.phi_out)
                            out.push_back(v[o]);
                        return out;
                    };
                    joint = std::make_shared<GenericCluster<T>>(
                        bodies, joints, std::make_shared<GenericImplicitConstraint<T>>(ind, phi));
                    break;
                }
                case 6:
                {
                    // bodies given in registration order; axes/roles: c.independent holds the
                    // sub-indices {link1, rotor1, rotor2, link2}
                    auto conv = [](const std::vector<double> &v)
                    {
                        std::vector<T> o;
                        for (double x : v)
                            o.push_back(T(x));
                        return o;
                    };
                    const int l1 = c.independent[0], r1 = c.independent[1], r2 = c.independent[2],
                              l2 = c.independent[3];
                    ParallelBeltTransmissionModule<T> m1{bodies[l1], bodies[r1], (Axis)c.axes[l1],
                                                         (Axis)c.axes[r1], T(c.gear[0]), conv(c.belt1)};
                    ParallelBeltTransmissionModule<T> m2{bodies[l2], bodies[r2], (Axis)c.axes[l2],
                                                         (Axis)c.axes[r2], T(c.gear[1]), conv(c.belt2)};
                    joint = std::make_shared<RevolutePairWithRotorCluster<T>>(m1, m2);
                    break;
                }
                case 7:
                    joint = std::make_shared<RevolutePairCluster<T>>(bodies[0], bodies[1],
                                                                     (Axis)c.axes[0], (Axis)c.axes[1]);
                    break;
                case 9:
                {
                    // RevoluteTripleWithRotor: bodies in the reference's fixed order link1..3, rotor1..3
                    auto conv = [](const std::vector<double> &v)
                    {
                        std::vector<T> o;
                        for (double x : v)
                            o.push_back(T(x));
                        return o;
                    };
                    ParallelBeltTransmissionModule<T> m1{bodies[0], bodies[3], (Axis)c.axes[0], (Axis)c.axes[3],
                                                         T(c.gear[0]), conv(c.belt1)};
                    ParallelBeltTransmissionModule<T> m2{bodies[1], bodies[4], (Axis)c.axes[1], (Axis)c.axes[4],
                                                         T(c.gear[1]), conv(c.belt2)};
                    ParallelBeltTransmissionModule<T> m3{bodies[2], bodies[5], (Axis)c.axes[2], (Axis)c.axes[5],
                                                         T(c.gear[2]), conv(c.belt3)};
                    joint = std::make_shared<RevoluteTripleWithRotorCluster<T>>(m1, m2, m3);
                    break;
                }
                default:
                    throw std::runtime_error("oracle: unknown cluster kind");
                }
                model.appendRegisteredBodiesAsCluster(c.name, joint);
            }
            return model;
        }
    };

    inline uint64_t nextUid()
    {
        static std::atomic<uint64_t> n{1};
        return n++;
    }
    struct Handle
    {
        uint64_t uid = nextUid(); // identity for per-thread caches (a freed handle's address may be reused)
        ModelSpec spec;
        std::unique_ptr<ClusterTreeModel<double>> model; // instance 0 (sizes, introspection)
        std::vector<std::unique_ptr<ClusterTreeModel<double>>> workers;
    };

    thread_local std::string g_error;

    template <typename F>
    int guarded(F f)
    {
        try
        {
            f();
            return 0;
        }
        catch (const std::exception &e)
        {
            g_error = e.what();
            return 1;
        }
    }

    ClusterTreeModel<double> &worker(Handle *h, int tid)
    {
        return *h->workers[tid];
    }

    void ensureWorkers(Handle *h, int n)
    {
        while ((int)h->workers.size() < n)
        {
            h->workers.push_back(
                std::make_unique<ClusterTreeModel<double>>(h->spec.instantiate<double>()));
            h->workers.back()->contact_points = h->model->contact_points;
        }
    }

    int resolveThreads(int threads)
    {
        if (threads <= 0)
            threads = (int)std::thread::hardware_concurrency();
        return threads > 0 ? threads : 1;
    }

    Mat<double> toVec(const double *p, int n)
    {
        Mat<double> v(n, 1);
        for (int i = 0; i < n; i++)
            v[i] = p[i];
        return v;
    }

    // One state per iteration, contiguous static chunks per worker thread (std::thread; the
    // reference itself is single threaded, one model instance per thread is required because the
    // model caches per-state data in its nodes).
    template <typename Body>
    int runBatch(Handle *h, int64_t batch, int threads, Body body)
    {
        threads = resolveThreads(threads);
        if (batch < threads)
            threads = batch > 0 ? (int)batch : 1;
        ensureWorkers(h, threads);
        std::vector<std::string> errors(threads);
        auto work = [&](int tid)
        {
            const int64_t lo = batch * tid / threads, hi = batch * (tid + 1) / threads;
            try
            {
                for (int64_t b = lo; b < hi; b++)
                    body(worker(h, tid), b);
            }
            catch (const std::exception &e)
            {
                errors[tid] = e.what();
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++)
            pool.emplace_back(work, t);
        work(0);
        for (auto &t : pool)
            t.join();
        for (auto &e : errors)
            if (!e.empty())
            {
                g_error = e;
                return 1;
            }
        return 0;
    }
} // namespace

extern "C"
{
    const char *oracle_last_error() { return g_error.c_str(); }

    // name: one of the robots in robots.h; append ":generic" for the Generic re-build
    // (UnitTests/testHelpers.hpp:10-45)
    void *oracle_model_create(const char *name)
    {
        Handle *h = new Handle();
        std::string n = name;
        const std::string suffix = ":generic";
        if (n.size() > suffix.size() && n.compare(n.size() - suffix.size(), suffix.size(), suffix) == 0)
        {
            h->spec.generic = true;
            n = n.substr(0, n.size() - suffix.size());
        }
        h->spec.named = n;
        if (guarded([&]
                    { h->model = std::make_unique<ClusterTreeModel<double>>(h->spec.instantiate<double>()); }))
        {
            delete h;
            return nullptr;
        }
        return h;
    }

    void *oracle_builder_create(const double *gravity3)
    {
        Handle *h = new Handle();
        for (int i = 0; i < 3; i++)
            h->spec.gravity[i] = gravity3[i];
        return h;
    }
    // inertia: row-major 6x6; E row-major 3x3; r 3
    void oracle_builder_register_body(void *hv, const char *name, const char *parent,
                                      const double *inertia36, const double *E9, const double *r3)
    {
        Handle *h = (Handle *)hv;
        BodyCmd b;
        b.name = name;
        b.parent = parent;
        std::memcpy(b.inertia, inertia36, sizeof(b.inertia));
        std::memcpy(b.E, E9, sizeof(b.E));
        std::memcpy(b.r, r3, sizeof(b.r));
        h->spec.pending.bodies.push_back(b);
    }
    void oracle_builder_append_simple(void *hv, const char *name, int kind, const int *axes,
                                      double gear_ratio)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        c.name = name;
        c.kind = kind;
        for (size_t i = 0; i < c.bodies.size(); i++)
            c.axes.push_back(axes ? axes[i] : 0);
        c.gear[0] = gear_ratio;
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    void oracle_builder_append_generic_static(void *hv, const char *name, const int *axes,
                                              const double *K, int num_constraints,
                                              const double *G, int num_independent)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        const int N = (int)c.bodies.size();
        c.name = name;
        c.kind = 4;
        c.axes.assign(axes, axes + N);
        c.num_constraints = num_constraints;
        c.K.assign(K, K + (size_t)num_constraints * N);
        c.G.assign(G, G + (size_t)N * num_independent);
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    // chains: sub-indices from the NCA (exclusive) to predecessor / successor; origins: E9 + r3 of
    // the constraint frame in the predecessor / successor link; keep_rows[3]: which of x,y,z rows
    // of (r_pred - r_succ) depend on q (ClusterTreeParsing.cpp:362, casadi which_depends)
    void oracle_builder_add_loop(void *hv, const int *pred_chain, int n_pred, const int *succ_chain,
                                 int n_succ, const double *pred_E9, const double *pred_r3,
                                 const double *succ_E9, const double *succ_r3, const int *keep_rows)
    {
        Handle *h = (Handle *)hv;
        LoopCapture l;
        l.nca_to_predecessor.assign(pred_chain, pred_chain + n_pred);
        l.nca_to_successor.assign(succ_chain, succ_chain + n_succ);
        std::memcpy(l.pred_E, pred_E9, sizeof(l.pred_E));
        std::memcpy(l.pred_r, pred_r3, sizeof(l.pred_r));
        std::memcpy(l.succ_E, succ_E9, sizeof(l.succ_E));
        std::memcpy(l.succ_r, succ_r3, sizeof(l.succ_r));
        h->spec.pending.loops.push_back(l);
        for (int j = 0; j < 3; j++)
            h->spec.pending.loop_rows.push_back(keep_rows[j]);
    }
    void oracle_builder_append_generic_loops(void *hv, const char *name, const int *axes,
                                             const int *independent)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        const int N = (int)c.bodies.size();
        c.name = name;
        c.kind = 5;
        c.axes.assign(axes, axes + N);
        c.independent.assign(independent, independent + N);
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    // ClusterJoints::FourBar: the reference's manual builders (src/Robots/PlanarLegLinkage.cpp:66-83)
    void oracle_builder_append_four_bar(void *hv, const char *name, const int *axes, const double *path1, int n1,
                                        const double *path2, int n2, const double *offset, int independent_coordinate)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        const int N = (int)c.bodies.size();
        c.name = name;
        c.kind = 10;
        c.axes.assign(axes, axes + N);
        c.belt1.assign(path1, path1 + n1);
        c.belt2.assign(path2, path2 + n2);
        c.belt3.assign(offset, offset + 2);
        c.gear[0] = independent_coordinate;
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    // phi as an op list: op/a/b (int, n_ops), val (double, n_ops), outputs (int, n_out)
    void oracle_builder_append_generic_phi(void *hv, const char *name, const int *axes, const int *independent,
                                           const int *op, const int *a, const int *b, const double *val,
                                           int n_ops, const int *outputs, int n_out)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        const int N = (int)c.bodies.size();
        c.name = name;
        c.kind = 8;
        c.axes.assign(axes, axes + N);
        c.independent.assign(independent, independent + N);
        c.phi_op.assign(op, op + n_ops);
        c.phi_a.assign(a, a + n_ops);
        c.phi_b.assign(b, b + n_ops);
        c.phi_val.assign(val, val + n_ops);
        c.phi_out.assign(outputs, outputs + n_out);
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    // roles[4] = sub-indices {link1, rotor1, rotor2, link2}
    void oracle_builder_append_revolute_pair_with_rotor(void *hv, const char *name, const int *axes,
                                                        const int *roles, double gear1, double gear2,
                                                        const double *belt1, int nb1,
                                                        const double *belt2, int nb2)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        c.name = name;
        c.kind = 6;
        c.axes.assign(axes, axes + 4);
        c.independent.assign(roles, roles + 4);
        c.gear[0] = gear1;
        c.gear[1] = gear2;
        c.belt1.assign(belt1, belt1 + nb1);
        c.belt2.assign(belt2, belt2 + nb2);
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    // bodies registered in the order link1, link2, link3, rotor1, rotor2, rotor3; belts: 1 + 2 + 3 ratios
    void oracle_builder_append_revolute_triple_with_rotor(void *hv, const char *name, const int *axes,
                                                          const double *gears3, const double *belts6)
    {
        Handle *h = (Handle *)hv;
        ClusterCmd &c = h->spec.pending;
        c.name = name;
        c.kind = 9;
        c.axes.assign(axes, axes + 6);
        for (int i = 0; i < 3; i++)
            c.gear[i] = gears3[i];
        c.belt1.assign(belts6, belts6 + 1);
        c.belt2.assign(belts6 + 1, belts6 + 3);
        c.belt3.assign(belts6 + 3, belts6 + 6);
        h->spec.clusters.push_back(c);
        h->spec.pending = ClusterCmd();
    }
    int oracle_builder_finish(void *hv, int generic)
    {
        Handle *h = (Handle *)hv;
        h->spec.generic = generic != 0;
        return guarded([&]
                       { h->model = std::make_unique<ClusterTreeModel<double>>(h->spec.instantiate<double>()); });
    }

    void oracle_model_destroy(void *h) { delete (Handle *)h; }
    int oracle_num_positions(void *h) { return ((Handle *)h)->model->getNumPositions(); }
    int oracle_num_dof(void *h) { return ((Handle *)h)->model->getNumDegreesOfFreedom(); }
    int oracle_num_bodies(void *h) { return ((Handle *)h)->model->getNumBodies(); }
    int oracle_num_clusters(void *h) { return (int)((Handle *)h)->model->nodes.size(); }

    // Per cluster: {parent, num_bodies, num_positions, num_velocities, position_index,
    // velocity_index, is_implicit}; type name copied into names (64 bytes per cluster).
    void oracle_cluster_info(void *hv, int *info7, char *names)
    {
        Handle *h = (Handle *)hv;
        int i = 0;
        for (auto &n : h->model->nodes)
        {
            int *o = info7 + 7 * i;
            o[0] = n->parent_index;
            o[1] = (int)n->bodies.size();
            o[2] = n->num_positions;
            o[3] = n->num_velocities;
            o[4] = n->position_index;
            o[5] = n->velocity_index;
            o[6] = n->joint->loop_constraint->isImplicit();
            if (names)
                std::snprintf(names + 64 * i, 64, "%s", n->joint->typeName());
            i++;
        }
    }
    // Per body: name (64 bytes), parent, cluster, sub index, Xtree E(9) r(3), inertia (36)
    void oracle_body_info(void *hv, int body, char *name64, int *ints3, double *E9, double *r3,
                          double *I36)
    {
        Handle *h = (Handle *)hv;
        const Body<double> &b = h->model->bodies[body];
        std::snprintf(name64, 64, "%s", b.name.c_str());
        ints3[0] = b.parent_index;
        ints3[1] = h->model->clusterContainingBody(body);
        ints3[2] = b.sub_index_within_cluster;
        for (int i = 0; i < 9; i++)
            E9[i] = b.Xtree.E.a[i];
        for (int i = 0; i < 3; i++)
            r3[i] = b.Xtree.r[i];
        for (int i = 0; i < 36; i++)
            I36[i] = b.inertia.a[i];
    }

    // Layout for every batched call: one state per column, contiguous per state
    // (q + b*nq, yd + b*nv, ...). threads <= 0: all OpenMP threads.
    int oracle_inverse_dynamics(void *hv, const double *q, const double *yd, const double *ydd,
                                double *tau, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), toVec(yd + b * nv, nv));
                            Mat<double> t = m.inverseDynamics(toVec(ydd + b * nv, nv));
                            for (int i = 0; i < nv; i++)
                                tau[b * nv + i] = t[i]; });
    }
    int oracle_forward_dynamics(void *hv, const double *q, const double *yd, const double *tau,
                                double *ydd, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), toVec(yd + b * nv, nv));
                            Mat<double> a = m.forwardDynamics(toVec(tau + b * nv, nv));
                            for (int i = 0; i < nv; i++)
                                ydd[b * nv + i] = a[i]; });
    }
    // f_ext: optional [6*nb per state] world-frame spatial forces on every body (may be null)
    int oracle_dynamics_with_external_forces(void *hv, const double *q, const double *yd,
                                             const double *in3, const double *f_ext, double *out,
                                             int forward, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        const int nb = h->model->getNumBodies();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), toVec(yd + b * nv, nv));
                            if (f_ext)
                                for (int i = 0; i < nb; i++)
                                    m.applyExternalForce(i, toVec(f_ext + (b * nb + i) * 6, 6));
                            Mat<double> r = forward ? m.forwardDynamics(toVec(in3 + b * nv, nv))
                                                    : m.inverseDynamics(toVec(in3 + b * nv, nv));
                            for (int i = 0; i < nv; i++)
                                out[b * nv + i] = r[i]; });
    }
    // H: nv x nv per state (symmetric, so row/column-major agree)
    int oracle_mass_matrix(void *hv, const double *q, double *H, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), Mat<double>(nv, 1));
                            Mat<double> Hm = m.getMassMatrix();
                            for (int i = 0; i < nv * nv; i++)
                                H[b * nv * nv + i] = Hm.a[i]; });
    }
    // per state and body: p[3], R[9] row-major (body-to-world), [w_world(3); v_world(3)]
    int oracle_forward_kinematics(void *hv, const double *q, const double *yd, double *p, double *R,
                                  double *v, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        const int nb = h->model->getNumBodies();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), toVec(yd + b * nv, nv));
                            m.forwardKinematics();
                            for (int i = 0; i < nb; i++)
                                m.bodyKinematics(i, p + (b * nb + i) * 3, R + (b * nb + i) * 9,
                                                 v + (b * nb + i) * 6); });
    }
    // valid[b] = 1 when every implicit cluster satisfies |phi(q)| < 1e-8 (ClusterJoint.cpp:41-48)
    int oracle_validate_states(void *hv, const double *q, int *valid, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), Mat<double>(nv, 1));
                            m.forwardKinematics();
                            valid[b] = m.all_positions_valid; });
    }
    // Loop-constraint quantities of one cluster at one state: G (N_span x n), K (nc x N_span),
    // g (N_span), k (nc), row-major. Sizes returned in dims[4] = {N_span, n, nc, 0}.
    int oracle_cluster_constraint(void *hv, int cluster, const double *q, const double *yd,
                                  double *G, double *K, double *g, double *k, int *dims)
    {
        Handle *h = (Handle *)hv;
        return guarded([&]
                       {
            auto &m = *h->model;
            const int nq = m.getNumPositions(), nv = m.getNumDegreesOfFreedom();
            m.setState(toVec(q, nq), toVec(yd, nv));
            m.forwardKinematics();
            auto &j = m.nodes[cluster]->joint;
            dims[0] = j->G().r; dims[1] = j->G().c; dims[2] = j->K().r; dims[3] = 0;
            for (int i = 0; i < j->G().size(); i++) G[i] = j->G().a[i];
            for (int i = 0; i < j->K().size(); i++) K[i] = j->K().a[i];
            for (int i = 0; i < j->g().size(); i++) g[i] = j->g().a[i];
            for (int i = 0; i < j->k().size(); i++) k[i] = j->k().a[i]; });
    }

    int oracle_generate_states(void *hv, uint64_t seed, int64_t first_index, int64_t count,
                               double *q, double *yd, double *aux, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, count, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        { generateState(m, seed, (uint64_t)(first_index + b), q + b * nq, yd + b * nv,
                                        aux + b * nv); });
    }

    // Operation counts of one evaluation with the counting scalar (F_alg definition, scalar.h).
    // algo: 0 ID, 1 FD, 2 FK, 3 H. out[10] = {add,mul,div,sqrt,trig}_all, {..}_alg
    // ---- operational space -----------------------------------------------------------------------------------
    int oracle_set_contact_points(void *hv, int count, const int *bodies, const double *offsets, const int *is_ee)
    {
        Handle *h = (Handle *)hv;
        return guarded([&]
                       {
            auto apply = [&](ClusterTreeModel<double> &m) {
                m.contact_points.clear();
                for (int i = 0; i < count; i++)
                    m.appendContactPoint(bodies[i], toVec(offsets + 3 * i, 3), is_ee && is_ee[i]);
            };
            apply(*h->model);
            for (auto &w : h->workers)
                apply(*w);
        });
    }
    // p, v: [batch][n_cp][3]
    int oracle_contact_kinematics(void *hv, const double *q, const double *yd, double *p, double *v, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        const int nc = (int)h->model->contact_points.size();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), toVec(yd + b * nv, nv));
                            m.forwardKinematics();
                            for (int c = 0; c < nc; c++)
                                m.contactPointKinematics(c, p + (b * nc + c) * 3, v + (b * nc + c) * 3); });
    }
    // J: [batch][n_cp][6][nv]; world != 0: contactJacobianWorldFrame, else contactJacobianBodyFrame
    int oracle_contact_jacobians(void *hv, const double *q, double *J, int world, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        const int nc = (int)h->model->contact_points.size();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), Mat<double>(nv, 1));
                            m.forwardKinematics();
                            for (int c = 0; c < nc; c++)
                            {
                                const Mat<double> Jc = m.contactJacobian(c, world != 0);
                                for (int i = 0; i < 6 * nv; i++)
                                    J[((b * nc + c) * 6) * nv + i] = Jc.a[i];
                            } });
    }
    // force: [batch][n_cp][3] -> dstate [batch][n_cp][nv], lambda_inv [batch][n_cp]
    int oracle_apply_test_force(void *hv, const double *q, const double *force, double *dstate, double *lambda_inv,
                                int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        const int nc = (int)h->model->contact_points.size();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), Mat<double>(nv, 1));
                            for (int c = 0; c < nc; c++)
                                m.applyTestForceReference(c, force + (b * nc + c) * 3, dstate + (b * nc + c) * nv,
                                                          lambda_inv + b * nc + c); });
    }
    // Lambda^-1: [batch][6 n_ee][6 n_ee]
    int oracle_inverse_osim(void *hv, const double *q, double *L, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        const int n = 6 * h->model->getNumEndEffectors();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            m.setState(toVec(q + b * nq, nq), Mat<double>(nv, 1));
                            const Mat<double> Li = m.inverseOperationalSpaceInertiaMatrixReference();
                            for (int i = 0; i < n * n; i++)
                                L[b * n * n + i] = Li.a[i]; });
    }
    // semi-implicit Euler step (rng.h integrateState); flags[b] = 1 where an implicit cluster could not be projected
    int oracle_integrate(void *hv, const double *q, const double *yd, const double *ydd, double dt, double *q_out,
                         double *yd_out, int *flags, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
                            const bool ok = integrateState(m, q + b * nq, yd + b * nv, ydd + b * nv, dt, q_out + b * nq,
                                                           yd_out + b * nv);
                            if (flags)
                                flags[b] = ok ? 0 : 1; });
    }
    int oracle_count_flops(void *hv, int algo, uint64_t *out)
    {
        Handle *h = (Handle *)hv;
        return guarded([&]
                       {
            ClusterTreeModel<Counter> m = h->spec.instantiate<Counter>();
            const int nq = m.getNumPositions(), nv = m.getNumDegreesOfFreedom();
            std::vector<double> q(nq), yd(nv), aux(nv);
            generateState(*h->model, 0x6772626461ull, 0, q.data(), yd.data(), aux.data());
            Mat<Counter> qc(nq, 1), ydc(nv, 1), auxc(nv, 1);
            for (int i = 0; i < nq; i++) qc[i] = Counter::variable(q[i]);
            for (int i = 0; i < nv; i++) ydc[i] = Counter::variable(yd[i]);
            for (int i = 0; i < nv; i++) auxc[i] = Counter::variable(aux[i]);
            m.setState(qc, ydc);
            op_counts().reset();
            if (algo == 0) m.inverseDynamics(auxc);
            else if (algo == 1) m.forwardDynamics(auxc);
            else if (algo == 2) {
                m.forwardKinematics();
                std::vector<Counter> p(3), R(9), v(6);
                for (int i = 0; i < m.getNumBodies(); i++) m.bodyKinematics(i, p.data(), R.data(), v.data());
            }
            else m.getMassMatrix();
            OpCounts c = op_counts();
            out[0] = c.add_all; out[1] = c.mul_all; out[2] = c.div_all; out[3] = c.sqrt_all; out[4] = c.trig_all;
            out[5] = c.add_alg; out[6] = c.mul_alg; out[7] = c.div_alg; out[8] = c.sqrt_alg; out[9] = c.trig_alg; });
    }

    // ---- derivatives (test infrastructure for grbda_cuda_{inverse,forward}_dynamics_derivatives_f64) -------
    // The reference test's Jacobians (testRigidBodyDynamicsAlgosDerivatives.cpp:126-155): derivative of the third
    // argument's counterpart with respect to a tangent-space perturbation dq (testHelpers.hpp:50-112 `plus`), the
    // velocities and (forward dynamics) tau, by forward-mode dual arithmetic through the restated algorithms.
    // mode 0: out0 = d ID / d dq, out1 = d ID / d yd               (in3 = ydd)
    // mode 1: out0 = d FD / d dq, out1 = d FD / d yd, out2 = d FD / d tau   (in3 = tau)
    // all nv x nv row-major, [i][j] = d out_i / d x_j. Clusters with an implicit constraint move along the
    // constraint manifold (dq_span = G dy; the reference test leaves them out).
    int oracle_dynamics_derivatives(void *hv, int mode, const double *q, const double *yd, const double *in3, double *out0,
                                    double *out1, double *out2, int64_t batch, int threads)
    {
        Handle *h = (Handle *)hv;
        const int nq = h->model->getNumPositions(), nv = h->model->getNumDegreesOfFreedom();
        return runBatch(h, batch, threads, [&](ClusterTreeModel<double> &m, int64_t b)
                        {
            static thread_local std::unique_ptr<ClusterTreeModel<Dual>> md;
            static thread_local uint64_t owner = 0;
            if (owner != h->uid)
            {
                md.reset(new ClusterTreeModel<Dual>(h->spec.instantiate<Dual>()));
                owner = h->uid;
            }
            const double *qb = q + b * nq, *ydb = yd + b * nv, *xb = in3 + b * nv;
            // tangent map T (nq x nv) at this state
            m.setState(toVec(qb, nq), toVec(ydb, nv));
            m.forwardKinematics();
            std::vector<double> T((size_t)nq * nv, 0.0);
            for (auto &n : m.nodes)
            {
                const int pi = n->position_index, vi = n->velocity_index;
                if (std::string(n->joint->typeName()) == "Free" && n->num_positions == 7)
                {
                    Mat<double> quat(4, 1);
                    for (int i = 0; i < 4; i++) quat[i] = qb[pi + 3 + i];
                    const Mat<double> R = quaternionToRotationMatrix(quat);
                    for (int k = 0; k < 3; k++)
                    {
                        const double w[4] = {0.0, k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
                        const double qq[4] = {quat[0], quat[1], quat[2], quat[3]};
                        double prod[4];
                        quatProduct(qq, w, prod);
                        for (int r = 0; r < 4; r++) T[(size_t)(pi + 3 + r) * nv + vi + k] = 0.5 * prod[r];
                        for (int r = 0; r < 3; r++) T[(size_t)(pi + r) * nv + vi + 3 + k] = R(k, r); // R^T e_k
                    }
                }
                else if (n->joint->loop_constraint->isImplicit())
                {
                    const Mat<double> &G = n->joint->G();
                    for (int i = 0; i < n->num_positions; i++)
                        for (int k = 0; k < n->num_velocities; k++)
                            T[(size_t)(pi + i) * nv + vi + k] = G(i, k);
                }
                else
                    for (int k = 0; k < n->num_velocities; k++)
                        T[(size_t)(pi + k) * nv + vi + k] = 1.0;
            }
            Mat<Dual> qd(nq, 1), ydd(nv, 1), xd(nv, 1);
            const int n_kinds = mode == 0 ? 2 : 3;
            double *outs[3] = {out0, out1, out2};
            for (int kind = 0; kind < n_kinds; kind++)
                for (int j = 0; j < nv; j++)
                {
                    for (int i = 0; i < nq; i++) qd[i] = Dual(qb[i], kind == 0 ? T[(size_t)i * nv + j] : 0.0);
                    for (int i = 0; i < nv; i++) ydd[i] = Dual(ydb[i], kind == 1 && i == j ? 1.0 : 0.0);
                    for (int i = 0; i < nv; i++) xd[i] = Dual(xb[i], kind == 2 && i == j ? 1.0 : 0.0);
                    md->setState(qd, ydd);
                    const Mat<Dual> r = mode == 0 ? md->inverseDynamics(xd) : md->forwardDynamics(xd);
                    for (int i = 0; i < nv; i++)
                        outs[kind][(size_t)b * nv * nv + (size_t)i * nv + j] = r[i].d;
                } });
    }

    int oracle_max_threads() { return resolveThreads(0); }
}
