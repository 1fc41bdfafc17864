// Host-side model API. Mirrors the construction interface of the reference's ClusterTreeModel
// (reference: include/grbda/Dynamics/ClusterTreeModel.h:23-228, Body.h:12-43,
// ClusterJoints/*.h, ClusterJoints/Transmissions.h:11-44) so that the reference's robot builders
// translate one to one; but here a cluster joint is a *description* (type tag, axes, G / phi) that
// the device-side model compiler consumes — all per-state arithmetic runs in generated sm_100a
// kernels (no CPU evaluation path exists in the product).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include "../compiler/sym.h"
#include "types.h"

struct grbda_model; // C-ABI handle (include/grbda_cuda.h)

namespace grbda
{
    // reference: include/grbda/Dynamics/Body.h:12-43
    struct Body
    {
        int index_ = -1;
        std::string name_;
        int parent_index_ = -1;
        spatial::Transform Xtree_;
        SpatialInertia inertia_;
        int sub_index_within_cluster_ = 0;
        int cluster_ancestor_index_ = -1;
        int cluster_ancestor_sub_index_within_cluster_ = 0;
    };

    // Serialisable straight-line program for an implicit constraint function phi(q_spanning).
    // ops[i] refers to earlier ops; OP_INPUT reads spanning coordinate `b` of the cluster.
    struct PhiOp
    {
        int32_t op, a, b;
        double val;
    };
    struct PhiProgram
    {
        std::vector<PhiOp> ops;
        std::vector<int32_t> outputs;
        int num_inputs = 0;

        using SymFcn = std::function<std::vector<sym::Sym>(const std::vector<sym::Sym> &)>;

        // Record a function written over sym::Sym (the counterpart of the casadi::SX lambdas in
        // src/Robots/Tello.cpp:139-154,237-252 and ClusterTreeParsing.cpp:310-376).
        static PhiProgram record(int num_inputs, const SymFcn &f)
        {
            sym::Graph g;
            sym::GraphScope scope(g);
            std::vector<sym::Sym> q;
            for (int i = 0; i < num_inputs; i++)
                q.push_back(sym::Sym::input(0, i));
            std::vector<sym::Sym> out = f(q);
            // keep only nodes reachable from the outputs
            std::vector<char> live(g.nodes.size(), 0);
            for (auto &o : out)
                live[o.id] = 1;
            for (int i = (int)g.nodes.size() - 1; i >= 0; i--)
                if (live[i])
                {
                    const sym::Node &n = g.nodes[i];
                    if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                        continue;
                    for (int32_t c : {n.a, n.b})
                        if (c >= 0)
                            live[c] = 1;
                }
            PhiProgram p;
            p.num_inputs = num_inputs;
            std::vector<int32_t> remap(g.nodes.size(), -1);
            for (size_t i = 0; i < g.nodes.size(); i++)
                if (live[i])
                {
                    const sym::Node &n = g.nodes[i];
                    PhiOp op{(int32_t)n.op, n.a, n.b, n.val};
                    if (n.op != sym::OP_CONST && n.op != sym::OP_INPUT)
                    {
                        op.a = n.a >= 0 ? remap[n.a] : -1;
                        op.b = n.b >= 0 ? remap[n.b] : -1;
                    }
                    remap[i] = (int32_t)p.ops.size();
                    p.ops.push_back(op);
                }
            for (auto &o : out)
                p.outputs.push_back(remap[o.id]);
            return p;
        }

        // Replay over any scalar with the sym:: operator set (Sym, or dual numbers over Sym).
        template <typename S>
        std::vector<S> evaluate(const std::vector<S> &q) const
        {
            std::vector<S> v;
            v.reserve(ops.size());
            for (const PhiOp &o : ops)
            {
                switch (o.op)
                {
                case sym::OP_CONST: v.push_back(S(o.val)); break;
                case sym::OP_INPUT: v.push_back(q[o.b]); break;
                case sym::OP_ADD: v.push_back(v[o.a] + v[o.b]); break;
                case sym::OP_SUB: v.push_back(v[o.a] - v[o.b]); break;
                case sym::OP_MUL: v.push_back(v[o.a] * v[o.b]); break;
                case sym::OP_DIV: v.push_back(v[o.a] / v[o.b]); break;
                case sym::OP_NEG: v.push_back(-v[o.a]); break;
                case sym::OP_SIN: v.push_back(sin(v[o.a])); break;
                case sym::OP_COS: v.push_back(cos(v[o.a])); break;
                default: throw std::runtime_error("PhiProgram: unsupported op");
                }
            }
            std::vector<S> out;
            for (int32_t i : outputs)
                out.push_back(v[i]);
            return out;
        }
    };

    enum class ClusterType : int32_t
    {
        FreeQuaternion = 0,  // 1 body, nq 7 [p; quat w,x,y,z], 6 dof [w_body; v_body]
        FreeRollPitchYaw = 1,// 1 body, nq 6 [p; rpy]
        Explicit = 2,        // revolute bodies, constant G: q_span = G y
        Implicit = 3,        // revolute bodies, phi(q_span) = 0; positions given in spanning coords
    };

    // What a ClusterJoints::* object boils down to for the model compiler.
    struct ClusterDesc
    {
        ClusterType type = ClusterType::Explicit;
        std::string joint_type_name;
        int num_bodies = 0, num_positions = 0, num_velocities = 0, num_constraints = 0;
        std::vector<ori::CoordinateAxis> axes; // per body, registration (sub-index) order
        std::vector<double> G, K;              // Explicit: N x n and nc x N, row-major
        std::vector<bool> independent;         // Implicit: per spanning coordinate
        PhiProgram phi;                        // Implicit
    };

    namespace LoopConstraint
    {
        // reference: ClusterJoints/LoopConstraint.h:63-78, LoopConstraint.cpp:39-52
        struct Static
        {
            std::vector<double> G, K;
            int num_spanning = 0, num_independent = 0, num_constraints = 0;
            Static(const std::vector<double> &G_, int n_span, int n_ind, const std::vector<double> &K_,
                   int n_cnstr)
                : G(G_), K(K_), num_spanning(n_span), num_independent(n_ind), num_constraints(n_cnstr)
            {
            }
        };
        // reference: ClusterJoints/GenericJoint.h (GenericImplicit), GenericJoint.cpp:10-109
        struct GenericImplicit
        {
            std::vector<bool> is_coordinate_independent;
            PhiProgram phi;
            GenericImplicit(const std::vector<bool> &independent, const PhiProgram::SymFcn &phi_fcn)
                : is_coordinate_independent(independent),
                  phi(PhiProgram::record((int)independent.size(), phi_fcn)) {}
            GenericImplicit(const std::vector<bool> &independent, const PhiProgram &program)
                : is_coordinate_independent(independent), phi(program) {}
        };
    } // namespace LoopConstraint

    namespace ClusterJoints
    {
        // reference: ClusterJoints/Transmissions.h:11-44
        struct GearedTransmissionModule
        {
            Body body_, rotor_;
            std::string body_joint_name_, rotor_joint_name_;
            ori::CoordinateAxis joint_axis_, rotor_axis_;
            double gear_ratio_;
        };
        struct ParallelBeltTransmissionModule
        {
            Body body_, rotor_;
            ori::CoordinateAxis joint_axis_, rotor_axis_;
            double gear_ratio_;
            std::vector<double> belt_ratios_;
        };
        inline std::vector<double> beltMatrixRowFromBeltRatios(std::vector<double> ratios)
        {
            for (size_t i = 1; i < ratios.size(); ++i)
                ratios[i] = ratios[i - 1] * ratios[i];
            return ratios;
        }

        // Each class below only assembles the ClusterDesc; see the file header.

        // reference: src/Dynamics/ClusterJoints/FreeJoint.cpp:10-27
        struct Free
        {
            ClusterDesc desc;
            explicit Free(const Body &body, bool quaternion = true, std::string = "")
            {
                if (body.parent_index_ >= 0)
                    throw std::runtime_error("Free joint is only valid as the first joint in a tree "
                                             "and thus cannot have a parent body");
                desc.type = quaternion ? ClusterType::FreeQuaternion : ClusterType::FreeRollPitchYaw;
                desc.joint_type_name = "Free";
                desc.num_bodies = 1;
                desc.num_positions = quaternion ? 7 : 6;
                desc.num_velocities = 6;
                desc.axes = {ori::CoordinateAxis::X};
            }
        };

        // reference: RevoluteJoint.cpp:9-24
        struct Revolute
        {
            ClusterDesc desc;
            Revolute(const Body &, ori::CoordinateAxis joint_axis, std::string = "")
            {
                desc.joint_type_name = "Revolute";
                desc.num_bodies = desc.num_positions = desc.num_velocities = 1;
                desc.axes = {joint_axis};
                desc.G = {1.0};
            }
        };

        // reference: RevoluteWithRotorJoint.cpp:9-32   G = [1; N], K = [N, -1]
        struct RevoluteWithRotor
        {
            ClusterDesc desc;
            explicit RevoluteWithRotor(const GearedTransmissionModule &m)
            {
                desc.joint_type_name = "RevoluteWithRotor";
                desc.num_bodies = 2;
                desc.num_positions = desc.num_velocities = 1;
                desc.num_constraints = 1;
                desc.axes = {m.joint_axis_, m.rotor_axis_};
                desc.G = {1.0, m.gear_ratio_};
                desc.K = {m.gear_ratio_, -1.0};
            }
        };

        // reference: RevolutePairJoint.cpp:10-35
        struct RevolutePair
        {
            ClusterDesc desc;
            RevolutePair(const Body &, const Body &, ori::CoordinateAxis a1, ori::CoordinateAxis a2)
            {
                desc.joint_type_name = "RevolutePair";
                desc.num_bodies = desc.num_positions = desc.num_velocities = 2;
                desc.axes = {a1, a2};
                desc.G = {1, 0, 0, 1};
            }
        };

        // reference: RevolutePairWithRotorJoint.cpp:10-70
        struct RevolutePairWithRotor
        {
            ClusterDesc desc;
            RevolutePairWithRotor(const ParallelBeltTransmissionModule &m1,
                                  const ParallelBeltTransmissionModule &m2)
            {
                const int l1 = m1.body_.sub_index_within_cluster_, l2 = m2.body_.sub_index_within_cluster_;
                const int r1 = m1.rotor_.sub_index_within_cluster_, r2 = m2.rotor_.sub_index_within_cluster_;
                desc.joint_type_name = "RevolutePairWithRotor";
                desc.num_bodies = 4;
                desc.num_positions = desc.num_velocities = 2;
                desc.num_constraints = 2;
                desc.axes.resize(4);
                desc.axes[l1] = m1.joint_axis_;
                desc.axes[r1] = m1.rotor_axis_;
                desc.axes[r2] = m2.rotor_axis_;
                desc.axes[l2] = m2.joint_axis_;
                const auto b1 = beltMatrixRowFromBeltRatios(m1.belt_ratios_);
                const auto b2 = beltMatrixRowFromBeltRatios(m2.belt_ratios_);
                desc.G.assign(8, 0.0);
                auto G = [&](int i, int j) -> double & { return desc.G[2 * i + j]; };
                G(l1, 0) = 1.;
                G(r1, 0) = m1.gear_ratio_ * b1.at(0);
                G(r2, 0) = m2.gear_ratio_ * b2.at(0);
                G(r2, 1) = m2.gear_ratio_ * b2.at(1);
                G(l2, 1) = 1.;
                desc.K.assign(8, 0.0);
                auto K = [&](int i, int j) -> double & { return desc.K[4 * i + j]; };
                const int c1 = r1 > r2, c2 = r2 > r1;
                K(c1, r1) = -1.;
                K(c1, l1) = G(r1, 0);
                K(c2, r2) = -1.;
                K(c2, l1) = G(r2, 0);
                K(c2, l2) = G(r2, 1);
            }
        };

        // reference: RevoluteTripleWithRotorJoint.cpp:10-60. Bodies must be registered in the order the
        // reference hard-codes: link 1, 2, 3 (a serial chain), then rotor 1, 2, 3 (on the parent cluster's body).
        // G = [1; diag(gear ratios) * belt matrix] (6 x 3), K = [-G_rotors, 1] (3 x 6).
        struct RevoluteTripleWithRotor
        {
            ClusterDesc desc;
            RevoluteTripleWithRotor(const ParallelBeltTransmissionModule &m1, const ParallelBeltTransmissionModule &m2,
                                    const ParallelBeltTransmissionModule &m3)
            {
                const ParallelBeltTransmissionModule *m[3] = {&m1, &m2, &m3};
                for (int i = 0; i < 3; i++)
                    if (m[i]->body_.sub_index_within_cluster_ != i || m[i]->rotor_.sub_index_within_cluster_ != 3 + i ||
                        (int)m[i]->belt_ratios_.size() != i + 1)
                        throw std::runtime_error("RevoluteTripleWithRotor: bodies must be registered as link 1, 2, 3, "
                                                 "rotor 1, 2, 3 and module i needs i belt ratios");
                desc.joint_type_name = "RevoluteTripleWithRotor";
                desc.num_bodies = 6;
                desc.num_positions = desc.num_velocities = 3;
                desc.num_constraints = 3;
                desc.axes = {m1.joint_axis_, m2.joint_axis_, m3.joint_axis_, m1.rotor_axis_, m2.rotor_axis_, m3.rotor_axis_};
                desc.G.assign(18, 0.0);
                desc.K.assign(18, 0.0);
                for (int i = 0; i < 3; i++)
                {
                    desc.G[3 * i + i] = 1.0;
                    const auto belt = beltMatrixRowFromBeltRatios(m[i]->belt_ratios_);
                    for (int j = 0; j <= i; j++)
                    {
                        desc.G[3 * (3 + i) + j] = m[i]->gear_ratio_ * belt[j];
                        desc.K[6 * i + j] = -desc.G[3 * (3 + i) + j];
                    }
                    desc.K[6 * i + 3 + i] = 1.0;
                }
            }
        };

        // reference: GenericJoint.cpp:243-288. `joint_axes` replaces the vector of
        // Joints::Revolute pointers (only revolute single joints occur inside multi-body clusters,
        // ClusterTreeParsing.cpp:232-258).
        struct Generic
        {
            ClusterDesc desc;
            Generic(const std::vector<Body> &bodies, const std::vector<ori::CoordinateAxis> &joint_axes,
                    const LoopConstraint::Static &lc)
            {
                desc.joint_type_name = "Generic";
                desc.num_bodies = (int)bodies.size();
                if (lc.num_spanning != desc.num_bodies)
                    throw std::runtime_error("Generic: G must have one row per body");
                desc.num_positions = desc.num_velocities = lc.num_independent;
                desc.num_constraints = lc.num_constraints;
                desc.axes = joint_axes;
                desc.G = lc.G;
                desc.K = lc.K;
            }
            Generic(const std::vector<Body> &bodies, const std::vector<ori::CoordinateAxis> &joint_axes,
                    const LoopConstraint::GenericImplicit &lc)
            {
                desc.type = ClusterType::Implicit;
                desc.joint_type_name = "Generic";
                desc.num_bodies = (int)bodies.size();
                if ((int)lc.is_coordinate_independent.size() != desc.num_bodies)
                    throw std::runtime_error("Generic: one independence flag per body is required");
                desc.num_positions = desc.num_bodies; // spanning coordinates (GenericJoint.cpp:246-249)
                desc.num_velocities = 0;
                for (bool b : lc.is_coordinate_independent)
                    desc.num_velocities += b;
                desc.num_constraints = (int)lc.phi.outputs.size();
                if (desc.num_constraints != desc.num_bodies - desc.num_velocities)
                    throw std::runtime_error("Generic: phi must have one row per dependent coordinate");
                desc.axes = joint_axes;
                desc.independent = lc.is_coordinate_independent;
                desc.phi = lc.phi;
            }
        };
    } // namespace ClusterJoints

    // reference: include/grbda/Dynamics/StateRepresentation.h (ContactPoint: body, offset in the body frame,
    // end-effector flag); appended with TreeModel::appendContactPoint / appendEndEffector
    struct ContactPoint
    {
        int body_index_ = -1;
        Vec3 local_offset_ = {0., 0., 0.};
        std::string name_;
        bool is_end_effector_ = false;
    };

    // reference: Nodes/TreeNode.h:16-75 (topology part only)
    struct ClusterTreeNode
    {
        int index_ = 0, parent_index_ = -1, num_parent_bodies_ = 1;
        std::string name_;
        int position_index_ = 0, num_positions_ = 0;
        int velocity_index_ = 0, num_velocities_ = 0;
        int motion_subspace_index_ = 0, motion_subspace_dimension_ = 0;
        int first_body_ = 0;
        std::vector<Body> bodies_;
        ClusterDesc joint_;
    };

    class ClusterTreeModel
    {
    public:
        // reference: ClusterTreeModel.h:27-31
        ClusterTreeModel() { body_name_to_body_index_["ground"] = -1; }
        explicit ClusterTreeModel(const std::string &urdf_filename) : ClusterTreeModel()
        {
            buildModelFromURDF(urdf_filename);
        }

        // reference: ClusterTreeModel.h:48-53
        explicit ClusterTreeModel(const std::vector<std::string> &urdf_filenames) : ClusterTreeModel()
        {
            buildModelFromURDF(urdf_filenames);
        }

        void buildModelFromURDF(const std::string &urdf_filename);                // host/urdf.cpp
        void buildModelFromURDF(const std::vector<std::string> &urdf_filenames); // several files, one model
        // the parse as JSON (link order, parents, children, loop links, supporting chains, clusters): what the
        // reference's parser tests check of the urdfdom fork's ModelInterface
        static std::string describeURDF(const std::vector<std::string> &urdf_filenames);

        // reference: ClusterTreeModel.cpp:9-32
        Body registerBody(const std::string &name, const SpatialInertia &inertia,
                          const std::string &parent_name, const spatial::Transform &Xtree)
        {
            Body body;
            body.index_ = (int)bodies_.size();
            body.name_ = name;
            auto it = body_name_to_body_index_.find(parent_name);
            if (it == body_name_to_body_index_.end())
                throw std::runtime_error("registerBody: unknown parent body '" + parent_name + "'");
            body.parent_index_ = it->second;
            body.Xtree_ = Xtree;
            body.inertia_ = inertia;
            body.sub_index_within_cluster_ = (int)bodies_in_current_cluster_.size();
            // getClusterAncestorIndexFromParent (:407-416)
            int anc = body.parent_index_;
            while (anc != -1 && getIndexOfClusterContainingBodyOrMinus2(anc) == -2)
                anc = bodies_[anc].parent_index_;
            body.cluster_ancestor_index_ = anc;
            body.cluster_ancestor_sub_index_within_cluster_ =
                anc >= 0 ? bodies_[anc].sub_index_within_cluster_ : 0;
            body_name_to_body_index_[name] = body.index_;
            bodies_.push_back(body);
            bodies_in_current_cluster_.push_back(body);
            return body;
        }

        // reference: ClusterTreeModel.h:61-66
        template <typename ClusterJointType, typename... Args>
        void appendRegisteredBodiesAsCluster(const std::string &name, Args &&...args)
        {
            ClusterJointType joint(std::forward<Args>(args)...);
            appendCluster(name, joint.desc);
        }

        // reference: ClusterTreeModel.h:69-78
        template <typename ClusterJointType, typename... Args>
        void appendBody(const std::string &name, const SpatialInertia &inertia,
                        const std::string &parent_name, const spatial::Transform &Xtree,
                        Args &&...args)
        {
            Body body = registerBody(name, inertia, parent_name, Xtree);
            ClusterJointType joint(body, std::forward<Args>(args)...);
            appendCluster(name, joint.desc);
        }

        // reference: ClusterTreeModel.cpp:34-67
        void appendCluster(const std::string &name, const ClusterDesc &joint)
        {
            if ((int)bodies_in_current_cluster_.size() != joint.num_bodies)
                throw std::runtime_error("appendRegisteredBodiesAsCluster: cluster joint '" + name +
                                         "' does not match the number of registered bodies");
            // getIndexOfParentClusterFromBodies (:461-489); ground maps to cluster -1
            int parent_cluster_index = -2;
            for (const Body &b : bodies_in_current_cluster_)
            {
                const int c = getIndexOfClusterContainingBodyOrMinus2(b.parent_index_);
                if (c != -2)
                {
                    parent_cluster_index = c;
                    break;
                }
            }
            if (parent_cluster_index == -2)
                throw std::runtime_error("At least one body in every cluster must have a parent in "
                                         "a different clusters");

            ClusterTreeNode node;
            node.index_ = (int)cluster_nodes_.size();
            node.name_ = name;
            node.parent_index_ = parent_cluster_index;
            node.num_parent_bodies_ =
                parent_cluster_index >= 0 ? (int)cluster_nodes_[parent_cluster_index].bodies_.size() : 1;
            node.bodies_ = bodies_in_current_cluster_;
            node.first_body_ = bodies_in_current_cluster_.front().index_;
            node.joint_ = joint;
            node.position_index_ = position_index_;
            node.num_positions_ = joint.num_positions;
            node.velocity_index_ = velocity_index_;
            node.num_velocities_ = joint.num_velocities;
            node.motion_subspace_index_ = motion_subspace_index_;
            node.motion_subspace_dimension_ = 6 * joint.num_bodies;
            cluster_name_to_cluster_index_[name] = node.index_;
            cluster_nodes_.push_back(node);
            for (const Body &b : node.bodies_)
                body_index_to_cluster_index_[b.index_] = node.index_;

            // checkValidParentClusterForBodiesInCluster (:112-127)
            for (const Body &b : node.bodies_)
            {
                const int other = getIndexOfClusterContainingBodyOrMinus2(b.parent_index_);
                if (other != node.index_ && other != parent_cluster_index)
                    throw std::runtime_error("The parents of all bodies in a cluster must have "
                                             "parents in the current cluster OR in the same parent "
                                             "cluster");
            }
            position_index_ += joint.num_positions;
            velocity_index_ += joint.num_velocities;
            motion_subspace_index_ += node.motion_subspace_dimension_;
            bodies_in_current_cluster_.clear();
            device_model_.reset();
        }

        // ---- batched extension of the model interface (host/batched.cpp) -------------------------------
        // Counterparts of setState + inverseDynamics / forwardDynamics / getMassMatrix / forwardKinematics
        // + getters (reference: include/grbda/Dynamics/ClusterTreeModel.h:91-97,143-165, TreeModel.h:62)
        // over `batch` states at once. Arrays are DEVICE pointers on the CUDA device that is current when
        // the first batched call is made, one state per column, contiguous per state: exactly the vectors
        // setState() takes, concatenated (spanning positions for implicit clusters). Calls are
        // asynchronous on `stream` (a cudaStream_t). Errors throw std::runtime_error like the reference.
        // The device-side model (kernels compiled ahead of time, or by NVRTC on first use) is created
        // lazily from the current topology and dropped when the model is modified.
        void inverseDynamicsBatch(const double *q, const double *yd, const double *ydd, double *tau, int64_t batch,
                                  void *stream = nullptr) const;
        void forwardDynamicsBatch(const double *q, const double *yd, const double *tau, double *ydd, int64_t batch,
                                  void *stream = nullptr) const;
        void massMatrixBatch(const double *q, double *H, int64_t batch, void *stream = nullptr) const;
        // p[3 Nb], R[9 Nb] (row-major body-to-world), v[6 Nb] ([world angular; world linear]) per state
        void forwardKinematicsBatch(const double *q, const double *yd, double *p, double *R, double *v, int64_t batch,
                                    void *stream = nullptr) const;
        // C(q, yd) = inverseDynamics with ydd = 0 (getBiasForceVector, ClusterTreeModel.h:165); `zeros` is a
        // device array of batch * nv zeros provided by the caller
        void biasForceBatch(const double *q, const double *yd, const double *zeros, double *C, int64_t batch,
                            void *stream = nullptr) const;
        // random valid states for global indices [first_index, first_index + count) (ClusterJoint.cpp:74-81,
        // FreeJoint.cpp:49-60, GenericJoint.cpp:290-385); aux = nv uniform values (a random ydd or tau)
        void randomStatesBatch(uint64_t seed, int64_t first_index, int64_t count, double *q, double *yd, double *aux,
                               void *stream = nullptr) const;
        // the C-ABI handle behind the batched methods (include/grbda_cuda.h), e.g. for the *_ext entry points
        ::grbda_model *deviceModel() const;

        // reference: TreeModel.h:56
        void setGravity(const Vec3 &g)
        {
            gravity_ = g;
            device_model_.reset();
        }
        const Vec3 &getGravity() const { return gravity_; }

        // reference: ClusterTreeModel::appendContactPoint / appendEndEffector (src/Dynamics/ClusterTreeModel.cpp:129-190)
        void appendContactPoint(const std::string &body_name, const Vec3 &local_offset, const std::string &cp_name,
                                bool is_end_effector = false)
        {
            auto it = body_name_to_body_index_.find(body_name);
            if (it == body_name_to_body_index_.end() || it->second < 0)
                throw std::runtime_error("appendContactPoint: unknown body '" + body_name + "'");
            for (const ContactPoint &c : contact_points_)
                if (c.name_ == cp_name)
                    throw std::runtime_error("appendContactPoint: contact point '" + cp_name + "' exists already");
            contact_points_.push_back(ContactPoint{it->second, local_offset, cp_name, is_end_effector});
            device_model_.reset();
        }
        void appendEndEffector(const std::string &body_name, const Vec3 &local_offset, const std::string &cp_name)
        {
            appendContactPoint(body_name, local_offset, cp_name, true);
        }
        void setContactPoints(const std::vector<ContactPoint> &points)
        {
            for (const ContactPoint &c : points)
                if (c.body_index_ < 0 || c.body_index_ >= (int)bodies_.size())
                    throw std::runtime_error("setContactPoints: body index out of range");
            contact_points_ = points;
            device_model_.reset();
        }
        const std::vector<ContactPoint> &contactPoints() const { return contact_points_; }
        int getNumEndEffectors() const
        {
            int n = 0;
            for (const ContactPoint &c : contact_points_)
                n += c.is_end_effector_;
            return n;
        }

        // Bodies that take external forces in the batched *_ext entry points (TreeModel::setExternalForces,
        // TreeModel.cpp:215-239, names any set of bodies per call; the batched path fixes the set per model so
        // that the force programs can be specialised). Empty: the default set, every terminal link.
        void setExternalForceBodies(const std::vector<int> &bodies)
        {
            std::vector<char> seen(bodies_.size(), 0);
            for (int b : bodies)
            {
                if (b < 0 || b >= (int)bodies_.size() || seen[b])
                    throw std::runtime_error("setExternalForceBodies: body indices must be distinct and in range");
                seen[b] = 1;
            }
            external_force_bodies_ = bodies;
            device_model_.reset();
        }
        const std::vector<int> &externalForceBodies() const { return external_force_bodies_; }

        // reference: TreeModel.h:25-28, ClusterTreeModel.h:97
        int getNumPositions() const { return position_index_; }
        int getNumDegreesOfFreedom() const { return velocity_index_; }
        int getNumBodies() const { return (int)bodies_.size(); }
        int getNumClusters() const { return (int)cluster_nodes_.size(); }

        const std::vector<Body> &bodies() const { return bodies_; }
        const std::vector<ClusterTreeNode> &clusters() const { return cluster_nodes_; }
        const Body &body(const std::string &name) const
        {
            return bodies_.at(body_name_to_body_index_.at(name));
        }
        int getIndexOfClusterContainingBody(int body_index) const
        {
            const int c = getIndexOfClusterContainingBodyOrMinus2(body_index);
            if (c == -2)
                throw std::runtime_error("Body is not found in any registered cluster");
            return c;
        }

    private:
        int getIndexOfClusterContainingBodyOrMinus2(int body_index) const
        {
            if (body_index == -1)
                return -1;
            auto it = body_index_to_cluster_index_.find(body_index);
            return it == body_index_to_cluster_index_.end() ? -2 : it->second;
        }

        std::vector<Body> bodies_;
        std::vector<Body> bodies_in_current_cluster_;
        std::vector<ClusterTreeNode> cluster_nodes_;
        std::map<std::string, int> body_name_to_body_index_;
        std::map<std::string, int> cluster_name_to_cluster_index_;
        std::map<int, int> body_index_to_cluster_index_;
        int position_index_ = 0, velocity_index_ = 0, motion_subspace_index_ = 0;
        Vec3 gravity_ = {0., 0., -9.81};
        std::vector<int> external_force_bodies_;
        std::vector<ContactPoint> contact_points_;
        mutable std::shared_ptr<void> device_model_; // grbda_model handle + deleter; copies of the model share it
    };

} // namespace grbda
