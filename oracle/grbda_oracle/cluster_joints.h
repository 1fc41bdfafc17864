// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// Cluster joints and loop constraints restated from the reference:
//   include/grbda/Dynamics/Body.h, Joints/Joint.h, ClusterJoints/*.h
//   src/Dynamics/ClusterJoints/{ClusterJoint,LoopConstraint,RevoluteJoint,FreeJoint,
//   RevoluteWithRotorJoint,RevolutePairJoint,RevolutePairWithRotorJoint,GenericJoint}.cpp
#pragma once
#include <functional>
#include <memory>
#include <string>
#include "spatial.h"
#include "taylor2.h"

namespace grbda_oracle
{
    // reference: include/grbda/Dynamics/Body.h:12-43
    template <typename T>
    struct Body
    {
        int index;
        std::string name;
        int parent_index;
        Transform<T> Xtree;
        Mat<T> inertia; // 6x6
        int sub_index_within_cluster;
        int cluster_ancestor_index;
        int cluster_ancestor_sub_index_within_cluster;
    };

    // reference: include/grbda/Utils/StateRepresentation.h:9-57
    template <typename T>
    struct JointState
    {
        Mat<T> position;
        bool position_is_spanning = false;
        Mat<T> velocity;
        bool velocity_is_spanning = false;
    };

    ////////////////////////////////////////////////////////////////////////////////////////////
    // Loop constraints  (reference: ClusterJoints/LoopConstraint.h:14-100, LoopConstraint.cpp)
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    struct LoopConstraintBase
    {
        Mat<T> G, g, K, k;
        virtual ~LoopConstraintBase() {}
        virtual bool isImplicit() const { return false; }
        bool isExplicit() const { return !isImplicit(); }
        virtual int numSpanningPos() const { return G.r; }
        virtual int numIndependentPos() const { return G.c; }
        int numIndependentVel() const { return G.c; }
        int numConstraints() const { return K.r; }
        virtual void updateJacobians(const Mat<T> &) {}
        virtual void updateBiases(const Mat<T> &, const Mat<T> &) {}
        virtual Mat<T> gamma(const Mat<T> &y) const = 0;
        virtual Mat<T> phi(const Mat<T> &) const { return Mat<T>(0, 1); }
    };

    // reference: LoopConstraint.cpp:39-52
    template <typename T>
    struct StaticConstraint : LoopConstraintBase<T>
    {
        StaticConstraint(const Mat<T> &G_, const Mat<T> &K_)
        {
            this->G = G_;
            this->g = Mat<T>(G_.r, 1);
            this->K = K_;
            this->k = Mat<T>(K_.r, 1);
        }
        Mat<T> gamma(const Mat<T> &y) const override { return this->G * y; }
    };

    // reference: ClusterJoints/FreeJoint.h (LoopConstraint::Free): G = I6, K empty, gamma = identity
    template <typename T>
    struct FreeConstraint : LoopConstraintBase<T>
    {
        int nq;
        explicit FreeConstraint(int num_positions) : nq(num_positions)
        {
            this->G = Mat<T>::Identity(6);
            this->g = Mat<T>(6, 1);
            this->K = Mat<T>(0, 6);
            this->k = Mat<T>(0, 1);
        }
        int numSpanningPos() const override { return nq; }
        int numIndependentPos() const override { return nq; }
        Mat<T> gamma(const Mat<T> &y) const override { return y; }
    };

    // reference: GenericJoint.cpp:10-129 (LoopConstraint::GenericImplicit). K, k by Taylor
    // arithmetic instead of CasADi (see taylor2.h); G = P [I; -Kd^-1 Ki], g = P [0; Kd^-1 k].
    template <typename T>
    struct GenericImplicitConstraint : LoopConstraintBase<T>
    {
        using S = Taylor2<T>;
        using PhiFcn = std::function<std::vector<S>(const std::vector<S> &)>;
        std::vector<bool> is_independent;
        std::vector<int> ind_coords, dep_coords;
        PhiFcn phi_fcn;

        GenericImplicitConstraint(const std::vector<bool> &is_coordinate_independent, PhiFcn f)
            : is_independent(is_coordinate_independent), phi_fcn(f)
        {
            const int n = (int)is_independent.size();
            for (int i = 0; i < n; i++)
                (is_independent[i] ? ind_coords : dep_coords).push_back(i);
            std::vector<S> q0(n);
            const int nc = (int)phi_fcn(q0).size();
            this->K = Mat<T>(nc, n);
            this->G = Mat<T>(n, (int)ind_coords.size());
            this->k = Mat<T>(nc, 1);
            this->g = Mat<T>(n, 1);
        }
        bool isImplicit() const override { return true; }
        Mat<T> gamma(const Mat<T> &) const override
        {
            throw std::runtime_error("GenericImplicit::gamma() not implemented");
        }
        Mat<T> phi(const Mat<T> &q) const override
        {
            std::vector<S> qs(q.r);
            for (int i = 0; i < q.r; i++)
                qs[i] = S(q[i], T(0.0), T(0.0));
            std::vector<S> p = phi_fcn(qs);
            Mat<T> out((int)p.size(), 1);
            for (size_t i = 0; i < p.size(); i++)
                out[i] = p[i].c0;
            return out;
        }
        Mat<T> Kd() const
        {
            Mat<T> m(this->K.r, (int)dep_coords.size());
            for (int i = 0; i < m.r; i++)
                for (int j = 0; j < m.c; j++)
                    m(i, j) = this->K(i, dep_coords[j]);
            return m;
        }
        Mat<T> Ki() const
        {
            Mat<T> m(this->K.r, (int)ind_coords.size());
            for (int i = 0; i < m.r; i++)
                for (int j = 0; j < m.c; j++)
                    m(i, j) = this->K(i, ind_coords[j]);
            return m;
        }
        // GenericJoint.cpp:118-122
        void updateJacobians(const Mat<T> &q) override
        {
            const int n = q.r;
            for (int j = 0; j < n; j++)
            {
                std::vector<S> qs(n);
                for (int i = 0; i < n; i++)
                    qs[i] = S(q[i], i == j ? T(1.0) : T(0.0), T(0.0));
                std::vector<S> p = phi_fcn(qs);
                for (size_t i = 0; i < p.size(); i++)
                    this->K(i, j) = p[i].c1;
            }
            Mat<T> Gd = -solve(Kd(), Ki());
            this->G.setZero();
            for (size_t i = 0; i < ind_coords.size(); i++)
                this->G(ind_coords[i], i) = T(1.0);
            for (size_t i = 0; i < dep_coords.size(); i++)
                for (int j = 0; j < Gd.c; j++)
                    this->G(dep_coords[i], j) = Gd(i, j);
        }
        // GenericJoint.cpp:124-129
        void updateBiases(const Mat<T> &q, const Mat<T> &qd) override
        {
            const int n = q.r;
            std::vector<S> qs(n);
            for (int i = 0; i < n; i++)
                qs[i] = S(q[i], qd[i], T(0.0));
            std::vector<S> p = phi_fcn(qs);
            for (size_t i = 0; i < p.size(); i++)
                this->k[i] = -(T(2.0) * p[i].c2);
            Mat<T> gd = solve(Kd(), this->k);
            this->g.setZero();
            for (size_t i = 0; i < dep_coords.size(); i++)
                this->g[dep_coords[i]] = gd[i];
        }
    };

    // reference: FourBarJoint.cpp:7-199 (LoopConstraint::FourBar): planar four-bar linkage, spanning coordinates
    // {path-1 joint 0, path-2 joint 0, path-1 joint 1}. The reference codes K, G, k and g in closed form; they are
    // restated the same way (NOT through the Taylor arithmetic of GenericImplicit), so a model built on this
    // constraint checks the generic loop machinery independently.
    template <typename T>
    struct FourBarConstraint : LoopConstraintBase<T>
    {
        std::vector<T> l1, l2; // path link lengths (two links on path 1, one on path 2)
        T off[2];
        int independent_coordinate;
        Mat<T> map;    // indepenent_coordinate_map_ (FourBarJoint.cpp:57-77)
        Mat<T> Kd_mat; // dependent columns of K at the last updateJacobians()

        FourBarConstraint(const std::vector<T> &path1, const std::vector<T> &path2, T off_x, T off_y, int ind)
            : l1(path1), l2(path2), independent_coordinate(ind), map(3, 3)
        {
            if (l1.size() + l2.size() != 3 || l1.size() != 2)
                throw std::runtime_error("FourBar: Must contain 3 links");
            off[0] = off_x, off[1] = off_y;
            this->G = Mat<T>(3, 1);
            this->g = Mat<T>(3, 1);
            this->K = Mat<T>(2, 3);
            this->k = Mat<T>(2, 1);
            const int rows[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}}; // row r of the map has its 1 in column rows[ind][r]
            if (ind < 0 || ind > 2)
                throw std::runtime_error("FourBar: Invalid independent coordinate");
            for (int r = 0; r < 3; r++)
                map(r, rows[ind][r]) = T(1.0);
        }
        bool isImplicit() const override { return true; }
        Mat<T> gamma(const Mat<T> &) const override { throw std::runtime_error("FourBar::gamma() not implemented"); }
        // FourBarJoint.cpp:22-50
        Mat<T> phi(const Mat<T> &q) const override
        {
            const T q1[2] = {q[0], q[2]}, q2[1] = {q[1]};
            T a(0.0), p1x(0.0), p1y(0.0);
            for (int i = 0; i < 2; i++)
            {
                a = a + q1[i];
                p1x = p1x + l1[i] * cos(a);
                p1y = p1y + l1[i] * sin(a);
            }
            a = T(0.0);
            T p2x = off[0], p2y = off[1];
            for (int i = 0; i < 1; i++)
            {
                a = a + q2[i];
                p2x = p2x + l2[i] * cos(a);
                p2y = p2y + l2[i] * sin(a);
            }
            Mat<T> out(2, 1);
            out[0] = p1x - p2x, out[1] = p1y - p2y;
            return out;
        }
        // FourBarJoint.cpp:81-146
        void updateJacobians(const Mat<T> &q) override
        {
            const T q1[2] = {q[0], q[2]}, q2[1] = {q[1]};
            Mat<T> K1(2, 2), K2(2, 1);
            T a(0.0);
            for (int i = 0; i < 2; i++)
            {
                a = a + q1[i];
                for (int j = 0; j <= i; j++)
                {
                    K1(0, j) = K1(0, j) - l1[i] * sin(a);
                    K1(1, j) = K1(1, j) + l1[i] * cos(a);
                }
            }
            a = T(0.0);
            for (int i = 0; i < 1; i++)
            {
                a = a + q2[i];
                for (int j = 0; j <= i; j++)
                {
                    K2(0, j) = K2(0, j) + l2[i] * sin(a);
                    K2(1, j) = K2(1, j) - l2[i] * cos(a);
                }
            }
            for (int r = 0; r < 2; r++)
                this->K(r, 0) = K1(r, 0), this->K(r, 1) = K2(r, 0), this->K(r, 2) = K1(r, 1);
            Mat<T> Ki(2, 1);
            Kd_mat = Mat<T>(2, 2);
            int dep = 0;
            for (int i = 0; i < 3; i++)
            {
                if (i == independent_coordinate)
                    Ki(0, 0) = this->K(0, i), Ki(1, 0) = this->K(1, i);
                else
                {
                    Kd_mat(0, dep) = this->K(0, i), Kd_mat(1, dep) = this->K(1, i);
                    dep++;
                }
            }
            const Mat<T> Gd = -solve(Kd_mat, Ki);
            Mat<T> Gt(3, 1);
            Gt[0] = T(1.0), Gt[1] = Gd[0], Gt[2] = Gd[1];
            this->G = map * Gt;
        }
        // FourBarJoint.cpp:148-199
        void updateBiases(const Mat<T> &q, const Mat<T> &qd) override
        {
            const T q1[2] = {q[0], q[2]}, q2[1] = {q[1]}, qd1[2] = {qd[0], qd[2]}, qd2[1] = {qd[1]};
            Mat<T> Kd1(2, 2), Kd2(2, 1);
            T a(0.0), w(0.0);
            for (int i = 0; i < 2; i++)
            {
                a = a + q1[i];
                w = w + qd1[i];
                for (int j = 0; j <= i; j++)
                {
                    Kd1(0, j) = Kd1(0, j) - l1[i] * w * cos(a);
                    Kd1(1, j) = Kd1(1, j) - l1[i] * w * sin(a);
                }
            }
            a = T(0.0), w = T(0.0);
            for (int i = 0; i < 1; i++)
            {
                a = a + q2[i];
                w = w + qd2[i];
                for (int j = 0; j <= i; j++)
                {
                    Kd2(0, j) = Kd2(0, j) + l2[i] * w * cos(a);
                    Kd2(1, j) = Kd2(1, j) + l2[i] * w * sin(a);
                }
            }
            Mat<T> Kdot(2, 3);
            for (int r = 0; r < 2; r++)
                Kdot(r, 0) = Kd1(r, 0), Kdot(r, 1) = Kd2(r, 0), Kdot(r, 2) = Kd1(r, 1);
            this->k = -(Kdot * qd);
            const Mat<T> gd = solve(Kd_mat, this->k);
            Mat<T> gt(3, 1);
            gt[0] = T(0.0), gt[1] = gd[0], gt[2] = gd[1];
            this->g = map * gt;
        }
    };

    ////////////////////////////////////////////////////////////////////////////////////////////
    // Single joints  (reference: include/grbda/Dynamics/Joints/Joint.h:43-102)
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    struct SingleJoint
    {
        Mat<T> S;
        Transform<T> XJ;
        int num_positions = 1, num_velocities = 1;
        virtual ~SingleJoint() {}
        virtual void updateKinematics(const Mat<T> &q) = 0;
    };

    template <typename T>
    struct SingleRevolute : SingleJoint<T>
    {
        Axis axis;
        explicit SingleRevolute(Axis a) : axis(a)
        {
            // Spatial.h:296-316 (jointMotionSubspace, revolute)
            this->S = Mat<T>(6, 1);
            this->S((int)a, 0) = T(1.0);
        }
        void updateKinematics(const Mat<T> &q) override
        {
            this->XJ = Transform<T>(coordinateRotation(axis, q[0]));
        }
    };

    template <typename T>
    struct SingleFree : SingleJoint<T>
    {
        bool quaternion;
        explicit SingleFree(bool use_quaternion) : quaternion(use_quaternion)
        {
            this->S = Mat<T>::Identity(6);
            this->num_positions = quaternion ? 7 : 6;
            this->num_velocities = 6;
        }
        // Joint.h:61-68, OrientationRepresentation.h:11-49
        void updateKinematics(const Mat<T> &q) override
        {
            Mat<T> R = quaternion ? quaternionToRotationMatrix(q.segment(3, 4))
                                  : rpyToRotMat(q.segment(3, 3));
            this->XJ = Transform<T>(R, q.segment(0, 3));
        }
    };

    ////////////////////////////////////////////////////////////////////////////////////////////
    // Cluster joints
    ////////////////////////////////////////////////////////////////////////////////////////////
    // reference: ClusterJoints/ClusterJoint.h:33-98, ClusterJoint.cpp:10-81
    template <typename T>
    struct ClusterJointBase
    {
        int num_bodies, num_positions, num_velocities;
        Mat<T> S, vJ, cJ;
        std::shared_ptr<LoopConstraintBase<T>> loop_constraint;
        std::vector<std::shared_ptr<SingleJoint<T>>> single_joints;
        bool last_position_valid = true;

        ClusterJointBase(int nb, int np, int nv)
            : num_bodies(nb), num_positions(np), num_velocities(nv),
              S(6 * nb, nv), vJ(6 * nb, 1), cJ(6 * nb, 1) {}
        virtual ~ClusterJointBase() {}

        virtual void updateKinematics(const JointState<T> &joint_state) = 0;
        virtual void computeXup(GeneralizedTransform<T> &Xup) const = 0;
        virtual const char *typeName() const = 0;

        const Mat<T> &G() const { return loop_constraint->G; }
        const Mat<T> &g() const { return loop_constraint->g; }
        const Mat<T> &K() const { return loop_constraint->K; }
        const Mat<T> &k() const { return loop_constraint->k; }

        // ClusterJoint.cpp:23-71. The reference throws on an invalid spanning position; the
        // oracle records it (last_position_valid) so that batch drivers can report it.
        JointState<T> toSpanningTreeState(const JointState<T> &js)
        {
            JointState<T> out;
            out.position_is_spanning = out.velocity_is_spanning = true;
            if (!js.position_is_spanning && loop_constraint->isExplicit())
                out.position = loop_constraint->gamma(js.position);
            else if (!js.position_is_spanning && loop_constraint->isImplicit())
                throw std::runtime_error("Independent positions cannot be converted to spanning "
                                         "positions when the constraint is implicit.");
            else if (js.position_is_spanning && loop_constraint->isExplicit())
                out.position = js.position;
            else
            {
                last_position_valid = loop_constraint->phi(js.position).norm() < 1e-8;
                out.position = js.position;
            }
            loop_constraint->updateJacobians(out.position);
            if (!js.velocity_is_spanning)
                out.velocity = G() * js.velocity;
            else
                out.velocity = js.velocity;
            loop_constraint->updateBiases(out.position, out.velocity);
            return out;
        }
    };

    // reference: RevoluteJoint.cpp:9-41
    template <typename T>
    struct RevoluteCluster : ClusterJointBase<T>
    {
        Body<T> body;
        RevoluteCluster(const Body<T> &b, Axis axis) : ClusterJointBase<T>(1, 1, 1), body(b)
        {
            this->single_joints.push_back(std::make_shared<SingleRevolute<T>>(axis));
            this->S = this->single_joints[0]->S;
            this->loop_constraint =
                std::make_shared<StaticConstraint<T>>(Mat<T>::Identity(1), Mat<T>(0, 1));
        }
        const char *typeName() const override { return "Revolute"; }
        void updateKinematics(const JointState<T> &js) override
        {
            this->single_joints[0]->updateKinematics(js.position);
            this->vJ = this->S * js.velocity;
        }
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            Xup.X[0] = this->single_joints[0]->XJ * body.Xtree;
        }
    };

    // reference: FreeJoint.cpp:10-46
    template <typename T>
    struct FreeCluster : ClusterJointBase<T>
    {
        Body<T> body;
        FreeCluster(const Body<T> &b, bool quaternion)
            : ClusterJointBase<T>(1, quaternion ? 7 : 6, 6), body(b)
        {
            if (b.parent_index >= 0)
                throw std::runtime_error("Free joint is only valid as the first joint in a tree "
                                         "and thus cannot have a parent body");
            this->S = Mat<T>::Identity(6);
            this->single_joints.push_back(std::make_shared<SingleFree<T>>(quaternion));
            this->loop_constraint = std::make_shared<FreeConstraint<T>>(quaternion ? 7 : 6);
        }
        const char *typeName() const override { return "Free"; }
        void updateKinematics(const JointState<T> &js) override
        {
            this->single_joints[0]->updateKinematics(js.position);
            this->vJ = this->S * js.velocity;
        }
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            Xup.X[0] = this->single_joints[0]->XJ; // Xtree ignored (FreeJoint.cpp:45)
        }
    };

    // reference: ClusterJoints/Transmissions.h:11-44
    template <typename T>
    struct GearedTransmissionModule
    {
        Body<T> body, rotor;
        Axis joint_axis, rotor_axis;
        T gear_ratio;
    };
    template <typename T>
    struct ParallelBeltTransmissionModule
    {
        Body<T> body, rotor;
        Axis joint_axis, rotor_axis;
        T gear_ratio;
        std::vector<T> belt_ratios;
    };
    template <typename T>
    std::vector<T> beltMatrixRowFromBeltRatios(std::vector<T> ratios)
    {
        for (size_t i = 1; i < ratios.size(); ++i)
            ratios[i] = ratios[i - 1] * ratios[i];
        return ratios;
    }

    // reference: RevoluteWithRotorJoint.cpp:9-58
    template <typename T>
    struct RevoluteWithRotorCluster : ClusterJointBase<T>
    {
        Body<T> link, rotor;
        explicit RevoluteWithRotorCluster(const GearedTransmissionModule<T> &m)
            : ClusterJointBase<T>(2, 1, 1), link(m.body), rotor(m.rotor)
        {
            this->single_joints.push_back(std::make_shared<SingleRevolute<T>>(m.joint_axis));
            this->single_joints.push_back(std::make_shared<SingleRevolute<T>>(m.rotor_axis));
            Mat<T> G(2, 1), K(1, 2);
            G(0, 0) = T(1.0);
            G(1, 0) = m.gear_ratio;
            K(0, 0) = m.gear_ratio;
            K(0, 1) = T(-1.0);
            this->loop_constraint = std::make_shared<StaticConstraint<T>>(G, K);
            this->S.setBlock(0, 0, this->single_joints[0]->S);
            this->S.setBlock(6, 0, m.gear_ratio * this->single_joints[1]->S);
        }
        const char *typeName() const override { return "RevoluteWithRotor"; }
        void updateKinematics(const JointState<T> &js) override
        {
            JointState<T> s = this->toSpanningTreeState(js);
            this->single_joints[0]->updateKinematics(s.position.segment(0, 1));
            this->single_joints[1]->updateKinematics(s.position.segment(1, 1));
            this->vJ.setSegment(0, this->single_joints[0]->S * s.velocity[0]);
            this->vJ.setSegment(6, this->single_joints[1]->S * s.velocity[1]);
        }
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            Xup.X[0] = this->single_joints[0]->XJ * link.Xtree;
            Xup.X[1] = this->single_joints[1]->XJ * rotor.Xtree;
        }
    };

    // reference: RevolutePairJoint.cpp:10-71
    template <typename T>
    struct RevolutePairCluster : ClusterJointBase<T>
    {
        Body<T> link1, link2;
        Mat<T> X_intra_S_span, X_intra_S_span_ring;
        RevolutePairCluster(const Body<T> &l1, const Body<T> &l2, Axis a1, Axis a2)
            : ClusterJointBase<T>(2, 2, 2), link1(l1), link2(l2),
              X_intra_S_span(12, 2), X_intra_S_span_ring(12, 2)
        {
            this->single_joints.push_back(std::make_shared<SingleRevolute<T>>(a1));
            this->single_joints.push_back(std::make_shared<SingleRevolute<T>>(a2));
            this->loop_constraint =
                std::make_shared<StaticConstraint<T>>(Mat<T>::Identity(2), Mat<T>(0, 2));
            X_intra_S_span.setBlock(0, 0, this->single_joints[0]->S);
            X_intra_S_span.setBlock(6, 1, this->single_joints[1]->S);
            this->S = X_intra_S_span * this->G();
        }
        const char *typeName() const override { return "RevolutePair"; }
        void updateKinematics(const JointState<T> &js) override
        {
            JointState<T> s = this->toSpanningTreeState(js);
            const Mat<T> &q = s.position, &qd = s.velocity;
            this->single_joints[0]->updateKinematics(q.segment(0, 1));
            this->single_joints[1]->updateKinematics(q.segment(1, 1));
            Transform<T> X21 = this->single_joints[1]->XJ * link2.Xtree;
            Mat<T> v2_relative = this->single_joints[1]->S * qd[1];
            Mat<T> X21S1 = X21.transformMotionVector(this->single_joints[0]->S);
            X_intra_S_span.setBlock(6, 0, X21S1);
            this->S.setBlock(6, 0, X21S1);
            X_intra_S_span_ring.setBlock(6, 0, -(motionCrossMatrix(v2_relative) * X21S1));
            this->vJ = X_intra_S_span * qd;
            this->cJ = X_intra_S_span_ring * qd;
        }
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            Xup.X[0] = this->single_joints[0]->XJ * link1.Xtree;
            Xup.X[1] = this->single_joints[1]->XJ * link2.Xtree * Xup.X[0];
        }
    };

    // reference: RevolutePairWithRotorJoint.cpp:10-116
    template <typename T>
    struct RevolutePairWithRotorCluster : ClusterJointBase<T>
    {
        Body<T> link1, link2, rotor1, rotor2;
        int link1_index, link2_index, rotor1_index, rotor2_index;
        std::shared_ptr<SingleJoint<T>> link1_joint, rotor1_joint, rotor2_joint, link2_joint;
        Mat<T> X_intra_S_span, X_intra_S_span_ring;

        RevolutePairWithRotorCluster(const ParallelBeltTransmissionModule<T> &m1,
                                     const ParallelBeltTransmissionModule<T> &m2)
            : ClusterJointBase<T>(4, 2, 2), link1(m1.body), link2(m2.body), rotor1(m1.rotor),
              rotor2(m2.rotor), link1_index(m1.body.sub_index_within_cluster),
              link2_index(m2.body.sub_index_within_cluster),
              rotor1_index(m1.rotor.sub_index_within_cluster),
              rotor2_index(m2.rotor.sub_index_within_cluster),
              X_intra_S_span(24, 4), X_intra_S_span_ring(24, 4)
        {
            link1_joint = std::make_shared<SingleRevolute<T>>(m1.joint_axis);
            rotor1_joint = std::make_shared<SingleRevolute<T>>(m1.rotor_axis);
            rotor2_joint = std::make_shared<SingleRevolute<T>>(m2.rotor_axis);
            link2_joint = std::make_shared<SingleRevolute<T>>(m2.joint_axis);
            this->single_joints = {link1_joint, rotor1_joint, rotor2_joint, link2_joint};

            // ratio_product = diag(gear ratios) * [belt row 1, 0; belt row 2]
            std::vector<T> b1 = beltMatrixRowFromBeltRatios(m1.belt_ratios);
            std::vector<T> b2 = beltMatrixRowFromBeltRatios(m2.belt_ratios);
            T rp00 = m1.gear_ratio * b1[0];
            T rp10 = m2.gear_ratio * b2[0];
            T rp11 = m2.gear_ratio * b2[1];

            Mat<T> G(4, 2);
            G(link1_index, 0) = T(1.0);
            G(rotor1_index, 0) = rp00;
            G(rotor2_index, 0) = rp10;
            G(rotor2_index, 1) = rp11;
            G(link2_index, 1) = T(1.0);

            Mat<T> K(2, 4);
            int c1 = rotor1_index > rotor2_index;
            int c2 = rotor2_index > rotor1_index;
            K(c1, rotor1_index) = T(-1.0);
            K(c1, link1_index) = G(rotor1_index, 0);
            K(c2, rotor2_index) = T(-1.0);
            K(c2, link1_index) = G(rotor2_index, 0);
            K(c2, link2_index) = G(rotor2_index, 1);
            this->loop_constraint = std::make_shared<StaticConstraint<T>>(G, K);

            X_intra_S_span.setBlock(6 * link1_index, link1_index, link1_joint->S);
            X_intra_S_span.setBlock(6 * rotor1_index, rotor1_index, rotor1_joint->S);
            X_intra_S_span.setBlock(6 * rotor2_index, rotor2_index, rotor2_joint->S);
            X_intra_S_span.setBlock(6 * link2_index, link2_index, link2_joint->S);
            this->S = X_intra_S_span * this->G();
        }
        const char *typeName() const override { return "RevolutePairWithRotor"; }
        void updateKinematics(const JointState<T> &js) override
        {
            JointState<T> s = this->toSpanningTreeState(js);
            const Mat<T> &q = s.position, &qd = s.velocity;
            link1_joint->updateKinematics(q.segment(link1_index, 1));
            rotor1_joint->updateKinematics(q.segment(rotor1_index, 1));
            rotor2_joint->updateKinematics(q.segment(rotor2_index, 1));
            link2_joint->updateKinematics(q.segment(link2_index, 1));

            Transform<T> X21 = link2_joint->XJ * link2.Xtree;
            Mat<T> v2_relative = link2_joint->S * qd[link2_index];
            Mat<T> X21S1 = X21.transformMotionVector(link1_joint->S);
            X_intra_S_span.setBlock(6 * link2_index, link1_index, X21S1);
            this->S.setBlock(6 * link2_index, 0, X21S1);
            X_intra_S_span_ring.setBlock(6 * link2_index, link1_index,
                                         -(motionCrossMatrix(v2_relative) * X21S1));
            this->vJ = X_intra_S_span * qd;
            this->cJ = X_intra_S_span_ring * qd;
        }
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            Xup.X[link1_index] = link1_joint->XJ * link1.Xtree;
            Xup.X[rotor1_index] = rotor1_joint->XJ * rotor1.Xtree;
            Xup.X[rotor2_index] = rotor2_joint->XJ * rotor2.Xtree;
            Xup.X[link2_index] = link2_joint->XJ * link2.Xtree * Xup.X[link1_index];
        }
    };

    // reference: src/Dynamics/ClusterJoints/RevoluteTripleWithRotorJoint.cpp:10-120
    // (sub-indices are fixed by the reference: links 0, 1, 2, rotors 3, 4, 5)
    template <typename T>
    struct RevoluteTripleWithRotorCluster : ClusterJointBase<T>
    {
        Body<T> link1, link2, link3, rotor1, rotor2, rotor3;
        std::shared_ptr<SingleJoint<T>> link1_joint, link2_joint, link3_joint, rotor1_joint, rotor2_joint, rotor3_joint;
        Mat<T> X_intra_S_span, X_intra_S_span_ring;
        Transform<T> X21, X32, X31;

        RevoluteTripleWithRotorCluster(const ParallelBeltTransmissionModule<T> &m1,
                                       const ParallelBeltTransmissionModule<T> &m2,
                                       const ParallelBeltTransmissionModule<T> &m3)
            : ClusterJointBase<T>(6, 3, 3), link1(m1.body), link2(m2.body), link3(m3.body), rotor1(m1.rotor),
              rotor2(m2.rotor), rotor3(m3.rotor), X_intra_S_span(36, 6), X_intra_S_span_ring(36, 6)
        {
            // :20-26
            link1_joint = std::make_shared<SingleRevolute<T>>(m1.joint_axis);
            link2_joint = std::make_shared<SingleRevolute<T>>(m2.joint_axis);
            link3_joint = std::make_shared<SingleRevolute<T>>(m3.joint_axis);
            rotor1_joint = std::make_shared<SingleRevolute<T>>(m1.rotor_axis);
            rotor2_joint = std::make_shared<SingleRevolute<T>>(m2.rotor_axis);
            rotor3_joint = std::make_shared<SingleRevolute<T>>(m3.rotor_axis);
            this->single_joints = {link1_joint, link2_joint, link3_joint, rotor1_joint, rotor2_joint, rotor3_joint};

            // :31-47  G = [1; diag(gear ratios) * belt matrix], K = [-G_bottom, 1]
            const std::vector<T> b1 = beltMatrixRowFromBeltRatios(m1.belt_ratios);
            const std::vector<T> b2 = beltMatrixRowFromBeltRatios(m2.belt_ratios);
            const std::vector<T> b3 = beltMatrixRowFromBeltRatios(m3.belt_ratios);
            const T gear[3] = {m1.gear_ratio, m2.gear_ratio, m3.gear_ratio};
            Mat<T> G(6, 3), K(3, 6);
            for (int i = 0; i < 3; i++)
                G(i, i) = T(1.0);
            G(3, 0) = gear[0] * b1[0];
            G(4, 0) = gear[1] * b2[0];
            G(4, 1) = gear[1] * b2[1];
            G(5, 0) = gear[2] * b3[0];
            G(5, 1) = gear[2] * b3[1];
            G(5, 2) = gear[2] * b3[2];
            for (int i = 0; i < 3; i++)
            {
                for (int j = 0; j < 3; j++)
                    K(i, j) = -G(3 + i, j);
                K(i, 3 + i) = T(1.0);
            }
            this->loop_constraint = std::make_shared<StaticConstraint<T>>(G, K);

            // :49-59
            X_intra_S_span.setBlock(0, 0, link1_joint->S);
            X_intra_S_span.setBlock(6, 1, link2_joint->S);
            X_intra_S_span.setBlock(12, 2, link3_joint->S);
            X_intra_S_span.setBlock(18, 3, rotor1_joint->S);
            X_intra_S_span.setBlock(24, 4, rotor2_joint->S);
            X_intra_S_span.setBlock(30, 5, rotor3_joint->S);
            this->S = X_intra_S_span * this->G();
        }
        const char *typeName() const override { return "RevoluteTripleWithRotor"; }
        // :62-106
        void updateKinematics(const JointState<T> &js) override
        {
            JointState<T> s = this->toSpanningTreeState(js);
            const Mat<T> &q = s.position, &qd = s.velocity;
            link1_joint->updateKinematics(q.segment(0, 1));
            link2_joint->updateKinematics(q.segment(1, 1));
            link3_joint->updateKinematics(q.segment(2, 1));
            rotor1_joint->updateKinematics(q.segment(3, 1));
            rotor2_joint->updateKinematics(q.segment(4, 1));
            rotor3_joint->updateKinematics(q.segment(5, 1));

            X21 = link2_joint->XJ * link2.Xtree;
            X32 = link3_joint->XJ * link3.Xtree;
            X31 = X32 * X21;

            Mat<T> v2_relative1 = link2_joint->S * qd[1];
            Mat<T> X21_S1 = X21.transformMotionVector(link1_joint->S);
            Mat<T> v3_relative1 = X32.transformMotionVector(v2_relative1) + link3_joint->S * qd[2];
            Mat<T> X31_S1 = X31.transformMotionVector(link1_joint->S);
            Mat<T> v3_relative2 = link3_joint->S * qd[2];
            Mat<T> X32_S2 = X32.transformMotionVector(link2_joint->S);

            X_intra_S_span.setBlock(6, 0, X21_S1);
            X_intra_S_span.setBlock(12, 0, X31_S1);
            X_intra_S_span.setBlock(12, 1, X32_S2);
            // S top-left 18 x 3 = X_intra_S_span top-left 18 x 3 (:88-89)
            for (int i = 0; i < 18; i++)
                for (int j = 0; j < 3; j++)
                    this->S(i, j) = X_intra_S_span(i, j);

            X_intra_S_span_ring.setBlock(6, 0, -(motionCrossMatrix(v2_relative1) * X21_S1));
            X_intra_S_span_ring.setBlock(12, 0, -(motionCrossMatrix(v3_relative1) * X31_S1));
            X_intra_S_span_ring.setBlock(12, 1, -(motionCrossMatrix(v3_relative2) * X32_S2));

            this->vJ = X_intra_S_span * qd;
            this->cJ = X_intra_S_span_ring * qd;
        }
        // :108-120
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            Xup.X[0] = link1_joint->XJ * link1.Xtree;
            Xup.X[1] = X21 * Xup.X[0];
            Xup.X[2] = X31 * Xup.X[0];
            Xup.X[3] = rotor1_joint->XJ * rotor1.Xtree;
            Xup.X[4] = rotor2_joint->XJ * rotor2.Xtree;
            Xup.X[5] = rotor3_joint->XJ * rotor3.Xtree;
        }
    };

    // reference: GenericJoint.cpp:243-505 (ClusterJoints::Generic)
    template <typename T>
    struct GenericCluster : ClusterJointBase<T>
    {
        std::vector<Body<T>> bodies;
        std::vector<std::vector<bool>> connectivity;
        Mat<T> S_spanning, X_intra, X_intra_ring;

        GenericCluster(const std::vector<Body<T>> &bodies_,
                       const std::vector<std::shared_ptr<SingleJoint<T>>> &joints,
                       std::shared_ptr<LoopConstraintBase<T>> lc)
            : ClusterJointBase<T>((int)bodies_.size(),
                                  lc->isExplicit() ? lc->numIndependentPos() : lc->numSpanningPos(),
                                  lc->numIndependentVel()),
              bodies(bodies_)
        {
            this->loop_constraint = lc;
            this->single_joints = joints;
            extractConnectivity();
            // S_spanning = blockdiag(S_i)   (:282-284, appendEigenMatrix)
            int rows = 0, cols = 0;
            for (auto &j : joints)
            {
                rows += j->S.r;
                cols += j->S.c;
            }
            S_spanning = Mat<T>(rows, cols);
            int r0 = 0, c0 = 0;
            for (auto &j : joints)
            {
                S_spanning.setBlock(r0, c0, j->S);
                r0 += j->S.r;
                c0 += j->S.c;
            }
            X_intra = Mat<T>::Identity(6 * this->num_bodies);
            X_intra_ring = Mat<T>(6 * this->num_bodies, 6 * this->num_bodies);
        }
        const char *typeName() const override { return "Generic"; }

        // :472-485
        void extractConnectivity()
        {
            const int N = this->num_bodies;
            connectivity.assign(N, std::vector<bool>(N, false));
            for (int i = 0; i < N; i++)
            {
                int j = i;
                while (bodyInCurrentCluster(bodies[j].parent_index))
                {
                    j = getBody(bodies[j].parent_index).sub_index_within_cluster;
                    connectivity[i][j] = true;
                }
            }
        }
        bool bodyInCurrentCluster(int body_index) const
        {
            for (auto &b : bodies)
                if (b.index == body_index)
                    return true;
            return false;
        }
        const Body<T> &getBody(int body_index) const
        {
            for (auto &b : bodies)
                if (b.index == body_index)
                    return b;
            throw std::runtime_error("Body is not in the current cluster");
        }

        // :388-451
        void updateKinematics(const JointState<T> &js) override
        {
            JointState<T> s = this->toSpanningTreeState(js);
            const Mat<T> &q = s.position, &qd = s.velocity;
            const int N = this->num_bodies;

            int pos_idx = 0;
            for (int i = 0; i < N; i++)
            {
                auto joint = this->single_joints[i];
                joint->updateKinematics(q.segment(pos_idx, joint->num_positions));
                int k = i;
                for (int j = i - 1; j >= 0; j--)
                    if (connectivity[i][j])
                    {
                        Mat<T> Xup_prev = X_intra.block(6 * i, 6 * k, 6, 6);
                        Mat<T> Xint = (this->single_joints[k]->XJ * bodies[k].Xtree).toMatrix();
                        X_intra.setBlock(6 * i, 6 * j, Xup_prev * Xint);
                        k = j;
                    }
                pos_idx += joint->num_positions;
            }

            Mat<T> S_implicit = X_intra * S_spanning;
            this->S = S_implicit * this->G();
            this->vJ = S_implicit * qd;

            for (int i = 0; i < N; i++)
                for (int j = i - 1; j >= 0; j--)
                    if (connectivity[i][j])
                    {
                        Mat<T> Xup = X_intra.block(6 * i, 6 * j, 6, 6);
                        Mat<T> v_parent = Xup * this->vJ.segment(6 * j, 6);
                        Mat<T> v_child = this->vJ.segment(6 * i, 6);
                        Mat<T> v_relative = v_child - v_parent;
                        X_intra_ring.setBlock(6 * i, 6 * j, -(motionCrossMatrix(v_relative) * Xup));
                    }

            this->cJ = X_intra_ring * (S_spanning * qd) + S_implicit * this->g();
        }

        // :454-469
        void computeXup(GeneralizedTransform<T> &Xup) const override
        {
            for (int i = 0; i < this->num_bodies; i++)
            {
                Xup.X[i] = this->single_joints[i]->XJ * bodies[i].Xtree;
                for (int j = i - 1; j >= 0; j--)
                    if (connectivity[i][j])
                    {
                        Xup.X[i] = Xup.X[i] * Xup.X[j];
                        break;
                    }
            }
        }
    };

} // namespace grbda_oracle
