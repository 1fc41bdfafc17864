// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// ClusterTreeModel restated from the reference:
//   include/grbda/Dynamics/{TreeModel,ClusterTreeModel}.h, Nodes/{TreeNode,ClusterTreeNode}.h
//   src/Dynamics/{TreeModel,ClusterTreeModel,ClusterTreeDynamics}.cpp, Nodes/ClusterTreeNode.cpp
#pragma once
#include <map>
#include "cluster_joints.h"

namespace grbda_oracle
{
    // reference: Nodes/TreeNode.h:16-75 + Nodes/ClusterTreeNode.h:12-55
    template <typename T>
    struct ClusterTreeNode
    {
        int index, parent_index, num_parent_bodies;
        std::string name;
        int position_index, num_positions, velocity_index, num_velocities;
        int motion_subspace_index, motion_subspace_dimension;
        std::vector<Body<T>> bodies;
        std::shared_ptr<ClusterJointBase<T>> joint;
        JointState<T> joint_state;

        Mat<T> I;                     // block diagonal 6N x 6N (ClusterTreeNode.cpp:17-23)
        GeneralizedTransform<T> Xup;  // per body transform from its cluster-ancestor body
        std::vector<Transform<T>> Xa; // absolute transforms (GeneralizedAbsoluteTransform)
        Mat<T> v, a, f, f_ext, avp;
        Mat<T> Ic;
        // ABA scratch (ClusterTreeNode.h:43-50)
        Mat<T> IA, pA, U, D, u, D_inv_UT, D_inv_u, Ia;
    };

    // reference: include/grbda/Dynamics/StateRepresentation.h (ContactPoint)
    template <typename T>
    struct ContactPoint
    {
        int body_index;
        Mat<T> local_offset; // 3 x 1, body frame
        bool is_end_effector;
    };

    template <typename T>
    struct ClusterTreeModel
    {
        std::vector<ContactPoint<T>> contact_points;
        std::vector<Body<T>> bodies;
        std::vector<Body<T>> bodies_in_current_cluster;
        std::vector<std::shared_ptr<ClusterTreeNode<T>>> nodes;
        std::map<std::string, int> body_name_to_body_index;
        std::map<std::string, int> cluster_name_to_cluster_index;
        int position_index = 0, velocity_index = 0, motion_subspace_index = 0;
        Mat<T> gravity; // 6x1, TreeModel.h:19-22
        std::vector<int> nodes_with_external_forces;
        bool all_positions_valid = true;

        ClusterTreeModel() : gravity(6, 1)
        {
            gravity[5] = T(-9.81);
            body_name_to_body_index["ground"] = -1;
        }
        void setGravity(double gx, double gy, double gz)
        {
            gravity[3] = T(gx);
            gravity[4] = T(gy);
            gravity[5] = T(gz);
        }
        int getNumPositions() const { return position_index; }
        int getNumDegreesOfFreedom() const { return velocity_index; }
        int getNumBodies() const { return (int)bodies.size(); }

        ////////////////////////////////////////////////////////////////////////////////////////
        // Construction (ClusterTreeModel.cpp:9-67, :407-528)
        ////////////////////////////////////////////////////////////////////////////////////////
        int clusterContainingBody(int body_index) const
        {
            for (size_t i = 0; i < nodes.size(); i++)
                for (auto &b : nodes[i]->bodies)
                    if (b.index == body_index)
                        return (int)i;
            return -1;
        }

        Body<T> registerBody(const std::string &name, const Mat<T> &inertia,
                             const std::string &parent_name, const Transform<T> &Xtree)
        {
            const int body_index = (int)bodies.size();
            body_name_to_body_index[name] = body_index;
            const int parent_body_index = body_name_to_body_index.at(parent_name);
            // getClusterAncestorIndexFromParent (:407-416)
            int anc = parent_body_index;
            while (anc != -1 && clusterContainingBody(anc) < 0)
                anc = bodies[anc].parent_index;
            const int anc_sub = anc >= 0 ? bodies[anc].sub_index_within_cluster : 0;
            Body<T> body{body_index, name, parent_body_index, Xtree, inertia,
                         (int)bodies_in_current_cluster.size(), anc, anc_sub};
            bodies.push_back(body);
            bodies_in_current_cluster.push_back(body);
            return body;
        }

        void appendRegisteredBodiesAsCluster(const std::string &name,
                                             std::shared_ptr<ClusterJointBase<T>> joint)
        {
            // getIndexOfParentClusterFromBodies (:461-489)
            // body index -1 ("ground") maps to cluster -1 (ClusterTreeModel.h:27-31)
            int parent_cluster_index = -2;
            for (auto &b : bodies_in_current_cluster)
            {
                if (b.parent_index == -1)
                {
                    parent_cluster_index = -1;
                    break;
                }
                int c = clusterContainingBody(b.parent_index);
                if (c >= 0)
                {
                    parent_cluster_index = c;
                    break;
                }
            }
            if (parent_cluster_index == -2)
                throw std::runtime_error("At least one body in every cluster must have a parent "
                                         "in a different clusters");
            const int num_parent_bodies =
                parent_cluster_index >= 0 ? (int)nodes[parent_cluster_index]->bodies.size() : 1;

            auto node = std::make_shared<ClusterTreeNode<T>>();
            node->index = (int)nodes.size();
            node->name = name;
            node->parent_index = parent_cluster_index;
            node->num_parent_bodies = num_parent_bodies;
            node->bodies = bodies_in_current_cluster;
            node->joint = joint;
            node->position_index = position_index;
            node->num_positions = joint->num_positions;
            node->velocity_index = velocity_index;
            node->num_velocities = joint->num_velocities;
            node->motion_subspace_index = motion_subspace_index;
            const int N = (int)node->bodies.size();
            node->motion_subspace_dimension = 6 * N;
            node->I = Mat<T>(6 * N, 6 * N);
            node->Xup.num_parent_bodies = num_parent_bodies;
            for (int i = 0; i < N; i++)
            {
                node->I.setBlock(6 * i, 6 * i, node->bodies[i].inertia);
                node->Xup.X.push_back(Transform<T>());
                node->Xup.parent_sub.push_back(
                    node->bodies[i].cluster_ancestor_sub_index_within_cluster);
                node->Xa.push_back(Transform<T>());
            }
            node->v = node->a = node->f = node->f_ext = node->avp = Mat<T>(6 * N, 1);
            cluster_name_to_cluster_index[name] = node->index;
            nodes.push_back(node);

            // checkValidParentClusterForBodiesInCluster (:112-127)
            for (auto &b : node->bodies)
            {
                int other = b.parent_index >= 0 ? clusterContainingBody(b.parent_index) : -1;
                if (other != node->index && other != parent_cluster_index)
                    throw std::runtime_error("The parents of all bodies in a cluster must have "
                                             "parents in the current cluster OR in the same "
                                             "parent cluster");
            }

            position_index += joint->num_positions;
            velocity_index += joint->num_velocities;
            motion_subspace_index += 6 * N;
            bodies_in_current_cluster.clear();
        }

        ////////////////////////////////////////////////////////////////////////////////////////
        // State (ClusterTreeModel.cpp:256-308). q uses the batched layout of SURVEY Appendix F:
        // spanning coordinates for implicit clusters, independent coordinates otherwise.
        ////////////////////////////////////////////////////////////////////////////////////////
        void setState(const Mat<T> &q, const Mat<T> &yd)
        {
            for (auto &n : nodes)
            {
                n->joint_state.position = q.segment(n->position_index, n->num_positions);
                n->joint_state.position_is_spanning = n->joint->loop_constraint->isImplicit();
                n->joint_state.velocity = yd.segment(n->velocity_index, n->num_velocities);
                n->joint_state.velocity_is_spanning = false;
            }
            clearExternalForces();
        }
        void clearExternalForces()
        {
            for (int i : nodes_with_external_forces)
                nodes[i]->f_ext.setZero();
            nodes_with_external_forces.clear();
        }
        // TreeModel.cpp:215-239 ; force = 6x1 spatial force expressed in world coordinates
        void applyExternalForce(int body_index, const Mat<T> &force)
        {
            int c = clusterContainingBody(body_index);
            nodes[c]->f_ext.addSegment(6 * bodies[body_index].sub_index_within_cluster, force);
            bool found = false;
            for (int i : nodes_with_external_forces)
                found |= (i == c);
            if (!found)
                nodes_with_external_forces.push_back(c);
        }

        ////////////////////////////////////////////////////////////////////////////////////////
        // TreeModel::forwardKinematics (TreeModel.cpp:7-32)
        ////////////////////////////////////////////////////////////////////////////////////////
        void forwardKinematics()
        {
            all_positions_valid = true;
            for (auto &node : nodes)
            {
                // ClusterTreeNode::updateKinematics (ClusterTreeNode.cpp:27-31)
                node->joint->updateKinematics(node->joint_state);
                node->joint->computeXup(node->Xup);
                all_positions_valid = all_positions_valid && node->joint->last_position_valid;

                const int N = (int)node->bodies.size();
                if (node->parent_index >= 0)
                {
                    auto &parent = nodes[node->parent_index];
                    node->v = node->Xup.transformMotionVector(parent->v) + node->joint->vJ;
                    for (int i = 0; i < N; i++)
                        node->Xa[i] = node->Xup.X[i] * parent->Xa[node->Xup.parent_sub[i]];
                }
                else
                {
                    node->v = node->joint->vJ;
                    for (int i = 0; i < N; i++)
                        node->Xa[i] = node->Xup.X[i];
                }
                node->avp = generalMotionCrossProduct(node->v, node->joint->vJ);
            }
        }

        // TreeModel.cpp:35-57
        void forwardAccelerationKinematics(const Mat<T> &qdd)
        {
            forwardKinematics();
            for (auto &node : nodes)
            {
                Mat<T> ydd = qdd.segment(node->velocity_index, node->num_velocities);
                Mat<T> a_parent = node->parent_index >= 0 ? nodes[node->parent_index]->a : -gravity;
                node->a = node->Xup.transformMotionVector(a_parent) + node->joint->S * ydd +
                          node->joint->cJ + node->avp;
            }
        }

        // TreeModel.cpp:174-212 (= ClusterTreeModel::inverseDynamics, ClusterTreeDynamics.cpp:79-83)
        Mat<T> inverseDynamics(const Mat<T> &qdd)
        {
            forwardAccelerationKinematics(qdd);
            Mat<T> tau(qdd.r, 1);
            for (auto &node : nodes)
                node->f = node->I * node->a + generalForceCrossProduct(node->v, node->I * node->v);
            for (int idx : nodes_with_external_forces)
                node_subtract_external(nodes[idx]->f, *nodes[idx]);
            for (int i = (int)nodes.size() - 1; i >= 0; i--)
            {
                auto &node = nodes[i];
                tau.setSegment(node->velocity_index, node->joint->S.transpose() * node->f);
                if (node->parent_index >= 0)
                {
                    auto &parent = nodes[node->parent_index];
                    parent->f = parent->f + node->Xup.inverseTransformForceVector(node->f);
                }
            }
            return tau;
        }
        // ClusterTreeModel.cpp:105-110
        Mat<T> getBiasForceVector() { return inverseDynamics(Mat<T>(getNumDegreesOfFreedom(), 1)); }

        // Xa.transformExternalForceVector (SpatialTransforms.cpp:235-250)
        void node_subtract_external(Mat<T> &target, const ClusterTreeNode<T> &node)
        {
            for (size_t b = 0; b < node.bodies.size(); b++)
                target.addSegment(6 * b, -node.Xa[b].transformForceVector(node.f_ext.segment(6 * b, 6)));
        }

        // ClusterTreeDynamics.cpp:157-191
        void updateArticulatedBodies()
        {
            for (auto &c : nodes)
                c->IA = c->I;
            for (int i = (int)nodes.size() - 1; i >= 0; i--)
            {
                auto &c = nodes[i];
                const Mat<T> &S = c->joint->S;
                c->U = c->IA * S;
                c->D = S.transpose() * c->U;
                c->D_inv_UT = solve(c->D, c->U.transpose()); // updateDinv + solve
                if (c->parent_index >= 0)
                {
                    auto &p = nodes[c->parent_index];
                    c->Ia = c->IA - c->U * c->D_inv_UT;
                    p->IA = p->IA + c->Xup.inverseTransformSpatialInertia(c->Ia);
                }
            }
        }

        // ClusterTreeDynamics.cpp:85-155
        Mat<T> forwardDynamics(const Mat<T> &tau)
        {
            Mat<T> qdd(getNumDegreesOfFreedom(), 1);
            forwardKinematics();
            updateArticulatedBodies();
            for (auto &c : nodes)
                c->pA = generalForceCrossProduct(c->v, c->I * c->v);
            for (int idx : nodes_with_external_forces)
                node_subtract_external(nodes[idx]->pA, *nodes[idx]);
            for (int i = (int)nodes.size() - 1; i >= 0; i--)
            {
                auto &c = nodes[i];
                const Mat<T> &S = c->joint->S;
                c->u = tau.segment(c->velocity_index, c->num_velocities) - S.transpose() * c->pA;
                c->D_inv_u = solve(c->D, c->u);
                if (c->parent_index >= 0)
                {
                    auto &p = nodes[c->parent_index];
                    Mat<T> pa = c->pA + c->Ia * (c->joint->cJ + c->avp) + c->U * c->D_inv_u;
                    p->pA = p->pA + c->Xup.inverseTransformForceVector(pa);
                }
            }
            for (auto &c : nodes)
            {
                Mat<T> a_parent = c->parent_index >= 0 ? nodes[c->parent_index]->a : -gravity;
                Mat<T> a_temp = c->Xup.transformMotionVector(a_parent) + c->joint->cJ + c->avp;
                Mat<T> ydd = c->D_inv_u - c->D_inv_UT * a_temp;
                qdd.setSegment(c->velocity_index, ydd);
                c->a = a_temp + c->joint->S * ydd;
            }
            return qdd;
        }

        // TreeModel.cpp:116-160 (= getMassMatrix, ClusterTreeModel.cpp:98-103)
        Mat<T> getMassMatrix()
        {
            const int nv = getNumDegreesOfFreedom();
            Mat<T> H(nv, nv);
            forwardKinematics();
            for (auto &n : nodes)
                n->Ic = n->I;
            for (int i = (int)nodes.size() - 1; i >= 0; i--)
            {
                auto &ni = nodes[i];
                if (ni->parent_index >= 0)
                {
                    auto &p = nodes[ni->parent_index];
                    p->Ic = p->Ic + ni->Xup.inverseTransformSpatialInertia(ni->Ic);
                }
                Mat<T> F = ni->Ic * ni->joint->S;
                H.setBlock(ni->velocity_index, ni->velocity_index, ni->joint->S.transpose() * F);
                int j = i;
                while (nodes[j]->parent_index > -1)
                {
                    F = nodes[j]->Xup.inverseTransformForceSubspace(F);
                    j = nodes[j]->parent_index;
                    Mat<T> Hij = F.transpose() * nodes[j]->joint->S;
                    H.setBlock(ni->velocity_index, nodes[j]->velocity_index, Hij);
                    H.setBlock(nodes[j]->velocity_index, ni->velocity_index, Hij.transpose());
                }
            }
            return H;
        }

        ////////////////////////////////////////////////////////////////////////////////////////
        // Operational space: contact points (ClusterTreeModel.cpp:129-190 appendContactPoint / appendEndEffector)
        ////////////////////////////////////////////////////////////////////////////////////////
        void appendContactPoint(int body_index, const Mat<T> &local_offset, bool is_end_effector)
        {
            contact_points.push_back(ContactPoint<T>{body_index, local_offset, is_end_effector});
        }
        int getNumEndEffectors() const
        {
            int n = 0;
            for (auto &c : contact_points)
                n += c.is_end_effector;
            return n;
        }
        // reference: Spatial.h:238-250 (createSXform)
        static Mat<T> createSXform(const Mat<T> &R, const Mat<T> &r)
        {
            Mat<T> X(6, 6), rhat(3, 3);
            rhat(0, 1) = -r[2], rhat(0, 2) = r[1], rhat(1, 0) = r[2], rhat(1, 2) = -r[0], rhat(2, 0) = -r[1], rhat(2, 1) = r[0];
            X.setBlock(0, 0, R);
            X.setBlock(3, 3, R);
            X.setBlock(3, 0, -(R * rhat));
            return X;
        }
        // TreeModel::contactPointForwardKinematics, TreeModel.cpp:60-78 (forwardKinematics() done by the caller)
        void contactPointKinematics(int cp_index, T *pos, T *vel)
        {
            const ContactPoint<T> &cp = contact_points[cp_index];
            const Body<T> &b = bodies[cp.body_index];
            auto &node = nodes[clusterContainingBody(cp.body_index)];
            const Transform<T> &Xa = node->Xa[b.sub_index_within_cluster];
            const Mat<T> v_body = node->v.segment(6 * b.sub_index_within_cluster, 6);
            const Mat<T> position = Xa.inverseTransformPoint(cp.local_offset);
            const Mat<T> vw = Xa.inverseTransformMotionVector(v_body);
            // spatialToLinearVelocity(v, x) = v_lin + w x x   (Spatial.h:296-307)
            const Mat<T> w = vw.segment(0, 3), vl = vw.segment(3, 3);
            const T lin[3] = {vl[0] + w[1] * position[2] - w[2] * position[1], vl[1] + w[2] * position[0] - w[0] * position[2],
                              vl[2] + w[0] * position[1] - w[1] * position[0]};
            for (int i = 0; i < 3; i++)
            {
                pos[i] = position[i];
                vel[i] = lin[i];
            }
        }
        // contactJacobianWorldFrame, ClusterTreeDynamics.cpp:10-45 (world = true) and contactJacobianBodyFrame,
        // :47-77 (world = false): 6 x nv
        Mat<T> contactJacobian(int cp_index, bool world)
        {
            const ContactPoint<T> &cp = contact_points[cp_index];
            Mat<T> J(6, getNumDegreesOfFreedom());
            const Body<T> &body_i = bodies[cp.body_index];
            auto &cluster_i = nodes[clusterContainingBody(cp.body_index)];
            Mat<T> R(3, 3);
            for (int k = 0; k < 3; k++)
                R(k, k) = T(1.0);
            if (world)
                R = cluster_i->Xa[body_i.sub_index_within_cluster].E.transpose(); // R_link_to_world
            Mat<T> Xout = createSXform(R, cp.local_offset);
            int j = cp.body_index;
            while (j > -1)
            {
                const Body<T> &body_j = bodies[j];
                auto &cluster_j = nodes[clusterContainingBody(j)];
                const int sub = body_j.sub_index_within_cluster;
                const Mat<T> S = cluster_j->joint->S.block(6 * sub, 0, 6, cluster_j->num_velocities);
                J.setBlock(0, cluster_j->velocity_index, Xout * S);
                Xout = Xout * cluster_j->Xup.X[sub].toMatrix();
                j = body_j.cluster_ancestor_index;
            }
            return J;
        }
        // What the reference's own tests hold applyTestForce and the EFPA to
        // (UnitTests/testRigidBodyDynamicsAlgos.cpp:241-335): with H = getMassMatrix(),
        //   dstate = H^-1 J_lin^T f,  lambda_inv = f^T J_lin H^-1 J_lin^T f        (world-frame J, linear rows)
        //   Lambda^-1 = J H^-1 J^T over the end-effectors                          (body-frame 6-row J)
        void applyTestForceReference(int cp_index, const T *force, T *dstate, T *lambda_inv)
        {
            const int nv = getNumDegreesOfFreedom();
            const Mat<T> H = getMassMatrix();
            const Mat<T> J = contactJacobian(cp_index, true);
            Mat<T> rhs(nv, 1);
            for (int k = 0; k < nv; k++)
                rhs[k] = J(3, k) * force[0] + J(4, k) * force[1] + J(5, k) * force[2];
            const Mat<T> x = solve(H, rhs);
            T l = T(0.0);
            for (int k = 0; k < nv; k++)
            {
                dstate[k] = x[k];
                l = l + rhs[k] * x[k];
            }
            *lambda_inv = l;
        }
        Mat<T> inverseOperationalSpaceInertiaMatrixReference()
        {
            const int nv = getNumDegreesOfFreedom(), ne = getNumEndEffectors();
            const Mat<T> H = getMassMatrix();
            Mat<T> J(6 * ne, nv);
            int k = 0;
            for (size_t i = 0; i < contact_points.size(); i++)
                if (contact_points[i].is_end_effector)
                    J.setBlock(6 * k++, 0, contactJacobian((int)i, false));
            return J * solve(H, J.transpose());
        }

        ////////////////////////////////////////////////////////////////////////////////////////
        // Kinematic getters (ClusterTreeModel.cpp:319-373), offset = 0.
        // Batched FK output per body: p (world position of the body origin), R (row-major 3x3,
        // getOrientation = Xa.E^T), w (world angular velocity), vl (world linear velocity of origin)
        ////////////////////////////////////////////////////////////////////////////////////////
        void bodyKinematics(int body_index, T *p, T *R, T *vel)
        {
            const Body<T> &b = bodies[body_index];
            auto &node = nodes[clusterContainingBody(body_index)];
            const Transform<T> &Xa = node->Xa[b.sub_index_within_cluster];
            Mat<T> Rai = Xa.E.transpose();
            // getPosition: sXFormPoint(invertSXform(Xa), 0) = Xa.r
            for (int i = 0; i < 3; i++)
                p[i] = Xa.r[i];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    R[3 * i + j] = Rai(i, j);
            Mat<T> vb = node->v.segment(6 * b.sub_index_within_cluster, 6);
            Mat<T> w = Rai * vb.segment(0, 3);
            Mat<T> vl = Rai * vb.segment(3, 3); // spatialToLinearVelocity(v, 0) = v_lin
            for (int i = 0; i < 3; i++)
            {
                vel[i] = w[i];
                vel[3 + i] = vl[i];
            }
        }
    };

} // namespace grbda_oracle
