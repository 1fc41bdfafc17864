// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// Deterministic synthetic state generation. Counter-based Philox4x32-10
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) keyed by the seed and
// indexed by the GLOBAL state index, so shards are reproducible regardless of GPU count.
// Ranges follow the reference generators:
//   ClusterJoints::Base::randomJointState   src/Dynamics/ClusterJoints/ClusterJoint.cpp:74-81
//   ClusterJoints::Free::randomJointState   FreeJoint.cpp:49-60 (+ rpyToQuat)
//   ClusterJoints::Generic::randomJointState / findRootsForPhi   GenericJoint.cpp:290-385
// The product's device generator (generalized_rbda_b200/csrc) implements the same stream; tests
// compare the two.
#pragma once
#include "model.h"

namespace grbda_oracle
{
    struct Philox
    {
        uint32_t key[2];
        uint32_t ctr_state[2]; // global state index (lo, hi)
        uint32_t draw = 0;     // number of doubles drawn so far
        uint32_t buf[4];

        Philox(uint64_t seed, uint64_t state_index)
        {
            key[0] = (uint32_t)seed;
            key[1] = (uint32_t)(seed >> 32);
            ctr_state[0] = (uint32_t)state_index;
            ctr_state[1] = (uint32_t)(state_index >> 32);
        }
        static void round(uint32_t c[4], const uint32_t k[2])
        {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
            const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
            uint32_t n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
            uint32_t n3 = (uint32_t)p0;
            c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        }
        void block(uint32_t blk)
        {
            uint32_t c[4] = {ctr_state[0], ctr_state[1], blk, 0x67726264u /* "grbd" */};
            uint32_t k[2] = {key[0], key[1]};
            for (int i = 0; i < 10; i++)
            {
                round(c, k);
                k[0] += 0x9E3779B9u;
                k[1] += 0xBB67AE85u;
            }
            for (int i = 0; i < 4; i++)
                buf[i] = c[i];
        }
        // uniform double in [-1, 1) from 53 random bits
        double uniform()
        {
            if ((draw & 1u) == 0)
                block(draw >> 1);
            const uint32_t lo = buf[2 * (draw & 1u)], hi = buf[2 * (draw & 1u) + 1];
            draw++;
            const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
            return (double)bits * (2.0 / 9007199254740992.0) - 1.0;
        }
    };

    // Newton solve for the dependent coordinates of an implicit cluster.
    // Returns true on success; q (spanning) holds the independent coordinates on entry.
    inline bool solveImplicitPosition(GenericImplicitConstraint<double> &lc, Mat<double> &q)
    {
        for (int it = 0; it < 30; it++)
        {
            Mat<double> phi = lc.phi(q);
            double nrm = 0;
            for (int i = 0; i < phi.r; i++)
                nrm = std::max(nrm, std::fabs(phi[i]));
            if (!(nrm == nrm))
                return false;
            if (nrm < 1e-13)
                return true;
            // K only (updateJacobians also forms G; harmless)
            lc.updateJacobians(q);
            Mat<double> Kd = lc.Kd();
            Mat<double> dq;
            try
            {
                dq = solve(Kd, phi);
            }
            catch (...)
            {
                return false;
            }
            for (size_t i = 0; i < lc.dep_coords.size(); i++)
                q[lc.dep_coords[i]] -= dq[i];
        }
        return lc.phi(q).norm() < 1e-10;
    }

    // One state: q[nq], yd[nv], aux[nv] (aux = ydd for ID, tau for FD).
    inline void generateState(ClusterTreeModel<double> &model, uint64_t seed, uint64_t index,
                              double *q, double *yd, double *aux)
    {
        Philox rng(seed, index);
        for (auto &node : model.nodes)
        {
            double *qc = q + node->position_index;
            auto lc = node->joint->loop_constraint;
            if (dynamic_cast<FreeCluster<double> *>(node->joint.get()) ||
                dynamic_cast<FreeConstraint<double> *>(lc.get()))
            {
                for (int i = 0; i < 3; i++)
                    qc[i] = rng.uniform();
                Mat<double> rpy(3, 1);
                for (int i = 0; i < 3; i++)
                    rpy[i] = rng.uniform();
                if (node->num_positions == 7)
                {
                    Mat<double> quat = rpyToQuat(rpy);
                    for (int i = 0; i < 4; i++)
                        qc[3 + i] = quat[i];
                }
                else
                    for (int i = 0; i < 3; i++)
                        qc[3 + i] = rpy[i];
            }
            else if (lc->isExplicit())
            {
                for (int i = 0; i < node->num_positions; i++)
                    qc[i] = rng.uniform();
            }
            else
            {
                auto *gi = dynamic_cast<GenericImplicitConstraint<double> *>(lc.get());
                if (!gi)
                    throw std::runtime_error("implicit constraint without GenericImplicit");
                Mat<double> qs(node->num_positions, 1);
                bool ok = false;
                for (int attempt = 0; attempt < 45 && !ok; attempt++)
                {
                    for (int i : gi->ind_coords)
                        qs[i] = rng.uniform();
                    for (int i : gi->dep_coords)
                        qs[i] = 0.1 * rng.uniform();
                    ok = solveImplicitPosition(*gi, qs);
                    if (ok)
                    {
                        gi->updateJacobians(qs);
                        ok = cond_estimate(gi->Kd()) < 1e6;
                    }
                }
                if (!ok)
                    throw std::runtime_error("Failed to find valid roots for implicit loop constraint");
                for (int i = 0; i < node->num_positions; i++)
                    qc[i] = qs[i];
            }
        }
        const int nv = model.getNumDegreesOfFreedom();
        for (int i = 0; i < nv; i++)
            yd[i] = rng.uniform();
        for (int i = 0; i < nv; i++)
            aux[i] = rng.uniform();
    }

    // ---- integration step (test infrastructure for the product's grbda_cuda_integrate_f64) -------------------
    // reference: include/grbda/Utils/OrientationTools.h:365-377 (quatProduct)
    inline void quatProduct(const double a[4], const double b[4], double out[4])
    {
        const double r = a[0] * b[0] - (a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
        const double v[3] = {a[0] * b[1] + b[0] * a[1] + (a[2] * b[3] - a[3] * b[2]),
                             a[0] * b[2] + b[0] * a[2] + (a[3] * b[1] - a[1] * b[3]),
                             a[0] * b[3] + b[0] * a[3] + (a[1] * b[2] - a[2] * b[1])};
        out[0] = r, out[1] = v[0], out[2] = v[1], out[3] = v[2];
    }
    // reference: include/grbda/Utils/OrientationTools.h:387-413 (integrateQuat; omega in INERTIAL coordinates)
    inline void integrateQuat(const double quat[4], const double omega[3], double dt, double out[4])
    {
        double axis[3] = {1.0, 0.0, 0.0};
        double ang = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
        if (ang > 0)
            for (int i = 0; i < 3; i++)
                axis[i] = omega[i] / ang;
        ang *= dt;
        const double s = std::sin(ang / 2), quatD[4] = {std::cos(ang / 2), s * axis[0], s * axis[1], s * axis[2]};
        double qn[4];
        quatProduct(quatD, quat, qn);
        const double n = std::sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
        for (int i = 0; i < 4; i++)
            out[i] = qn[i] / n;
    }
    // Semi-implicit Euler step of one state: yd' = yd + dt ydd; positions advance with yd' - revolute
    // coordinates linearly, the free base through its body-frame twist (Joints::Free, Joint.h:61-68: the
    // joint transform is (E = R(quat)^T, p), so world rates are R omega_body and R v_body), implicit clusters
    // on their independent coordinates followed by the Newton projection onto phi = 0 used by
    // GenericJoint.cpp:290-385. Returns false when a projection did not converge.
    inline bool integrateState(ClusterTreeModel<double> &model, const double *q, const double *yd, const double *ydd,
                               double dt, double *q_out, double *yd_out)
    {
        const int nv = model.getNumDegreesOfFreedom();
        for (int i = 0; i < nv; i++)
            yd_out[i] = yd[i] + dt * ydd[i];
        bool ok = true;
        for (auto &node : model.nodes)
        {
            const double *qc = q + node->position_index, *vc = yd_out + node->velocity_index;
            double *qo = q_out + node->position_index;
            auto lc = node->joint->loop_constraint;
            if (dynamic_cast<FreeCluster<double> *>(node->joint.get()) || dynamic_cast<FreeConstraint<double> *>(lc.get()))
            {
                if (node->num_positions != 7)
                {
                    for (int i = 0; i < node->num_positions; i++)
                        qo[i] = qc[i];
                    ok = false;
                    continue;
                }
                Mat<double> quat(4, 1);
                for (int i = 0; i < 4; i++)
                    quat[i] = qc[3 + i];
                const Mat<double> R = quaternionToRotationMatrix(quat).transpose(); // body -> world
                double om[3], v[3];
                for (int i = 0; i < 3; i++)
                {
                    om[i] = R(i, 0) * vc[0] + R(i, 1) * vc[1] + R(i, 2) * vc[2];
                    v[i] = R(i, 0) * vc[3] + R(i, 1) * vc[4] + R(i, 2) * vc[5];
                }
                for (int i = 0; i < 3; i++)
                    qo[i] = qc[i] + dt * v[i];
                integrateQuat(qc + 3, om, dt, qo + 3);
            }
            else if (lc->isExplicit())
            {
                for (int i = 0; i < node->num_positions; i++)
                    qo[i] = qc[i] + dt * vc[i];
            }
            else
            {
                auto *gi = dynamic_cast<GenericImplicitConstraint<double> *>(lc.get());
                Mat<double> qs(node->num_positions, 1);
                for (int i = 0; i < node->num_positions; i++)
                    qs[i] = qc[i];
                for (size_t k = 0; k < gi->ind_coords.size(); k++)
                    qs[gi->ind_coords[k]] += dt * vc[k];
                ok = solveImplicitPosition(*gi, qs) && ok;
                for (int i = 0; i < node->num_positions; i++)
                    qo[i] = qs[i];
            }
        }
        return ok;
    }

} // namespace grbda_oracle
