// Device-side synthetic state generation (one state per thread), hand-written.
// Counter-based Philox4x32-10 keyed by the seed and indexed by the GLOBAL state index, so every
// GPU generates its own shard and the union is independent of the GPU count.
// Value ranges follow the reference's generators:
//   ClusterJoints::Base::randomJointState        src/Dynamics/ClusterJoints/ClusterJoint.cpp:74-81
//   ClusterJoints::Free::randomJointState        FreeJoint.cpp:49-60 (rpy -> quaternion,
//                                                include/grbda/Utils/OrientationTools.h:160-199,293-300)
//   ClusterJoints::Generic::randomJointState     GenericJoint.cpp:290-385 (independent coordinates
//                                                uniform, dependent ones by Newton on phi from a
//                                                guess in [-0.1, 0.1], up to 45 restarts)
// The per-cluster constraint function (phi, K_d) is generated code (struct C below).
#ifndef GRBDA_KERNELS_STATEGEN_CUH // (NVRTC sees this header under two include names: #pragma once is not enough)
#define GRBDA_KERNELS_STATEGEN_CUH
#include "batched_kernel.cuh" // integer typedefs that also work under NVRTC

namespace grbda_kernels
{
    struct Philox
    {
        uint32_t key0, key1, idx0, idx1, draw;
        uint32_t buf[4];

        __device__ Philox(uint64_t seed, uint64_t state_index)
            : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)), idx0((uint32_t)state_index),
              idx1((uint32_t)(state_index >> 32)), draw(0) {}

        __device__ void block(uint32_t blk)
        {
            uint32_t c0 = idx0, c1 = idx1, c2 = blk, c3 = 0x67726264u;
            uint32_t k0 = key0, k1 = key1;
#pragma unroll
            for (int i = 0; i < 10; i++)
            {
                const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
                const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
                const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
                c0 = n0;
                c1 = lo1;
                c2 = n2;
                c3 = lo0;
                k0 += 0x9E3779B9u;
                k1 += 0xBB67AE85u;
            }
            buf[0] = c0;
            buf[1] = c1;
            buf[2] = c2;
            buf[3] = c3;
        }
        // uniform double in [-1, 1) from 53 random bits
        __device__ double uniform()
        {
            if ((draw & 1u) == 0)
                block(draw >> 1);
            const uint32_t lo = buf[2 * (draw & 1u)], hi = buf[2 * (draw & 1u) + 1];
            draw++;
            const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
            return (double)bits * (2.0 / 9007199254740992.0) - 1.0;
        }
    };

    // rpy -> quaternion (w, x, y, z) through the coordinate rotation matrix, exactly as
    // ori::rpyToQuat = rotationMatrixToQuaternion(rpyToRotMat(rpy)).
    __device__ inline void rpyToQuat(const double rpy[3], double q[4])
    {
        double sr, cr, sp, cp, sy, cy;
        sincos(rpy[0], &sr, &cr);
        sincos(rpy[1], &sp, &cp);
        sincos(rpy[2], &sy, &cy);
        // R = Rx(r) Ry(p) Rz(y) with coordinate rotations
        const double Rx[9] = {1, 0, 0, 0, cr, sr, 0, -sr, cr};
        const double Ry[9] = {cp, 0, -sp, 0, 1, 0, sp, 0, cp};
        const double Rz[9] = {cy, sy, 0, -sy, cy, 0, 0, 0, 1};
        double A[9], R[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
            {
                double s = 0;
                for (int k = 0; k < 3; k++)
                    s += Rx[3 * i + k] * Ry[3 * k + j];
                A[3 * i + j] = s;
            }
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
            {
                double s = 0;
                for (int k = 0; k < 3; k++)
                    s += A[3 * i + k] * Rz[3 * k + j];
                R[3 * i + j] = s;
            }
        // r = R^T
        auto r = [&](int i, int j) { return R[3 * j + i]; };
        const double tr = r(0, 0) + r(1, 1) + r(2, 2);
        if (tr > 0.0)
        {
            const double S = sqrt(tr + 1.0) * 2.0;
            q[0] = 0.25 * S;
            q[1] = (r(2, 1) - r(1, 2)) / S;
            q[2] = (r(0, 2) - r(2, 0)) / S;
            q[3] = (r(1, 0) - r(0, 1)) / S;
        }
        else if ((r(0, 0) > r(1, 1)) && (r(0, 0) > r(2, 2)))
        {
            const double S = sqrt(1.0 + r(0, 0) - r(1, 1) - r(2, 2)) * 2.0;
            q[0] = (r(2, 1) - r(1, 2)) / S;
            q[1] = 0.25 * S;
            q[2] = (r(0, 1) + r(1, 0)) / S;
            q[3] = (r(0, 2) + r(2, 0)) / S;
        }
        else if (r(1, 1) > r(2, 2))
        {
            const double S = sqrt(1.0 + r(1, 1) - r(0, 0) - r(2, 2)) * 2.0;
            q[0] = (r(0, 2) - r(2, 0)) / S;
            q[1] = (r(0, 1) + r(1, 0)) / S;
            q[2] = 0.25 * S;
            q[3] = (r(1, 2) + r(2, 1)) / S;
        }
        else
        {
            const double S = sqrt(1.0 + r(2, 2) - r(0, 0) - r(1, 1)) * 2.0;
            q[0] = (r(1, 0) - r(0, 1)) / S;
            q[1] = (r(0, 2) + r(2, 0)) / S;
            q[2] = (r(1, 2) + r(2, 1)) / S;
            q[3] = 0.25 * S;
        }
    }

    // Solve the NC x NC system A x = b in place (partial pivoting); returns false when singular.
    template <int NC>
    __device__ inline bool smallSolve(double *A, double *b)
    {
        for (int k = 0; k < NC; k++)
        {
            int p = k;
            double best = fabs(A[k * NC + k]);
            for (int i = k + 1; i < NC; i++)
                if (fabs(A[i * NC + k]) > best)
                {
                    best = fabs(A[i * NC + k]);
                    p = i;
                }
            if (best == 0.0 || !(best == best))
                return false;
            if (p != k)
            {
                for (int j = 0; j < NC; j++)
                {
                    const double t = A[k * NC + j];
                    A[k * NC + j] = A[p * NC + j];
                    A[p * NC + j] = t;
                }
                const double t = b[k];
                b[k] = b[p];
                b[p] = t;
            }
            for (int i = k + 1; i < NC; i++)
            {
                const double f = A[i * NC + k] / A[k * NC + k];
                for (int j = k + 1; j < NC; j++)
                    A[i * NC + j] -= f * A[k * NC + j];
                b[i] -= f * b[k];
            }
        }
        for (int i = NC - 1; i >= 0; i--)
        {
            double s = b[i];
            for (int k = i + 1; k < NC; k++)
                s -= A[i * NC + k] * b[k];
            b[i] = s / A[i * NC + i];
        }
        return true;
    }

    // Frobenius-norm condition estimate |A| |A^-1| (same estimate as the oracle's generator).
    template <int NC>
    __device__ inline double condEstimate(const double *A)
    {
        double nA = 0, nI = 0;
        for (int i = 0; i < NC * NC; i++)
            nA += A[i] * A[i];
        for (int c = 0; c < NC; c++)
        {
            double M[NC * NC], e[NC];
            for (int i = 0; i < NC * NC; i++)
                M[i] = A[i];
            for (int i = 0; i < NC; i++)
                e[i] = i == c ? 1.0 : 0.0;
            if (!smallSolve<NC>(M, e))
                return 1e300;
            for (int i = 0; i < NC; i++)
                nI += e[i] * e[i];
        }
        return sqrt(nA) * sqrt(nI);
    }

    // Random valid spanning position of one implicit cluster. C provides
    //   N, NC, IND[], DEP[] and  static __device__ void eval(const double *q, double *phi, double *Kd)
    template <typename C>
    __device__ inline bool randomImplicitPosition(Philox &rng, double *q)
    {
        constexpr int N = C::N, NC = C::NC;
        double phi[NC], Kd[NC * NC];
        for (int attempt = 0; attempt < 45; attempt++)
        {
            for (int i = 0; i < N - NC; i++)
                q[C::ind(i)] = rng.uniform();
            for (int i = 0; i < NC; i++)
                q[C::dep(i)] = 0.1 * rng.uniform();
            bool ok = false, failed = false;
            for (int it = 0; it < 30 && !ok && !failed; it++)
            {
                C::eval(q, phi, Kd);
                double nrm = 0;
                for (int i = 0; i < NC; i++)
                    nrm = fmax(nrm, fabs(phi[i]));
                if (!(nrm == nrm))
                    failed = true;
                else if (nrm < 1e-13)
                    ok = true;
                else if (!smallSolve<NC>(Kd, phi))
                    failed = true;
                else
                    for (int i = 0; i < NC; i++)
                        q[C::dep(i)] -= phi[i];
            }
            if (!ok && !failed)
            {
                C::eval(q, phi, Kd);
                double s = 0;
                for (int i = 0; i < NC; i++)
                    s += phi[i] * phi[i];
                ok = sqrt(s) < 1e-10;
            }
            if (ok)
            {
                C::eval(q, phi, Kd);
                ok = condEstimate<NC>(Kd) < 1e6;
            }
            if (ok)
                return true;
        }
        return false;
    }

    // ---- integration step ------------------------------------------------------------------
    // quaternion (w, x, y, z) after rotating for dt with the WORLD-frame angular velocity omega:
    // ori::integrateQuat, include/grbda/Utils/OrientationTools.h:387-413 (axis-angle increment, product, renormalise)
    __device__ inline void integrateQuat(const double q[4], const double omega[3], double dt, double out[4])
    {
        double ang = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
        double axis[3] = {1.0, 0.0, 0.0};
        if (ang > 0.0)
            for (int i = 0; i < 3; i++)
                axis[i] = omega[i] / ang;
        ang *= dt;
        double s, c;
        sincos(0.5 * ang, &s, &c);
        const double d[4] = {c, s * axis[0], s * axis[1], s * axis[2]};
        // quatProduct(d, q) (OrientationTools.h:343-360)
        double r[4] = {d[0] * q[0] - d[1] * q[1] - d[2] * q[2] - d[3] * q[3],
                       d[0] * q[1] + d[1] * q[0] + d[2] * q[3] - d[3] * q[2],
                       d[0] * q[2] - d[1] * q[3] + d[2] * q[0] + d[3] * q[1],
                       d[0] * q[3] + d[1] * q[2] - d[2] * q[1] + d[3] * q[0]};
        const double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        for (int i = 0; i < 4; i++)
            out[i] = r[i] / n;
    }
    // Free joint (Joint.h:61-68, FreeJoint.cpp:29-46): q = [p world; quat], yd = [omega_body; v_body]. With
    // E = R(quat)^T (world -> body, OrientationTools.h:251-269): p' = p + dt E^T v_body, quat' =
    // integrateQuat(quat, E^T omega_body, dt).
    __device__ inline void integrateFreeQuaternion(const double *q, const double *yd, double dt, double *q_out)
    {
        const double w = q[3], x = q[4], y = q[5], z = q[6];
        // body -> world rotation R(quat)
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                             2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                             2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
        double om[3], v[3];
        for (int i = 0; i < 3; i++)
        {
            om[i] = R[3 * i] * yd[0] + R[3 * i + 1] * yd[1] + R[3 * i + 2] * yd[2];
            v[i] = R[3 * i] * yd[3] + R[3 * i + 1] * yd[4] + R[3 * i + 2] * yd[5];
        }
        for (int i = 0; i < 3; i++)
            q_out[i] = q[i] + dt * v[i];
        integrateQuat(q + 3, om, dt, q_out + 3);
    }
    // Dependent coordinates of an implicit cluster back onto phi(q) = 0 (Newton on the dependent coordinates,
    // independent ones held; the same iteration the state generator uses). Returns false when it did not converge.
    template <typename C>
    __device__ inline bool projectImplicitPosition(double *q)
    {
        constexpr int NC = C::NC;
        double phi[NC], Kd[NC * NC];
        for (int it = 0; it < 30; it++)
        {
            C::eval(q, phi, Kd);
            double nrm = 0;
            for (int i = 0; i < NC; i++)
                nrm = fmax(nrm, fabs(phi[i]));
            if (!(nrm == nrm))
                return false;
            if (nrm < 1e-13)
                return true;
            if (!smallSolve<NC>(Kd, phi))
                return false;
            for (int i = 0; i < NC; i++)
                q[C::dep(i)] -= phi[i];
        }
        C::eval(q, phi, Kd);
        double s = 0;
        for (int i = 0; i < NC; i++)
            s += phi[i] * phi[i];
        return sqrt(s) < 1e-10;
    }

    // Step::run advances one state (generated per model): semi-implicit Euler, yd' = yd + dt ydd, q' from yd'
    template <typename Step>
    __global__ void __launch_bounds__(128)
        grbda_integrate_kernel(const double *__restrict__ q, const double *__restrict__ yd, const double *__restrict__ ydd,
                               double dt, int64_t count, double *__restrict__ q_out, double *__restrict__ yd_out,
                               int32_t *__restrict__ flags)
    {
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= count)
            return;
        double ql[Step::NQ], ydl[Step::NV], qn[Step::NQ];
        for (int k = 0; k < Step::NQ; k++)
            ql[k] = q[i * Step::NQ + k];
        for (int k = 0; k < Step::NV; k++)
            ydl[k] = yd[i * Step::NV + k] + dt * ydd[i * Step::NV + k];
        const bool ok = Step::run(ql, ydl, dt, qn);
        for (int k = 0; k < Step::NQ; k++)
            q_out[i * Step::NQ + k] = qn[k];
        for (int k = 0; k < Step::NV; k++)
            yd_out[i * Step::NV + k] = ydl[k];
        if (flags)
            flags[i] = ok ? 0 : 1;
    }

    // Gen::run fills one state (generated per model, csrc/generated/<model>_gen.cu)
    template <typename Gen>
    __global__ void __launch_bounds__(128)
        grbda_generate_kernel(uint64_t seed, int64_t first_index, int64_t count, double *__restrict__ q,
                              double *__restrict__ yd, double *__restrict__ aux, int32_t *__restrict__ flags)
    {
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= count)
            return;
        Philox rng(seed, (uint64_t)(first_index + i));
        double ql[Gen::NQ], ydl[Gen::NV], auxl[Gen::NV];
        const bool ok = Gen::run(rng, ql, ydl, auxl);
        for (int k = 0; k < Gen::NQ; k++)
            q[i * Gen::NQ + k] = ql[k];
        for (int k = 0; k < Gen::NV; k++)
        {
            yd[i * Gen::NV + k] = ydl[k];
            aux[i * Gen::NV + k] = auxl[k];
        }
        if (flags)
            flags[i] = ok ? 0 : 1;
    }

} // namespace grbda_kernels
#endif // GRBDA_KERNELS_STATEGEN_CUH
