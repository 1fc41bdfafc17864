// Small fixed-size host types used while *describing* a model (construction-time constants only;
// no per-state arithmetic happens on the host). They stand in for the Eigen types of the
// reference's public API (reference: include/grbda/Utils/cppTypes.h:17-83).
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace grbda
{
    using Vec3 = std::array<double, 3>;
    using Mat3 = std::array<double, 9>;  // row-major
    using Mat6 = std::array<double, 36>; // row-major

    namespace ori
    {
        // reference: include/grbda/Utils/OrientationTools.h:34-39
        enum class CoordinateAxis
        {
            X = 0,
            Y = 1,
            Z = 2
        };

        inline Mat3 identity3() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }
        inline Mat3 mul(const Mat3 &A, const Mat3 &B)
        {
            Mat3 C{};
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                {
                    double s = 0;
                    for (int k = 0; k < 3; k++)
                        s += A[3 * i + k] * B[3 * k + j];
                    C[3 * i + j] = s;
                }
            return C;
        }
        inline Mat3 transpose(const Mat3 &A)
        {
            return {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
        }
        inline Vec3 mul(const Mat3 &A, const Vec3 &v)
        {
            return {A[0] * v[0] + A[1] * v[1] + A[2] * v[2], A[3] * v[0] + A[4] * v[1] + A[5] * v[2],
                    A[6] * v[0] + A[7] * v[1] + A[8] * v[2]};
        }
        inline Mat3 scale(const Mat3 &A, double s)
        {
            Mat3 B = A;
            for (auto &x : B)
                x *= s;
            return B;
        }
        // reference: OrientationTools.h:46-68 — coordinate rotation (transforms INTO the rotated frame)
        inline Mat3 coordinateRotation(CoordinateAxis axis, double theta)
        {
            const double s = std::sin(theta), c = std::cos(theta);
            if (axis == CoordinateAxis::X)
                return {1, 0, 0, 0, c, s, 0, -s, c};
            if (axis == CoordinateAxis::Y)
                return {c, 0, -s, 0, 1, 0, s, 0, c};
            return {c, s, 0, -s, c, 0, 0, 0, 1};
        }
        // reference: OrientationTools.h:121-130
        inline Mat3 rpyToRotMat(const Vec3 &v)
        {
            return mul(mul(coordinateRotation(CoordinateAxis::X, v[0]),
                           coordinateRotation(CoordinateAxis::Y, v[1])),
                       coordinateRotation(CoordinateAxis::Z, v[2]));
        }
        // reference: OrientationTools.h:251-269, q = (w, x, y, z)
        inline Mat3 quaternionToRotationMatrix(const std::array<double, 4> &q)
        {
            const double e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3];
            Mat3 R = {1 - 2 * (e2 * e2 + e3 * e3), 2 * (e1 * e2 - e0 * e3), 2 * (e1 * e3 + e0 * e2),
                      2 * (e1 * e2 + e0 * e3), 1 - 2 * (e1 * e1 + e3 * e3), 2 * (e2 * e3 - e0 * e1),
                      2 * (e1 * e3 - e0 * e2), 2 * (e2 * e3 + e0 * e1), 1 - 2 * (e1 * e1 + e2 * e2)};
            return transpose(R);
        }
        inline Mat3 skew(const Vec3 &v) { return {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}; }
    } // namespace ori

    namespace spatial
    {
        // reference: include/grbda/Utils/SpatialTransforms.h (spatial::Transform): rotation E and
        // translation r of the child frame expressed in the parent frame.
        struct Transform
        {
            Mat3 E = ori::identity3();
            Vec3 r = {0, 0, 0};
            Transform() {}
            Transform(const Mat3 &E_, const Vec3 &r_) : E(E_), r(r_) {}
            explicit Transform(const Mat3 &E_) : E(E_) {}
        };
    } // namespace spatial

    // reference: include/grbda/Utils/SpatialInertia.h:66-249 (construction helpers only)
    class SpatialInertia
    {
    public:
        SpatialInertia() { I_.fill(0.0); }
        // :74-82  mass, COM, rotational inertia about the COM
        SpatialInertia(double mass, const Vec3 &com, const Mat3 &inertia)
        {
            I_.fill(0.0);
            const Mat3 cS = ori::skew(com);
            const Mat3 ccT = ori::mul(cS, ori::transpose(cS));
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                {
                    I_[6 * i + j] = inertia[3 * i + j] + mass * ccT[3 * i + j];
                    I_[6 * i + 3 + j] = mass * cS[3 * i + j];
                    I_[6 * (3 + i) + j] = mass * cS[3 * j + i];
                }
            for (int i = 0; i < 3; i++)
                I_[6 * (3 + i) + 3 + i] = mass;
        }
        explicit SpatialInertia(const Mat6 &I) : I_(I) {}
        const Mat6 &getMatrix() const { return I_; }
        double getMass() const { return I_[35]; }

        // :233-245 via the pseudo-inertia (:130-142, :212-226)
        SpatialInertia flipAlongAxis(ori::CoordinateAxis axis) const
        {
            // h = m c, Ibar = rotational inertia about the frame origin. Mirroring coordinate `a`
            // negates h_a and the products of inertia Ibar_ab (b != a).
            const int a = (int)axis;
            Mat6 J = I_;
            auto flipSign = [&](int i, int j) { J[6 * i + j] = -J[6 * i + j]; };
            // top-left block: negate off-diagonal entries involving axis a
            for (int b = 0; b < 3; b++)
                if (b != a)
                {
                    flipSign(a, b);
                    flipSign(b, a);
                }
            // top-right = skew(h), bottom-left = skew(h)^T with h_a negated
            Vec3 h = {0.5 * (I_[6 * 2 + 4] - I_[6 * 1 + 5]), 0.5 * (I_[6 * 0 + 5] - I_[6 * 2 + 3]),
                      0.5 * (I_[6 * 1 + 3] - I_[6 * 0 + 4])};
            h[a] = -h[a];
            const Mat3 hS = ori::skew(h);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                {
                    J[6 * i + 3 + j] = hS[3 * i + j];
                    J[6 * (3 + i) + j] = hS[3 * j + i];
                }
            return SpatialInertia(J);
        }

    private:
        Mat6 I_;
    };

} // namespace grbda
