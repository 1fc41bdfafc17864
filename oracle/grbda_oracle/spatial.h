// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// Spatial algebra restated from the reference:
//   include/grbda/Utils/OrientationTools.h, Spatial.h, SpatialInertia.h,
//   src/Utils/SpatialTransforms.cpp
// Featherstone convention, angular part first: motion [w; v], force [n; f].
#pragma once
#include "linalg.h"

namespace grbda_oracle
{
    enum class Axis
    {
        X = 0,
        Y = 1,
        Z = 2
    };

    // reference: OrientationTools.h:46-68 (coordinateRotation) — the *coordinate* rotation
    template <typename T>
    Mat<T> coordinateRotation(Axis axis, const T &theta)
    {
        T s = sin(theta);
        T c = cos(theta);
        Mat<T> R = Mat<T>::Identity(3);
        if (axis == Axis::X)
        {
            R(1, 1) = c; R(1, 2) = s;
            R(2, 1) = -s; R(2, 2) = c;
        }
        else if (axis == Axis::Y)
        {
            R(0, 0) = c; R(0, 2) = -s;
            R(2, 0) = s; R(2, 2) = c;
        }
        else
        {
            R(0, 0) = c; R(0, 1) = s;
            R(1, 0) = -s; R(1, 1) = c;
        }
        return R;
    }

    // reference: OrientationTools.h:121-130 (rpyToRotMat)
    template <typename T>
    Mat<T> rpyToRotMat(const Mat<T> &v)
    {
        return coordinateRotation(Axis::X, v[0]) * coordinateRotation(Axis::Y, v[1]) *
               coordinateRotation(Axis::Z, v[2]);
    }

    // reference: OrientationTools.h:251-269 (quaternionToRotationMatrix), q = (w, x, y, z);
    // returns the coordinate transformation into the frame with that orientation (R^T).
    template <typename T>
    Mat<T> quaternionToRotationMatrix(const Mat<T> &q)
    {
        T e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3];
        Mat<T> R(3, 3);
        R(0, 0) = T(1.0) - T(2.0) * (e2 * e2 + e3 * e3);
        R(0, 1) = T(2.0) * (e1 * e2 - e0 * e3);
        R(0, 2) = T(2.0) * (e1 * e3 + e0 * e2);
        R(1, 0) = T(2.0) * (e1 * e2 + e0 * e3);
        R(1, 1) = T(1.0) - T(2.0) * (e1 * e1 + e3 * e3);
        R(1, 2) = T(2.0) * (e2 * e3 - e0 * e1);
        R(2, 0) = T(2.0) * (e1 * e3 - e0 * e2);
        R(2, 1) = T(2.0) * (e2 * e3 + e0 * e1);
        R(2, 2) = T(1.0) - T(2.0) * (e1 * e1 + e2 * e2);
        return R.transpose();
    }

    // reference: OrientationTools.h:160-199 (rotationMatrixToQuaternion)
    inline Mat<double> rotationMatrixToQuaternion(const Mat<double> &r1)
    {
        Mat<double> q(4, 1);
        Mat<double> r = r1.transpose();
        double tr = r(0, 0) + r(1, 1) + r(2, 2);
        if (tr > 0.0)
        {
            double S = std::sqrt(tr + 1.0) * 2.0;
            q[0] = 0.25 * S;
            q[1] = (r(2, 1) - r(1, 2)) / S;
            q[2] = (r(0, 2) - r(2, 0)) / S;
            q[3] = (r(1, 0) - r(0, 1)) / S;
        }
        else if ((r(0, 0) > r(1, 1)) && (r(0, 0) > r(2, 2)))
        {
            double S = std::sqrt(1.0 + r(0, 0) - r(1, 1) - r(2, 2)) * 2.0;
            q[0] = (r(2, 1) - r(1, 2)) / S;
            q[1] = 0.25 * S;
            q[2] = (r(0, 1) + r(1, 0)) / S;
            q[3] = (r(0, 2) + r(2, 0)) / S;
        }
        else if (r(1, 1) > r(2, 2))
        {
            double S = std::sqrt(1.0 + r(1, 1) - r(0, 0) - r(2, 2)) * 2.0;
            q[0] = (r(0, 2) - r(2, 0)) / S;
            q[1] = (r(0, 1) + r(1, 0)) / S;
            q[2] = 0.25 * S;
            q[3] = (r(1, 2) + r(2, 1)) / S;
        }
        else
        {
            double S = std::sqrt(1.0 + r(2, 2) - r(0, 0) - r(1, 1)) * 2.0;
            q[0] = (r(1, 0) - r(0, 1)) / S;
            q[1] = (r(0, 2) + r(2, 0)) / S;
            q[2] = (r(1, 2) + r(2, 1)) / S;
            q[3] = 0.25 * S;
        }
        return q;
    }

    // reference: OrientationTools.h:293-300 (rpyToQuat)
    inline Mat<double> rpyToQuat(const Mat<double> &rpy)
    {
        return rotationMatrixToQuaternion(rpyToRotMat(rpy));
    }

    // reference: src/Utils/SpatialTransforms.cpp:14-197 (spatial::Transform)
    template <typename T>
    struct Transform
    {
        Mat<T> E; // 3x3 rotation
        Mat<T> r; // 3x1 translation

        Transform() : E(Mat<T>::Identity(3)), r(3, 1) {}
        Transform(const Mat<T> &E_, const Mat<T> &r_) : E(E_), r(r_) {}
        explicit Transform(const Mat<T> &E_) : E(E_), r(3, 1) {}

        // :33-40
        Mat<T> toMatrix() const
        {
            Mat<T> X(6, 6);
            X.setBlock(0, 0, E);
            X.setBlock(3, 3, E);
            X.setBlock(3, 0, -(E * skew(r)));
            return X;
        }
        // :43-50   [E w; E (v - r x w)]
        Mat<T> transformMotionVector(const Mat<T> &m) const
        {
            Mat<T> w = m.segment(0, 3), v = m.segment(3, 3);
            Mat<T> out(6, 1);
            out.setSegment(0, E * w);
            out.setSegment(3, E * (v - skew(r) * w));
            return out;
        }
        // :53-61
        Mat<T> inverseTransformMotionVector(const Mat<T> &m) const
        {
            Mat<T> ET = E.transpose();
            Mat<T> w = ET * m.segment(0, 3);
            Mat<T> out(6, 1);
            out.setSegment(0, w);
            out.setSegment(3, skew(r) * w + ET * m.segment(3, 3));
            return out;
        }
        // :64-71
        Mat<T> transformForceVector(const Mat<T> &f) const
        {
            Mat<T> out(6, 1);
            out.setSegment(0, E * (f.segment(0, 3) - skew(r) * f.segment(3, 3)));
            out.setSegment(3, E * f.segment(3, 3));
            return out;
        }
        // :74-82   [E^T n + r x (E^T f); E^T f]
        Mat<T> inverseTransformForceVector(const Mat<T> &f) const
        {
            Mat<T> ET = E.transpose();
            Mat<T> lin = ET * f.segment(3, 3);
            Mat<T> out(6, 1);
            out.setSegment(0, ET * f.segment(0, 3) + skew(r) * lin);
            out.setSegment(3, lin);
            return out;
        }
        // :137-147
        Mat<T> transformPoint(const Mat<T> &p) const { return E * (p - r); }
        Mat<T> inverseTransformPoint(const Mat<T> &p) const { return E.transpose() * p + r; }

        // :150-157   (E1,r1)*(E2,r2) = (E1 E2, r2 + E2^T r1)
        Transform operator*(const Transform &X_in) const
        {
            return Transform(E * X_in.E, X_in.r + X_in.E.transpose() * r);
        }
    };

    // reference: Spatial.h:131-142 (motionCrossProduct)
    template <typename T>
    Mat<T> motionCrossProduct(const Mat<T> &a, const Mat<T> &b)
    {
        Mat<T> mv(6, 1);
        mv[0] = a[1] * b[2] - a[2] * b[1];
        mv[1] = a[2] * b[0] - a[0] * b[2];
        mv[2] = a[0] * b[1] - a[1] * b[0];
        mv[3] = a[1] * b[5] - a[2] * b[4] + a[4] * b[2] - a[5] * b[1];
        mv[4] = a[2] * b[3] - a[0] * b[5] - a[3] * b[2] + a[5] * b[0];
        mv[5] = a[0] * b[4] - a[1] * b[3] + a[3] * b[1] - a[4] * b[0];
        return mv;
    }

    // reference: Spatial.h:176-187 (forceCrossProduct)
    template <typename T>
    Mat<T> forceCrossProduct(const Mat<T> &a, const Mat<T> &b)
    {
        Mat<T> fv(6, 1);
        fv[0] = b[2] * a[1] - b[1] * a[2] - b[4] * a[5] + b[5] * a[4];
        fv[1] = b[0] * a[2] - b[2] * a[0] + b[3] * a[5] - b[5] * a[3];
        fv[2] = b[1] * a[0] - b[0] * a[1] - b[3] * a[4] + b[4] * a[3];
        fv[3] = b[5] * a[1] - b[4] * a[2];
        fv[4] = b[3] * a[2] - b[5] * a[0];
        fv[5] = b[4] * a[0] - b[3] * a[1];
        return fv;
    }

    // reference: Spatial.h:151-170, :196-215 (general*CrossProduct over 6N stacked vectors)
    template <typename T>
    Mat<T> generalMotionCrossProduct(const Mat<T> &a, const Mat<T> &b)
    {
        Mat<T> out(a.r, 1);
        for (int i = 0; i < a.r / 6; i++)
            out.setSegment(6 * i, motionCrossProduct(a.segment(6 * i, 6), b.segment(6 * i, 6)));
        return out;
    }
    template <typename T>
    Mat<T> generalForceCrossProduct(const Mat<T> &a, const Mat<T> &b)
    {
        Mat<T> out(a.r, 1);
        for (int i = 0; i < a.r / 6; i++)
            out.setSegment(6 * i, forceCrossProduct(a.segment(6 * i, 6), b.segment(6 * i, 6)));
        return out;
    }

    // reference: Spatial.h:53-65 (motionCrossMatrix)
    template <typename T>
    Mat<T> motionCrossMatrix(const Mat<T> &v)
    {
        Mat<T> m(6, 6);
        Mat<T> w = skew(v.segment(0, 3)), l = skew(v.segment(3, 3));
        m.setBlock(0, 0, w);
        m.setBlock(3, 0, l);
        m.setBlock(3, 3, w);
        return m;
    }

    // reference: SpatialInertia.h:74-82 (mass, com, rotational inertia about the COM)
    template <typename T>
    Mat<T> spatialInertia(const T &mass, const Mat<T> &com, const Mat<T> &inertia)
    {
        Mat<T> cS = skew(com);
        Mat<T> I(6, 6);
        I.setBlock(0, 0, inertia + mass * (cS * cS.transpose()));
        I.setBlock(0, 3, mass * cS);
        I.setBlock(3, 0, mass * cS.transpose());
        I.setBlock(3, 3, mass * Mat<T>::Identity(3));
        return I;
    }

    // reference: SpatialInertia.h:212-245 (getPseudoInertia, flipAlongAxis) and :130-142
    template <typename T>
    Mat<T> flipAlongAxis(const Mat<T> &I6, Axis axis)
    {
        Mat<T> h = matToSkewVec(I6.block(0, 3, 3, 3));
        Mat<T> Ibar = I6.block(0, 0, 3, 3);
        T m = I6(5, 5);
        T tr = Ibar(0, 0) + Ibar(1, 1) + Ibar(2, 2);
        Mat<T> P(4, 4);
        P.setBlock(0, 0, (T(0.5) * tr) * Mat<T>::Identity(3) - Ibar);
        P.setBlock(0, 3, h);
        P.setBlock(3, 0, h.transpose());
        P(3, 3) = m;
        Mat<T> X = Mat<T>::Identity(4);
        X((int)axis, (int)axis) = T(-1.0);
        P = X * P * X;
        // SpatialInertia(Mat4 P)
        T m2 = P(3, 3);
        Mat<T> h2 = P.block(0, 3, 3, 1);
        Mat<T> E = P.block(0, 0, 3, 3);
        T trE = E(0, 0) + E(1, 1) + E(2, 2);
        Mat<T> out(6, 6);
        out.setBlock(0, 0, trE * Mat<T>::Identity(3) - E);
        out.setBlock(0, 3, skew(h2));
        out.setBlock(3, 0, skew(h2).transpose());
        out.setBlock(3, 3, m2 * Mat<T>::Identity(3));
        return out;
    }

    // reference: SpatialTransforms.cpp:258-477 (GeneralizedTransform): per output body a Transform
    // plus the sub-index of the body of the parent cluster it is expressed relative to.
    template <typename T>
    struct GeneralizedTransform
    {
        int num_parent_bodies = 1;
        std::vector<Transform<T>> X;
        std::vector<int> parent_sub;

        int numOutputBodies() const { return (int)X.size(); }

        // :314-328
        Mat<T> transformMotionVector(const Mat<T> &m_in) const
        {
            Mat<T> out(6 * numOutputBodies(), 1);
            for (int i = 0; i < numOutputBodies(); i++)
                out.setSegment(6 * i, X[i].transformMotionVector(m_in.segment(6 * parent_sub[i], 6)));
            return out;
        }
        // :331-345
        Mat<T> inverseTransformForceVector(const Mat<T> &f_in) const
        {
            Mat<T> out(6 * num_parent_bodies, 1);
            for (int i = 0; i < numOutputBodies(); i++)
                out.addSegment(6 * parent_sub[i], X[i].inverseTransformForceVector(f_in.segment(6 * i, 6)));
            return out;
        }
        // :348-363
        Mat<T> inverseTransformForceSubspace(const Mat<T> &F_in) const
        {
            Mat<T> out(6 * num_parent_bodies, F_in.c);
            for (int i = 0; i < numOutputBodies(); i++)
                for (int j = 0; j < F_in.c; j++)
                    out.addBlock(6 * parent_sub[i], j,
                                 X[i].inverseTransformForceVector(F_in.block(6 * i, j, 6, 1)));
            return out;
        }
        // :366-369, :417-477   X^T I X accumulated into the parent's blocks
        Mat<T> inverseTransformSpatialInertia(const Mat<T> &I_in) const
        {
            const int No = numOutputBodies();
            // rightMultiplyMotionTransform
            Mat<T> M1(6 * No, 6 * num_parent_bodies);
            for (int b = 0; b < No; b++)
            {
                Mat<T> Xm = X[b].toMatrix();
                for (int i = 0; i < 6 * No; i += 6)
                    M1.addBlock(i, 6 * parent_sub[b], I_in.block(i, 6 * b, 6, 6) * Xm);
            }
            // leftMultiplyForceTransform
            Mat<T> M2(6 * num_parent_bodies, 6 * num_parent_bodies);
            for (int b = 0; b < No; b++)
            {
                Mat<T> XmT = X[b].toMatrix().transpose();
                for (int i = 0; i < 6 * num_parent_bodies; i += 6)
                    M2.addBlock(6 * parent_sub[b], i, XmT * M1.block(6 * b, i, 6, 6));
            }
            return M2;
        }
    };

} // namespace grbda_oracle
