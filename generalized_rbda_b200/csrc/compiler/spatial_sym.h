// Spatial algebra over symbolic scalars (E, r) transforms, rigid 10-parameter and general
// symmetric 6x6 inertias. Same conventions as the reference (Featherstone, angular first):
// reference: src/Utils/SpatialTransforms.cpp:14-197, include/grbda/Utils/Spatial.h:131-187,
// include/grbda/Utils/SpatialInertia.h:74-82, OrientationTools.h:46-68,251-269.
// Because every operation goes through sym::Sym, structural zeros/ones of a particular model
// (identity Xtree rotations, zero offsets, diagonal rotor inertias ...) vanish at compile time.
#pragma once
#include <array>
#include "../host/types.h"
#include "sym.h"

namespace grbda
{
    namespace compiler
    {
        using sym::Sym;

        struct V3
        {
            Sym v[3];
            Sym &operator[](int i) { return v[i]; }
            const Sym &operator[](int i) const { return v[i]; }
        };
        struct M3
        {
            Sym m[9]; // row-major
            Sym &operator()(int i, int j) { return m[3 * i + j]; }
            const Sym &operator()(int i, int j) const { return m[3 * i + j]; }
        };
        struct SV // spatial (6D) vector
        {
            Sym v[6];
            Sym &operator[](int i) { return v[i]; }
            const Sym &operator[](int i) const { return v[i]; }
            V3 ang() const { return V3{{v[0], v[1], v[2]}}; }
            V3 lin() const { return V3{{v[3], v[4], v[5]}}; }
            static SV make(const V3 &a, const V3 &l) { return SV{{a[0], a[1], a[2], l[0], l[1], l[2]}}; }
        };

        inline V3 operator+(const V3 &a, const V3 &b) { return V3{{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
        inline V3 operator-(const V3 &a, const V3 &b) { return V3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
        inline V3 operator-(const V3 &a) { return V3{{-a[0], -a[1], -a[2]}}; }
        inline V3 operator*(const Sym &s, const V3 &a) { return V3{{s * a[0], s * a[1], s * a[2]}}; }
        inline Sym dot(const V3 &a, const V3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
        inline V3 cross(const V3 &a, const V3 &b)
        {
            return V3{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
        }
        inline V3 mul(const M3 &A, const V3 &x)
        {
            V3 y;
            for (int i = 0; i < 3; i++)
                y[i] = A(i, 0) * x[0] + A(i, 1) * x[1] + A(i, 2) * x[2];
            return y;
        }
        inline V3 mulT(const M3 &A, const V3 &x) // A^T x
        {
            V3 y;
            for (int i = 0; i < 3; i++)
                y[i] = A(0, i) * x[0] + A(1, i) * x[1] + A(2, i) * x[2];
            return y;
        }
        inline M3 mul(const M3 &A, const M3 &B)
        {
            M3 C;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
            return C;
        }
        inline M3 transpose(const M3 &A)
        {
            M3 C;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    C(i, j) = A(j, i);
            return C;
        }
        inline M3 constM3(const Mat3 &A)
        {
            M3 C;
            for (int i = 0; i < 9; i++)
                C.m[i] = Sym(A[i]);
            return C;
        }
        inline V3 constV3(const Vec3 &a) { return V3{{Sym(a[0]), Sym(a[1]), Sym(a[2])}}; }

        inline SV operator+(const SV &a, const SV &b)
        {
            SV c;
            for (int i = 0; i < 6; i++)
                c[i] = a[i] + b[i];
            return c;
        }
        inline SV operator-(const SV &a, const SV &b)
        {
            SV c;
            for (int i = 0; i < 6; i++)
                c[i] = a[i] - b[i];
            return c;
        }
        inline SV operator*(const Sym &s, const SV &a)
        {
            SV c;
            for (int i = 0; i < 6; i++)
                c[i] = s * a[i];
            return c;
        }
        inline Sym dot(const SV &a, const SV &b)
        {
            Sym s = a[0] * b[0];
            for (int i = 1; i < 6; i++)
                s = s + a[i] * b[i];
            return s;
        }
        // reference: Spatial.h:131-142
        inline SV motionCross(const SV &a, const SV &b)
        {
            return SV::make(cross(a.ang(), b.ang()), cross(a.ang(), b.lin()) + cross(a.lin(), b.ang()));
        }
        // reference: Spatial.h:176-187   a x* f
        inline SV forceCross(const SV &a, const SV &f)
        {
            return SV::make(cross(a.ang(), f.ang()) + cross(a.lin(), f.lin()), cross(a.ang(), f.lin()));
        }

        // reference: OrientationTools.h:46-68
        inline M3 coordinateRotation(ori::CoordinateAxis axis, const Sym &s, const Sym &c)
        {
            M3 R;
            const Sym one(1.0), zero(0.0);
            if (axis == ori::CoordinateAxis::X)
                R = M3{{one, zero, zero, zero, c, s, zero, -s, c}};
            else if (axis == ori::CoordinateAxis::Y)
                R = M3{{c, zero, -s, zero, one, zero, s, zero, c}};
            else
                R = M3{{c, s, zero, -s, c, zero, zero, zero, one}};
            return R;
        }
        // reference: OrientationTools.h:251-269
        inline M3 quaternionToRotationMatrix(const Sym &e0, const Sym &e1, const Sym &e2, const Sym &e3)
        {
            const Sym one(1.0), two(2.0);
            M3 R;
            R(0, 0) = one - two * (e2 * e2 + e3 * e3);
            R(0, 1) = two * (e1 * e2 - e0 * e3);
            R(0, 2) = two * (e1 * e3 + e0 * e2);
            R(1, 0) = two * (e1 * e2 + e0 * e3);
            R(1, 1) = one - two * (e1 * e1 + e3 * e3);
            R(1, 2) = two * (e2 * e3 - e0 * e1);
            R(2, 0) = two * (e1 * e3 - e0 * e2);
            R(2, 1) = two * (e2 * e3 + e0 * e1);
            R(2, 2) = one - two * (e1 * e1 + e2 * e2);
            return transpose(R);
        }

        // reference: SpatialTransforms.cpp:14-197 (spatial::Transform)
        struct Xf
        {
            M3 E;
            V3 r;
            static Xf identity()
            {
                Xf X;
                X.E = constM3(ori::identity3());
                return X;
            }
            // :43-50   [E w; E (v - r x w)]
            SV applyMotion(const SV &m) const
            {
                return SV::make(mul(E, m.ang()), mul(E, m.lin() - cross(r, m.ang())));
            }
            // :74-82   [E^T n + r x (E^T f); E^T f]
            SV applyForceTranspose(const SV &f) const
            {
                const V3 l = mulT(E, f.lin());
                return SV::make(mulT(E, f.ang()) + cross(r, l), l);
            }
            // :64-71   [E (n - r x f); E f]
            SV applyForce(const SV &f) const
            {
                return SV::make(mul(E, f.ang() - cross(r, f.lin())), mul(E, f.lin()));
            }
            // :150-157  (E1, r1) * (E2, r2) = (E1 E2, r2 + E2^T r1)
            Xf operator*(const Xf &B) const
            {
                Xf C;
                C.E = mul(E, B.E);
                C.r = B.r + mulT(B.E, r);
                return C;
            }
        };

        // Symmetric 3x3 stored as xx, xy, xz, yy, yz, zz
        struct Sym3
        {
            Sym a[6];
            Sym at(int i, int j) const
            {
                static const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
                return a[idx[i][j]];
            }
            M3 full() const
            {
                M3 M;
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++)
                        M(i, j) = at(i, j);
                return M;
            }
            static Sym3 fromUpper(const M3 &M)
            {
                return Sym3{{M(0, 0), M(0, 1), M(0, 2), M(1, 1), M(1, 2), M(2, 2)}};
            }
        };
        inline Sym3 operator+(const Sym3 &x, const Sym3 &y)
        {
            Sym3 z;
            for (int i = 0; i < 6; i++)
                z.a[i] = x.a[i] + y.a[i];
            return z;
        }
        inline Sym3 operator-(const Sym3 &x, const Sym3 &y)
        {
            Sym3 z;
            for (int i = 0; i < 6; i++)
                z.a[i] = x.a[i] - y.a[i];
            return z;
        }
        // E^T A E for symmetric A (upper triangle only)
        inline Sym3 rotateToParent(const M3 &E, const Sym3 &A)
        {
            const M3 AE = mul(A.full(), E);
            Sym3 out;
            int k = 0;
            for (int i = 0; i < 3; i++)
                for (int j = i; j < 3; j++)
                    out.a[k++] = E(0, i) * AE(0, j) + E(1, i) * AE(1, j) + E(2, i) * AE(2, j);
            return out;
        }
        // E^T B E for a general 3x3
        inline M3 rotateToParent(const M3 &E, const M3 &B) { return mul(transpose(E), mul(B, E)); }

        // a b^T + b a^T - 2 (a.b) 1   ( = -(a^ b^ + b^ a^) )
        inline Sym3 symOuterMinusDot(const V3 &a, const V3 &b)
        {
            const Sym d = dot(a, b);
            Sym3 out;
            int k = 0;
            for (int i = 0; i < 3; i++)
                for (int j = i; j < 3; j++)
                {
                    Sym e = a[i] * b[j] + b[i] * a[j];
                    if (i == j)
                        e = e - Sym(2.0) * d;
                    out.a[k++] = e;
                }
            return out;
        }

        // Rigid-body inertia [[Ibar, h^],[h^T, m 1]] in 10 parameters (h = m c, Ibar about the frame
        // origin). reference: SpatialInertia.h:74-82 builds the same 6x6 matrix.
        struct RigidInertia
        {
            Sym m;
            V3 h;
            Sym3 Ibar;

            static RigidInertia fromMatrix(const Mat6 &I)
            {
                RigidInertia R;
                R.m = Sym(I[35]);
                // top-right block is skew(h)
                R.h = V3{{Sym(0.5 * (I[6 * 2 + 4] - I[6 * 1 + 5])), Sym(0.5 * (I[6 * 0 + 5] - I[6 * 2 + 3])),
                          Sym(0.5 * (I[6 * 1 + 3] - I[6 * 0 + 4]))}};
                R.Ibar = Sym3{{Sym(I[0]), Sym(0.5 * (I[1] + I[6])), Sym(0.5 * (I[2] + I[12])), Sym(I[7]),
                               Sym(0.5 * (I[8] + I[13])), Sym(I[14])}};
                return R;
            }
            // I * a  = [Ibar w + h x v ; m v - h x w]
            SV apply(const SV &a) const
            {
                return SV::make(mul(Ibar.full(), a.ang()) + cross(h, a.lin()), m * a.lin() - cross(h, a.ang()));
            }
            // X^T I X for X = (E, r): inertia expressed in the parent frame
            RigidInertia toParent(const Xf &X) const
            {
                RigidInertia P;
                P.m = m;
                const V3 hr = mulT(X.E, h);
                P.h = hr + m * X.r;
                // Ibar' = E^T Ibar E - hr^ r^ - r^ hr^ - m r^ r^
                const Sym3 A = rotateToParent(X.E, Ibar);
                const Sym3 t1 = symOuterMinusDot(X.r, hr);
                const Sym3 rr = symOuterMinusDot(X.r, X.r); // 2 (r r^T - |r|^2 1)
                P.Ibar = A - t1;
                for (int i = 0; i < 6; i++)
                    P.Ibar.a[i] = P.Ibar.a[i] - (Sym(0.5) * m) * rr.a[i];
                return P;
            }
        };
        inline RigidInertia operator+(const RigidInertia &x, const RigidInertia &y)
        {
            RigidInertia z;
            z.m = x.m + y.m;
            z.h = x.h + y.h;
            z.Ibar = x.Ibar + y.Ibar;
            return z;
        }

        // General 6x6 block (not necessarily symmetric)
        struct M6
        {
            Sym m[36];
            Sym &operator()(int i, int j) { return m[6 * i + j]; }
            const Sym &operator()(int i, int j) const { return m[6 * i + j]; }
            SV apply(const SV &x) const
            {
                SV y;
                for (int i = 0; i < 6; i++)
                {
                    Sym s = (*this)(i, 0) * x[0];
                    for (int j = 1; j < 6; j++)
                        s = s + (*this)(i, j) * x[j];
                    y[i] = s;
                }
                return y;
            }
            SV applyTranspose(const SV &x) const
            {
                SV y;
                for (int i = 0; i < 6; i++)
                {
                    Sym s = (*this)(0, i) * x[0];
                    for (int j = 1; j < 6; j++)
                        s = s + (*this)(j, i) * x[j];
                    y[i] = s;
                }
                return y;
            }
            bool isZero() const
            {
                for (int i = 0; i < 36; i++)
                    if (!m[i].isZero())
                        return false;
                return true;
            }
            M6 transposed() const
            {
                M6 t;
                for (int i = 0; i < 6; i++)
                    for (int j = 0; j < 6; j++)
                        t(i, j) = (*this)(j, i);
                return t;
            }
        };
        inline M6 operator+(const M6 &x, const M6 &y)
        {
            M6 z;
            for (int i = 0; i < 36; i++)
                z.m[i] = x.m[i] + y.m[i];
            return z;
        }
        inline M6 operator-(const M6 &x, const M6 &y)
        {
            M6 z;
            for (int i = 0; i < 36; i++)
                z.m[i] = x.m[i] - y.m[i];
            return z;
        }

        // Symmetric 6x6 [[A, B],[B^T, C]] with A, C symmetric: articulated-body inertia.
        struct SymInertia
        {
            Sym3 A, C;
            M3 B;
            static SymInertia fromRigid(const RigidInertia &R)
            {
                SymInertia I;
                I.A = R.Ibar;
                const Sym z(0.0);
                I.B = M3{{z, -R.h[2], R.h[1], R.h[2], z, -R.h[0], -R.h[1], R.h[0], z}};
                I.C = Sym3{{R.m, z, z, R.m, z, R.m}};
                return I;
            }
            Sym at(int i, int j) const
            {
                if (i < 3 && j < 3)
                    return A.at(i, j);
                if (i >= 3 && j >= 3)
                    return C.at(i - 3, j - 3);
                if (i < 3)
                    return B(i, j - 3);
                return B(j, i - 3);
            }
            SV apply(const SV &x) const
            {
                SV y;
                for (int i = 0; i < 6; i++)
                {
                    Sym s = at(i, 0) * x[0];
                    for (int j = 1; j < 6; j++)
                        s = s + at(i, j) * x[j];
                    y[i] = s;
                }
                return y;
            }
            M6 full() const
            {
                M6 M;
                for (int i = 0; i < 6; i++)
                    for (int j = 0; j < 6; j++)
                        M(i, j) = at(i, j);
                return M;
            }
            // X^T I X (reference: SpatialTransforms.cpp:113-135)
            SymInertia toParent(const Xf &X) const
            {
                const Sym3 Ar = rotateToParent(X.E, A), Cr = rotateToParent(X.E, C);
                const M3 Br = rotateToParent(X.E, B);
                const Sym z(0.0);
                const M3 rh{{z, -X.r[2], X.r[1], X.r[2], z, -X.r[0], -X.r[1], X.r[0], z}};
                SymInertia P;
                P.C = Cr;
                const M3 rC = mul(rh, Cr.full());
                // B' = Br + r^ Cr
                for (int i = 0; i < 9; i++)
                    P.B.m[i] = Br.m[i] + rC.m[i];
                // A' = Ar - Br r^ + r^ Br^T - r^ Cr r^  = Ar - B' r^ + r^ Br^T   (symmetric)
                const M3 Bpr = mul(P.B, rh);
                const M3 rBt = mul(rh, transpose(Br));
                int k = 0;
                for (int i = 0; i < 3; i++)
                    for (int j = i; j < 3; j++)
                    {
                        P.A.a[k] = Ar.a[k] - Bpr(i, j) + rBt(i, j);
                        k++;
                    }
                return P;
            }
        };
        inline SymInertia operator+(const SymInertia &x, const SymInertia &y)
        {
            SymInertia z;
            z.A = x.A + y.A;
            z.C = x.C + y.C;
            for (int i = 0; i < 9; i++)
                z.B.m[i] = x.B.m[i] + y.B.m[i];
            return z;
        }

        // 6x6 matrix of the motion transform (reference: SpatialTransforms.cpp:33-40)
        inline M6 toMatrix(const Xf &X)
        {
            M6 M;
            const Sym z(0.0);
            const M3 rh{{z, -X.r[2], X.r[1], X.r[2], z, -X.r[0], -X.r[1], X.r[0], z}};
            const M3 Erh = mul(X.E, rh);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                {
                    M(i, j) = X.E(i, j);
                    M(3 + i, 3 + j) = X.E(i, j);
                    M(3 + i, j) = -Erh(i, j);
                }
            return M;
        }
        // Xi^T M Xj for a general (off-diagonal) block
        inline M6 transformBlock(const Xf &Xi, const M6 &M, const Xf &Xj)
        {
            const M6 A = toMatrix(Xi), B = toMatrix(Xj);
            M6 MB, out;
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 6; j++)
                {
                    Sym s(0.0);
                    for (int k = 0; k < 6; k++)
                        s = s + M(i, k) * B(k, j);
                    MB(i, j) = s;
                }
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 6; j++)
                {
                    Sym s(0.0);
                    for (int k = 0; k < 6; k++)
                        s = s + A(k, i) * MB(k, j);
                    out(i, j) = s;
                }
            return out;
        }

    } // namespace compiler
} // namespace grbda
