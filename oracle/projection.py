"""ORACLE - TEST INFRASTRUCTURE ONLY (second oracle).

The reference checks its cluster algorithms against an independent formulation: the same robot as a spanning tree of
single rigid bodies (one joint per body) with the loop constraints applied to the whole tree at the end
(RigidBodyTreeModel, src/Dynamics/RigidBodyTreeModel.cpp:8-237; the cross checks are
UnitTests/testRigidBodyDynamicsAlgos.cpp:113-239). Restated here in numpy, vectorised over the batch:

  * body-level forward kinematics, RNEA and CRBA on the spanning tree (TreeModel.cpp:7-57,116-212 with one body per
    node: S is the joint's own 6 x 1 axis or I6 for the floating base);
  * inverse dynamics  tau = G^T RNEA(q, G yd, G ydd + g)                (RigidBodyTreeDynamics.cpp:137-146);
  * forward dynamics by projection, "Method 3 in Featherstone Ch 8.5":   (RigidBodyTreeDynamics.cpp:86-97)
        A = G^T H G,  b = tau - G^T (C + H g),  ydd = A^-1 b;
  * mass matrix  G^T H G.

G and g (block diagonal over the clusters) come from the cluster oracle's restated loop constraints per state
(oracle_cluster_constraint); the spanning tree (parents, joint axes, Xtree, inertias) from the model's body table.
Nothing here shares code with the cluster recursion of grbda_oracle/model.h or with the product's compiler.
"""
import numpy as np


def _rot(axis, th):  # ori::coordinateRotation (OrientationTools.h:46-68)
    c, s, o, z = np.cos(th), np.sin(th), np.ones_like(th), np.zeros_like(th)
    if axis == 0:
        R = [[o, z, z], [z, c, s], [z, -s, c]]
    elif axis == 1:
        R = [[c, z, -s], [z, o, z], [s, z, c]]
    else:
        R = [[c, s, z], [-s, c, z], [z, z, o]]
    return np.stack([np.stack(r, -1) for r in R], -2)


def _quat_to_R(q):  # ori::quaternionToRotationMatrix (OrientationTools.h:251-269)
    e0, e1, e2, e3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([np.stack([1 - 2 * (e2 * e2 + e3 * e3), 2 * (e1 * e2 - e0 * e3), 2 * (e1 * e3 + e0 * e2)], -1),
                  np.stack([2 * (e1 * e2 + e0 * e3), 1 - 2 * (e1 * e1 + e3 * e3), 2 * (e2 * e3 - e0 * e1)], -1),
                  np.stack([2 * (e1 * e3 - e0 * e2), 2 * (e2 * e3 + e0 * e1), 1 - 2 * (e1 * e1 + e2 * e2)], -1)], -2)
    return np.swapaxes(R, -1, -2)


def _skew(r):
    z = np.zeros_like(r[..., 0])
    return np.stack([np.stack([z, -r[..., 2], r[..., 1]], -1), np.stack([r[..., 2], z, -r[..., 0]], -1),
                     np.stack([-r[..., 1], r[..., 0], z], -1)], -2)


def _X(E, r):  # spatial::Transform(E, r).toMatrix() (SpatialTransforms.cpp:26-35): motion transform [[E, 0], [-E rx, E]]
    X = np.zeros(E.shape[:-2] + (6, 6))
    X[..., :3, :3] = E
    X[..., 3:, 3:] = E
    X[..., 3:, :3] = -E @ _skew(r)
    return X


def _crm(v):  # motionCrossMatrix (Spatial.h:131-150)
    M = np.zeros(v.shape[:-1] + (6, 6))
    M[..., :3, :3] = _skew(v[..., :3])
    M[..., 3:, 3:] = _skew(v[..., :3])
    M[..., 3:, :3] = _skew(v[..., 3:])
    return M


class SpanningTreeModel:
    """bodies: list of dicts (parent, axis, E, r, inertia) in registration order; clusters: list of dicts
    (first_body, num_bodies, position_index, num_positions, velocity_index, num_velocities, free: 0 no / 7 quaternion
    / 6 roll-pitch-yaw, implicit); cluster_oracle: the OracleModel of the same robot (loop constraints only)."""

    def __init__(self, bodies, clusters, cluster_oracle, gravity=(0.0, 0.0, -9.81)):
        self.bodies, self.clusters, self.o = bodies, clusters, cluster_oracle
        self.gravity = np.array(gravity, dtype=float)
        self.nb = len(bodies)
        self.sdof, n = [], 0
        for c in clusters:
            for i in range(c["num_bodies"]):
                self.sdof.append(n)
                n += 6 if c["free"] else 1
        self.ns = n
        self.nv = sum(c["num_velocities"] for c in clusters)
        self.cluster_of = [ci for ci, c in enumerate(clusters) for _ in range(c["num_bodies"])]

    # ---- per-state loop-constraint quantities: spanning positions, G (ns x nv), g (ns) ----------------------
    def _constraints(self, q, yd):
        B = q.shape[0]
        G = np.zeros((B, self.ns, self.nv))
        g = np.zeros((B, self.ns))
        qs = [None] * self.nb  # spanning position of every body
        for ci, c in enumerate(self.clusters):
            fb, pi, vi, n = c["first_body"], c["position_index"], c["velocity_index"], c["num_velocities"]
            s0 = self.sdof[fb]
            if c["free"]:
                G[:, s0:s0 + 6, vi:vi + 6] = np.eye(6)
                qs[fb] = q[:, pi:pi + c["num_positions"]]
                continue
            for b in range(B):
                Gc, _, gc, _ = self.o.cluster_constraint(ci, q[b], yd[b])
                G[b, s0:s0 + c["num_bodies"], vi:vi + n] = Gc
                g[b, s0:s0 + c["num_bodies"]] = gc
            if c["implicit"]:
                span = q[:, pi:pi + c["num_bodies"]]
            else:  # q_span = gamma(y) = G y (LoopConstraint.cpp:48-52)
                span = np.einsum("bij,bj->bi", G[:, s0:s0 + c["num_bodies"], vi:vi + n], q[:, pi:pi + n])
            for i in range(c["num_bodies"]):
                qs[fb + i] = span[:, i]
        return qs, G, g

    def _kinematics(self, qs, qd_s):
        """Xup, S (6 x dofs), joint velocity vJ of every body."""
        B = qd_s.shape[0]
        Xup, S, vJ = [], [], []
        for i, body in enumerate(self.bodies):
            c = self.clusters[self.cluster_of[i]]
            s0 = self.sdof[i]
            if c["free"]:
                R = _quat_to_R(qs[i][:, 3:7]) if c["free"] == 7 else (
                    _rot(0, qs[i][:, 3]) @ _rot(1, qs[i][:, 4]) @ _rot(2, qs[i][:, 5]))
                Xup.append(_X(R, qs[i][:, 0:3]))  # Xtree ignored (FreeJoint.cpp:45)
                S.append(np.broadcast_to(np.eye(6), (B, 6, 6)))
                vJ.append(qd_s[:, s0:s0 + 6])
            else:
                EJ = _rot(body["axis"], qs[i])
                E = EJ @ np.asarray(body["E"]).reshape(3, 3)
                Xup.append(_X(E, np.broadcast_to(np.asarray(body["r"]).reshape(3), (B, 3))))
                s = np.zeros((B, 6, 1))
                s[:, body["axis"], 0] = 1.0
                S.append(s)
                vJ.append(s[:, :, 0] * qd_s[:, s0:s0 + 1])
        return Xup, S, vJ

    def _rnea(self, Xup, S, vJ, qdd_s):
        B = qdd_s.shape[0]
        a0 = np.zeros((B, 6))
        a0[:, 3:] = -self.gravity
        v, a, f = [None] * self.nb, [None] * self.nb, [None] * self.nb
        for i, body in enumerate(self.bodies):
            p, s0, nd = body["parent"], self.sdof[i], S[i].shape[2]
            vp = v[p] if p >= 0 else np.zeros((B, 6))
            ap = a[p] if p >= 0 else a0
            v[i] = np.einsum("bij,bj->bi", Xup[i], vp) + vJ[i]
            a[i] = (np.einsum("bij,bj->bi", Xup[i], ap) + np.einsum("bij,bj->bi", S[i], qdd_s[:, s0:s0 + nd]) +
                    np.einsum("bij,bj->bi", _crm(v[i]), vJ[i]))
            I = np.asarray(body["inertia"]).reshape(6, 6)
            f[i] = a[i] @ I.T + np.einsum("bij,bj->bi", -np.swapaxes(_crm(v[i]), -1, -2), v[i] @ I.T)
        tau = np.zeros((B, self.ns))
        for i in range(self.nb - 1, -1, -1):
            s0, nd = self.sdof[i], S[i].shape[2]
            tau[:, s0:s0 + nd] = np.einsum("bji,bj->bi", S[i], f[i])
            p = self.bodies[i]["parent"]
            if p >= 0:
                f[p] = f[p] + np.einsum("bji,bj->bi", Xup[i], f[i])
        return tau

    def _crba(self, Xup, S):
        B = Xup[0].shape[0]
        Ic = [np.broadcast_to(np.asarray(b["inertia"]).reshape(6, 6), (B, 6, 6)).copy() for b in self.bodies]
        for i in range(self.nb - 1, -1, -1):
            p = self.bodies[i]["parent"]
            if p >= 0:
                Ic[p] += np.swapaxes(Xup[i], -1, -2) @ Ic[i] @ Xup[i]
        H = np.zeros((B, self.ns, self.ns))
        for i in range(self.nb):
            s0, nd = self.sdof[i], S[i].shape[2]
            F = Ic[i] @ S[i]
            H[:, s0:s0 + nd, s0:s0 + nd] = np.swapaxes(S[i], -1, -2) @ F
            j = i
            while self.bodies[j]["parent"] >= 0:
                F = np.swapaxes(Xup[j], -1, -2) @ F
                j = self.bodies[j]["parent"]
                t0, md = self.sdof[j], S[j].shape[2]
                blk = np.swapaxes(F, -1, -2) @ S[j]
                H[:, s0:s0 + nd, t0:t0 + md] = blk
                H[:, t0:t0 + md, s0:s0 + nd] = np.swapaxes(blk, -1, -2)
        return H

    # ---- the cluster model's entry points through the spanning tree ------------------------------------------
    def inverse_dynamics(self, q, yd, ydd):
        qs, G, g = self._constraints(q, yd)
        qd_s = np.einsum("bij,bj->bi", G, yd)
        Xup, S, vJ = self._kinematics(qs, qd_s)
        tau_s = self._rnea(Xup, S, vJ, np.einsum("bij,bj->bi", G, ydd) + g)
        return np.einsum("bji,bj->bi", G, tau_s)

    def mass_matrix(self, q):
        qs, G, _ = self._constraints(q, np.zeros((q.shape[0], self.nv)))
        Xup, S, _ = self._kinematics(qs, np.zeros((q.shape[0], self.ns)))
        return np.swapaxes(G, -1, -2) @ self._crba(Xup, S) @ G

    def forward_dynamics(self, q, yd, tau):
        qs, G, g = self._constraints(q, yd)
        qd_s = np.einsum("bij,bj->bi", G, yd)
        Xup, S, vJ = self._kinematics(qs, qd_s)
        H = self._crba(Xup, S)
        C = self._rnea(Xup, S, vJ, np.zeros((q.shape[0], self.ns)))
        Gt = np.swapaxes(G, -1, -2)
        A = Gt @ H @ G
        b = tau - np.einsum("bij,bj->bi", Gt, C + np.einsum("bij,bj->bi", H, g))
        return np.linalg.solve(A, b[..., None])[..., 0]


def from_models(product_model, cluster_oracle):
    """Spanning tree from the body table of the product's host model (parents, joint axes, Xtree, inertias) - pinned to
    the reference by tests/test_robot_constants.py and the URDF-vs-builder tests - with the oracle's loop constraints."""
    bodies = [dict(parent=b["parent"], axis=b["axis"], E=b["E"], r=b["r"], inertia=b["inertia"]) for b in product_model.bodies()]
    clusters = []
    for cl, oc in zip(product_model.clusters(), cluster_oracle.clusters()):
        free = oc["num_positions"] if oc["joint_type"] == "Free" else 0
        clusters.append(dict(first_body=cl["first_body"], num_bodies=cl["num_bodies"], position_index=oc["position_index"],
                             num_positions=oc["num_positions"], velocity_index=oc["velocity_index"],
                             num_velocities=oc["num_velocities"], free=free, implicit=bool(oc["implicit"])))
    return SpanningTreeModel(bodies, clusters, cluster_oracle, gravity=tuple(product_model.getGravity()))
