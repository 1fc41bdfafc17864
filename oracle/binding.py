"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (CPU restatement of the reference's ClusterTreeModel hot
path) and, when present, oracle/_ref/libgrbda_codegen.so (the reference's own CasADi-generated
closed-form dynamics compiled from /root/reference/src/Codegen).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package generalized_rbda_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libgrbda_codegen.so")

_dp = C.POINTER(C.c_double)


def build():
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    r = subprocess.run(["make", "-C", HERE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout)
    return r.stdout


def _P(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_model_create.restype = C.c_void_p
        _lib.oracle_builder_create.restype = C.c_void_p
        _lib.oracle_last_error.restype = C.c_char_p
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().oracle_last_error().decode())


class OracleModel:
    """One ClusterTreeModel of the oracle. name: a robot of oracle/grbda_oracle/robots.h
    (append ':generic' for the Generic re-build of UnitTests/testHelpers.hpp)."""

    def __init__(self, name=None, handle=None):
        L = lib()
        if handle is None:
            handle = L.oracle_model_create(name.encode())
            if not handle:
                raise OracleError(L.oracle_last_error().decode())
        self._h = C.c_void_p(handle)
        self.nq = L.oracle_num_positions(self._h)
        self.nv = L.oracle_num_dof(self._h)
        self.nb = L.oracle_num_bodies(self._h)
        self.nc = L.oracle_num_clusters(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.oracle_model_destroy(self._h)
            self._h = None

    def clusters(self):
        info = (C.c_int * (7 * self.nc))()
        names = C.create_string_buffer(64 * self.nc)
        lib().oracle_cluster_info(self._h, info, names)
        keys = ("parent", "num_bodies", "num_positions", "num_velocities", "position_index",
                "velocity_index", "implicit")
        out = []
        for c in range(self.nc):
            d = dict(zip(keys, info[7 * c:7 * c + 7]))
            d["joint_type"] = names.raw[64 * c:64 * c + 64].split(b"\0")[0].decode()
            out.append(d)
        return out

    def bodies(self):
        out = []
        for b in range(self.nb):
            name = C.create_string_buffer(64)
            ints = (C.c_int * 3)()
            E, r, I = np.zeros((3, 3)), np.zeros(3), np.zeros((6, 6))
            lib().oracle_body_info(self._h, b, name, ints, _P(E), _P(r), _P(I))
            out.append(dict(name=name.value.decode(), parent=ints[0], cluster=ints[1], sub_index=ints[2],
                            E=E, r=r, inertia=I))
        return out

    def generate_states(self, count, seed=0x6772626461, first_index=0, threads=0):
        q = np.zeros((count, self.nq))
        yd = np.zeros((count, self.nv))
        aux = np.zeros((count, self.nv))
        _check(lib().oracle_generate_states(self._h, C.c_uint64(seed), C.c_int64(first_index), C.c_int64(count),
                                            _P(q), _P(yd), _P(aux), threads))
        return q, yd, aux

    def inverse_dynamics(self, q, yd, ydd, threads=0):
        tau = np.zeros_like(ydd)
        _check(lib().oracle_inverse_dynamics(self._h, _P(q), _P(yd), _P(ydd), _P(tau), C.c_int64(q.shape[0]), threads))
        return tau

    def forward_dynamics(self, q, yd, tau, threads=0):
        ydd = np.zeros_like(tau)
        _check(lib().oracle_forward_dynamics(self._h, _P(q), _P(yd), _P(tau), _P(ydd), C.c_int64(q.shape[0]), threads))
        return ydd

    def dynamics_derivatives(self, q, yd, in3, forward, threads=0):
        """Jacobians of inverse (forward=False: in3 = ydd) or forward dynamics (in3 = tau) with respect to the
        tangent-space perturbation dq, yd and (forward only) tau; each [batch, nv, nv], [b, i, j] = d out_i / d x_j."""
        B = q.shape[0]
        outs = [np.zeros((B, self.nv, self.nv)) for _ in range(3 if forward else 2)]
        _check(lib().oracle_dynamics_derivatives(self._h, int(bool(forward)), _P(q), _P(yd), _P(in3), _P(outs[0]), _P(outs[1]),
                                                 _P(outs[2]) if forward else None, C.c_int64(B), threads))
        return tuple(outs)

    def dynamics_with_external_forces(self, q, yd, in3, f_ext, forward, threads=0):
        out = np.zeros_like(in3)
        _check(lib().oracle_dynamics_with_external_forces(self._h, _P(q), _P(yd), _P(in3),
                                                          _P(f_ext) if f_ext is not None else None, _P(out),
                                                          int(forward), C.c_int64(q.shape[0]), threads))
        return out

    def mass_matrix(self, q, threads=0):
        H = np.zeros((q.shape[0], self.nv, self.nv))
        _check(lib().oracle_mass_matrix(self._h, _P(q), _P(H), C.c_int64(q.shape[0]), threads))
        return H

    def forward_kinematics(self, q, yd, threads=0):
        B = q.shape[0]
        p, R, v = np.zeros((B, self.nb, 3)), np.zeros((B, self.nb, 3, 3)), np.zeros((B, self.nb, 6))
        _check(lib().oracle_forward_kinematics(self._h, _P(q), _P(yd), _P(p), _P(R), _P(v), C.c_int64(B), threads))
        return p, R, v

    def validate_states(self, q, threads=0):
        valid = np.zeros(q.shape[0], dtype=np.int32)
        _check(lib().oracle_validate_states(self._h, _P(q), valid.ctypes.data_as(C.POINTER(C.c_int)),
                                            C.c_int64(q.shape[0]), threads))
        return valid.astype(bool)

    # ---- operational space (contact points) ---------------------------------------------------------
    def set_contact_points(self, bodies, offsets, end_effector=None):
        n = len(bodies)
        b = (C.c_int * max(1, n))(*[int(x) for x in bodies])
        o = np.ascontiguousarray(offsets, dtype=np.float64).reshape(n, 3)
        e = (C.c_int * max(1, n))(*[int(x) for x in (end_effector if end_effector is not None else [0] * n)])
        _check(lib().oracle_set_contact_points(self._h, n, b, _P(o), e))
        self.ncp, self.nee = n, int(sum(e[:n]))

    def contact_kinematics(self, q, yd, threads=0):
        p, v = np.zeros((q.shape[0], self.ncp, 3)), np.zeros((q.shape[0], self.ncp, 3))
        _check(lib().oracle_contact_kinematics(self._h, _P(q), _P(yd), _P(p), _P(v), C.c_int64(q.shape[0]), threads))
        return p, v

    def contact_jacobians(self, q, world=True, threads=0):
        J = np.zeros((q.shape[0], self.ncp, 6, self.nv))
        _check(lib().oracle_contact_jacobians(self._h, _P(q), _P(J), int(world), C.c_int64(q.shape[0]), threads))
        return J

    def apply_test_force(self, q, force, threads=0):
        d, lam = np.zeros((q.shape[0], self.ncp, self.nv)), np.zeros((q.shape[0], self.ncp))
        f = np.ascontiguousarray(force, dtype=np.float64)
        _check(lib().oracle_apply_test_force(self._h, _P(q), _P(f), _P(d), _P(lam), C.c_int64(q.shape[0]), threads))
        return d, lam

    def inverse_osim(self, q, threads=0):
        n = 6 * self.nee
        L = np.zeros((q.shape[0], n, n))
        _check(lib().oracle_inverse_osim(self._h, _P(q), _P(L), C.c_int64(q.shape[0]), threads))
        return L

    def integrate(self, q, yd, ydd, dt, threads=0):
        """(q', yd', flags) of the semi-implicit Euler step (oracle/grbda_oracle/rng.h integrateState)."""
        qo, ydo = np.zeros_like(q), np.zeros_like(yd)
        flags = np.zeros(q.shape[0], dtype=np.int32)
        _check(lib().oracle_integrate(self._h, _P(q), _P(yd), _P(ydd), C.c_double(dt), _P(qo), _P(ydo),
                                      flags.ctypes.data_as(C.POINTER(C.c_int)), C.c_int64(q.shape[0]), threads))
        return qo, ydo, flags

    def cluster_constraint(self, cluster, q, yd):
        G, K, g, k = np.zeros(1024), np.zeros(1024), np.zeros(64), np.zeros(64)
        dims = (C.c_int * 4)()
        _check(lib().oracle_cluster_constraint(self._h, cluster, _P(q), _P(yd), _P(G), _P(K), _P(g), _P(k), dims))
        N, n, nc = dims[0], dims[1], dims[2]
        return (G[:N * n].reshape(N, n).copy(), K[:nc * N].reshape(nc, N).copy(), g[:N].copy(), k[:nc].copy())

    def count_flops(self, algo):
        """algo: 0 ID, 1 FD, 2 FK, 3 H -> dict of operation counts of one evaluation."""
        out = (C.c_uint64 * 10)()
        _check(lib().oracle_count_flops(self._h, algo, out))
        v = [int(x) for x in out]
        return dict(flops_all=sum(v[0:4]), trig_all=v[4], flops_alg=sum(v[5:9]), trig_alg=v[9],
                    add_alg=v[5], mul_alg=v[6], div_alg=v[7], sqrt_alg=v[8])


def max_threads():
    return lib().oracle_max_threads()


class OracleBuilder:
    """Assemble an oracle model through the reference's registerBody / append* call sequence."""

    def __init__(self, gravity=(0.0, 0.0, -9.81)):
        g = np.array(gravity, dtype=np.float64)
        self._h = C.c_void_p(lib().oracle_builder_create(_P(g)))

    def register_body(self, name, parent, inertia, E, r):
        lib().oracle_builder_register_body(self._h, name.encode(), parent.encode(),
                                           _P(np.ascontiguousarray(inertia, dtype=np.float64)),
                                           _P(np.ascontiguousarray(E, dtype=np.float64)),
                                           _P(np.ascontiguousarray(r, dtype=np.float64)))

    def append_simple(self, name, kind, axes=None, gear_ratio=0.0):
        """kind: 0 Free(quat) 1 Free(rpy) 2 Revolute 3 RevoluteWithRotor 7 RevolutePair"""
        ax = (C.c_int * len(axes))(*axes) if axes else None
        lib().oracle_builder_append_simple(self._h, name.encode(), kind, ax, C.c_double(gear_ratio))

    def append_generic_static(self, name, axes, K, G):
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1, len(axes))
        G = np.ascontiguousarray(G, dtype=np.float64).reshape(len(axes), -1)
        lib().oracle_builder_append_generic_static(self._h, name.encode(), (C.c_int * len(axes))(*axes),
                                                   _P(K), K.shape[0], _P(G), G.shape[1])

    def add_loop(self, pred_chain, succ_chain, pred_E, pred_r, succ_E, succ_r, keep_rows):
        lib().oracle_builder_add_loop(self._h, (C.c_int * len(pred_chain))(*pred_chain), len(pred_chain),
                                      (C.c_int * len(succ_chain))(*succ_chain), len(succ_chain),
                                      _P(np.ascontiguousarray(pred_E, dtype=np.float64)),
                                      _P(np.ascontiguousarray(pred_r, dtype=np.float64)),
                                      _P(np.ascontiguousarray(succ_E, dtype=np.float64)),
                                      _P(np.ascontiguousarray(succ_r, dtype=np.float64)),
                                      (C.c_int * 3)(*keep_rows))

    def append_generic_loops(self, name, axes, independent):
        lib().oracle_builder_append_generic_loops(self._h, name.encode(), (C.c_int * len(axes))(*axes),
                                                  (C.c_int * len(axes))(*[int(x) for x in independent]))

    def append_four_bar(self, name, axes, path1, path2, offset, independent_coordinate):
        """ClusterJoints::FourBar with the analytic LoopConstraint::FourBar (FourBarJoint.cpp:7-199)."""
        p1, p2, off = (np.ascontiguousarray(x, dtype=np.float64) for x in (path1, path2, offset))
        lib().oracle_builder_append_four_bar(self._h, name.encode(), (C.c_int * len(axes))(*axes), _P(p1), len(p1),
                                             _P(p2), len(p2), _P(off), int(independent_coordinate))

    def append_generic_phi(self, name, axes, independent, ops, outputs):
        """ops: structured array with fields op, a, b, val (generalized_rbda_b200.PHI_OP_DTYPE)"""
        ci = lambda x: np.ascontiguousarray(x, dtype=np.int32)
        op, a, b = ci(ops["op"]), ci(ops["a"]), ci(ops["b"])
        val = np.ascontiguousarray(ops["val"], dtype=np.float64)
        out = ci(outputs)
        ip = C.POINTER(C.c_int)
        lib().oracle_builder_append_generic_phi(self._h, name.encode(), (C.c_int * len(axes))(*axes),
                                                (C.c_int * len(axes))(*[int(x) for x in independent]),
                                                op.ctypes.data_as(ip), a.ctypes.data_as(ip), b.ctypes.data_as(ip),
                                                _P(val), len(op), out.ctypes.data_as(ip), len(out))

    def append_revolute_triple_with_rotor(self, name, axes, gears, belts):
        """bodies registered as link1, link2, link3, rotor1, rotor2, rotor3; belts = 1 + 2 + 3 ratios"""
        g = np.ascontiguousarray(gears, dtype=np.float64)
        b = np.ascontiguousarray(belts, dtype=np.float64)
        assert g.size == 3 and b.size == 6 and len(axes) == 6
        lib().oracle_builder_append_revolute_triple_with_rotor(self._h, name.encode(), (C.c_int * 6)(*axes), _P(g), _P(b))

    def finish(self, generic=False):
        _check(lib().oracle_builder_finish(self._h, int(generic)))
        h, self._h = self._h, None
        return OracleModel(handle=h.value)


# ---- the reference's own generated closed-form dynamics (known-answer functions) -----------------
_ref = None


def reference_codegen_available():
    return os.path.exists(REF_LIB_PATH)


def reference_codegen(name, args, n_out):
    """Call one of /root/reference/src/Codegen's functions, e.g. 'RevWithRotors2DofFwdDyn'
    with args = [y, yd, tau] (reference: include/grbda/Codegen/*.h, CasadiGen.cpp:8-83)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_LIB_PATH)
    f = getattr(_ref, name)
    work = getattr(_ref, name + "_work")
    sz = [C.c_longlong() for _ in range(4)]
    work(*[C.byref(s) for s in sz])
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in args]
    argv = (_dp * max(sz[0].value, len(args)))(*[_P(a) for a in args])
    out = np.zeros(n_out)
    resv = (_dp * max(sz[1].value, 1))(_P(out))
    iw = (C.c_longlong * max(sz[2].value, 1))()
    w = (C.c_double * max(sz[3].value, 1))()
    f(argv, resv, iw, w, 0)
    return out
