"""Derivatives of the dynamics.

The reference has no closed-form derivative algorithm: UnitTests/testRigidBodyDynamicsAlgosDerivatives.cpp
takes CasADi's jacobian() of the symbolic model with respect to a tangent-space perturbation dq of the positions
(UnitTests/testHelpers.hpp:50-112, `plus`), the velocities and the third argument, and checks it against central
finite differences (:271-383, tolerances 2e-5 / 1e-5). Restated here in two layers:
  * the oracle (dual-number arithmetic through the CPU restatement) against central finite differences through
    the same `plus` - the reference test itself, on its own robot list;
  * the product's derivative programs (compile-time differentiation of the model's program) against the oracle:
    as numpy-replayed tapes on the CPU, and through the C ABI on the GPU.
"""
import numpy as np
import pytest

from tape import load_tape, run_tape

# the reference test's robots that exist here (+ the pair chain), and the headline model with implicit clusters
REFERENCE_ROBOTS = ["revolute_chain_with_rotor_2", "revolute_chain_with_rotor_4", "revolute_chain_with_rotor_8",
                    "mini_cheetah", "mit_humanoid"]
ALL_ROBOTS = REFERENCE_ROBOTS + ["revolute_pair_chain_with_rotor_4", "tello_with_arms"]
TOL = 1e-10  # tapes / GPU against the oracle, relative to the largest entry of each matrix


def rel(a, b):
    a, b = a.reshape(b.shape[0], -1), b.reshape(b.shape[0], -1)
    return float((np.abs(a - b).max(1) / np.maximum(1e-300, np.abs(b).max(1))).max())


def quat_to_R(q):  # ori::quaternionToRotationMatrix (OrientationTools.h:251-269), q = (w, x, y, z)
    e0, e1, e2, e3 = q
    R = np.array([[1 - 2 * (e2 * e2 + e3 * e3), 2 * (e1 * e2 - e0 * e3), 2 * (e1 * e3 + e0 * e2)],
                  [2 * (e1 * e2 + e0 * e3), 1 - 2 * (e1 * e1 + e3 * e3), 2 * (e2 * e3 - e0 * e1)],
                  [2 * (e1 * e3 - e0 * e2), 2 * (e2 * e3 + e0 * e1), 1 - 2 * (e1 * e1 + e2 * e2)]])
    return R.T


def quat_product(a, b):  # ori::quatProduct (OrientationTools.h:365-377)
    r = a[0] * b[0] - a[1:] @ b[1:]
    v = a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])
    return np.concatenate([[r], v])


def plus(q, dq):
    """TestHelpers::plus (testHelpers.hpp:79-112) for one state: floating base first when nq != nv."""
    out = q.copy()
    pi = vi = 0
    if q.size != dq.size:
        R = quat_to_R(q[3:7])
        out[0:3] = q[0:3] + R.T @ dq[3:6]
        out[3:7] = q[3:7] + 0.5 * quat_product(q[3:7], np.concatenate([[0.0], dq[0:3]]))
        pi, vi = 7, 6
    out[pi:] = q[pi:] + dq[vi:]
    return out


@pytest.mark.parametrize("robot", REFERENCE_ROBOTS)
def test_oracle_derivatives_against_finite_differences(oracle, robot):
    """The reference's test (rnea, :339-383) on the oracle: Jacobians against central differences."""
    o = oracle.OracleModel(robot)
    q, yd, aux = o.generate_states(3, seed=11)[:3]
    h = 1e-6
    for forward in (False, True):
        f = o.forward_dynamics if forward else o.inverse_dynamics
        d = o.dynamics_derivatives(q, yd, aux, forward)
        num = [np.zeros_like(d[0]) for _ in d]
        for j in range(o.nv):
            e = np.zeros(o.nv)
            e[j] = h
            qp = np.stack([plus(q[b], e) for b in range(q.shape[0])])
            qm = np.stack([plus(q[b], -e) for b in range(q.shape[0])])
            num[0][:, :, j] = (f(qp, yd, aux) - f(qm, yd, aux)) / (2 * h)
            num[1][:, :, j] = (f(q, yd + e, aux) - f(q, yd - e, aux)) / (2 * h)
            if forward:
                num[2][:, :, j] = (f(q, yd, aux + e) - f(q, yd, aux - e)) / (2 * h)
        for got, want in zip(d, num):
            assert rel(got, want) < 2e-6  # finite-difference accuracy; the reference asserts 2e-5 absolute


@pytest.mark.parametrize("robot", ["mini_cheetah", "tello_with_arms"])
def test_oracle_derivative_identities(oracle, robot):
    """d FD / d tau = H^-1 and the implicit function theorem d FD / d x = -H^-1 d ID / d x at ydd = FD."""
    o = oracle.OracleModel(robot)
    q, yd, tau = o.generate_states(4, seed=5)[:3]
    dq, dv, dt = o.dynamics_derivatives(q, yd, tau, True)
    H = o.mass_matrix(q)
    assert rel(dt, np.linalg.inv(H)) < 1e-10
    ydd = o.forward_dynamics(q, yd, tau)
    idq, idv = o.dynamics_derivatives(q, yd, ydd, False)
    assert rel(dq, -np.linalg.solve(H, idq)) < 1e-9 and rel(dv, -np.linalg.solve(H, idv)) < 1e-9


@pytest.mark.parametrize("robot", ALL_ROBOTS)
def test_derivative_programs_match_oracle(grbda, oracle, robot, tmp_path):
    """The emitted derivative programs, replayed in numpy, against the oracle's dual-number Jacobians."""
    m = grbda.ClusterTreeModel.from_robot(robot, device=None)
    o = oracle.OracleModel(robot)
    q, yd, aux = o.generate_states(8, seed=3)[:3]
    for algo, forward in ((grbda.ALGO_ID_DERIV, False), (grbda.ALGO_FD_DERIV, True)):
        path = str(tmp_path / ("d%d.tape" % algo))
        counts = m.dump_program(algo, path)
        outs = run_tape(load_tape(path), [q, yd, aux])
        want = o.dynamics_derivatives(q, yd, aux, forward)
        assert len(outs) == len(want)
        for got, w in zip(outs, want):
            assert rel(got.reshape(w.shape).transpose(0, 2, 1), w) < TOL  # the programs write column-major matrices
        # sparsity is exploited: far fewer operations than 2 nv dense tangent sweeps of the program
        base = m.dump_program(grbda.ALGO_ID)["flops"]
        if not forward and o.nv >= 18:
            assert counts["flops"] < 0.35 * 2 * o.nv * 2.5 * base


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["revolute_chain_with_rotor_4", "mini_cheetah", "mit_humanoid", "tello_with_arms"])
def test_derivatives_on_gpu(grbda, oracle, robot):
    """grbda_cuda_{inverse,forward}_dynamics_derivatives_f64 against the oracle (run-time compiled programs)."""
    import torch
    m = grbda.ClusterTreeModel.from_robot(robot, device=0)
    o = oracle.OracleModel(robot)
    B = 300  # ragged: two full tiles and a tail
    q, yd, aux, flags = m.generateStates(B)
    assert int(flags.sum()) == 0
    qn, ydn, auxn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    got = m.inverseDynamicsDerivatives(q, yd, aux)
    want = o.dynamics_derivatives(qn, ydn, auxn, False)
    for g, w in zip(got, want):
        assert rel(g.cpu().numpy(), w) < TOL
    got = m.forwardDynamicsDerivatives(q, yd, aux)
    want = o.dynamics_derivatives(qn, ydn, auxn, True)
    for g, w in zip(got, want):
        assert rel(g.cpu().numpy(), w) < 1e-9
    # d tau / d ydd is the mass matrix; d ydd / d tau its inverse
    H = m.getMassMatrix(q)
    eye = torch.eye(m.nv, dtype=torch.float64, device=q.device).expand(B, -1, -1)
    assert float((torch.bmm(H, got[2]) - eye).abs().max()) < 1e-9
