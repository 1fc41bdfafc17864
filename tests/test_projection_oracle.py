"""Second oracle: the spanning-tree / projection formulation the reference uses as its own cross check
(RigidBodyTreeModel, src/Dynamics/RigidBodyTreeDynamics.cpp:86-97,137-146; reference tests
UnitTests/testRigidBodyDynamicsAlgos.cpp:113-239 compare the cluster algorithms with it at 1e-8 .. 1e-6), restated in
numpy (oracle/projection.py). Three independent implementations must agree: cluster oracle (dense cluster recursion),
projection oracle (body-level recursion + G^T H G), and the product's emitted programs."""
import numpy as np
import pytest

from tape import load_tape, run_tape

ROBOTS = ["tello_with_arms", "mit_humanoid", "mini_cheetah", "revolute_pair_chain_with_rotor_4", "four_bar",
          "revolute_triple_chain_with_rotor_6"]


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


@pytest.mark.parametrize("robot", ROBOTS)
def test_cluster_algorithms_against_projection(grbda, oracle, robot, tmp_path):
    from mirror import mirror_to_oracle
    from oracle.projection import from_models
    m = grbda.ClusterTreeModel.from_robot(robot, device=None)
    try:
        o = oracle.OracleModel(robot)
    except Exception:
        o = mirror_to_oracle(m, oracle)  # URDF-only models
    st = from_models(m, o)
    assert st.nv == o.nv and st.nb == o.nb
    q, yd, aux = o.generate_states(6, seed=29)[:3]
    # reference test tolerances: ID 1e-8 (testRigidBodyDynamicsAlgos.cpp:190), FD 1e-6 / 5e-6 (:150,233), H 1e-8 (:122)
    tau_p, tau_c = st.inverse_dynamics(q, yd, aux), o.inverse_dynamics(q, yd, aux)
    assert rel(tau_p, tau_c) < 1e-10
    H_p, H_c = st.mass_matrix(q), o.mass_matrix(q)
    assert rel(H_p, H_c) < 1e-10
    ydd_p, ydd_c = st.forward_dynamics(q, yd, aux), o.forward_dynamics(q, yd, aux)
    assert rel(ydd_p, ydd_c) < 1e-8
    # and the product's programs against the projection oracle directly
    ins = [q, yd, aux]
    for algo, want, tol in ((grbda.ALGO_ID, tau_p, 1e-10), (grbda.ALGO_FD, ydd_p, 1e-8), (grbda.PROGRAM_FD_LTL, ydd_p, 1e-8),
                            (grbda.ALGO_H, H_p.reshape(q.shape[0], -1), 1e-10)):
        path = str(tmp_path / ("p%d.tape" % algo))
        m.dump_program(algo, path)
        assert rel(run_tape(load_tape(path), ins)[0], want) < tol
