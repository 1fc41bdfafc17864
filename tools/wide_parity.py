"""Ad-hoc wide parity sweep: every compiled model, 2^15 states, all four entry points (and the external
force entry points) against the oracle; prints the largest per-state relative error of each."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import generalized_rbda_b200 as grbda
from oracle import binding as oracle
from mirror import mirror_to_oracle

def relrows(a, b, scale=None, floor=1e-3):
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    s = np.abs(b).max(1) if scale is None else scale
    return float((np.abs(a - b).max(1) / np.maximum(floor, s)).max())

B = 1 << 15
for robot in ("tello_with_arms", "tello", "mit_humanoid", "mini_cheetah", "jvrc1_humanoid", "mit_humanoid_leg",
              "revolute_rotor_chain", "revolute_chain_with_rotor_16", "revolute_pair_chain_with_rotor_4", "four_bar",
              "six_bar", "planar_leg_linkage"):
    m = grbda.ClusterTreeModel.from_robot(robot)
    try:
        o = oracle.OracleModel(robot)
    except Exception:
        o = mirror_to_oracle(m, oracle)
    q, yd, aux, flags = m.generateStates(B, seed=2024)
    qn, ydn, auxn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    tau_o = o.inverse_dynamics(qn, ydn, auxn)
    ydd_o = o.forward_dynamics(qn, ydn, auxn)
    res = {"model": robot, "invalid_states": int(flags.sum()),
           "id": relrows(m.inverseDynamics(q, yd, aux).cpu().numpy(), tau_o)}
    ydd = m.forwardDynamics(q, yd, aux)
    e = np.abs(ydd.cpu().numpy() - ydd_o).max(1) / np.maximum(1e-3, np.abs(ydd_o).max(1))
    res["fd_max"], res["fd_p999"], res["fd_median"] = float(e.max()), float(np.quantile(e, 0.999)), float(np.median(e))
    back = m.inverseDynamics(q, yd, ydd).cpu().numpy()
    res["id_of_fd_median"] = float(np.median(np.abs(back - auxn).max(1) / np.abs(auxn).max(1)))
    n = 2048
    res["h"] = relrows(m.getMassMatrix(q[:n]).cpu().numpy(), o.mass_matrix(qn[:n]))
    p, R, v = m.forwardKinematics(q[:n], yd[:n]); po, Ro, vo = o.forward_kinematics(qn[:n], ydn[:n])
    res["fk"] = max(relrows(p.cpu().numpy(), po), relrows(R.cpu().numpy(), Ro), relrows(v.cpu().numpy(), vo))
    print(json.dumps(res), flush=True)
