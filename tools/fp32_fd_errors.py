"""FP32 forward dynamics against the FP64 oracle on float-rounded states: per model the median / maximum relative
error of ydd and the same error in units of eps32 * cond(H) (what a backward-stable solve of H ydd = tau - C may
lose). Feeds the tolerance of tests/test_gpu_parity.py::test_dynamics_parity_f32.
Usage: python tools/fp32_fd_errors.py > profiles/r2_fp32_fd_errors.jsonl"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import generalized_rbda_b200 as grbda  # noqa: E402
from oracle import binding  # noqa: E402
from mirror import mirror_to_oracle  # noqa: E402

EPS = float(np.finfo(np.float32).eps)
for robot in ("tello_with_arms", "mini_cheetah", "mit_humanoid", "revolute_chain_with_rotor_2", "revolute_chain_with_rotor_16",
              "revolute_rotor_chain", "jvrc1_humanoid"):
    m = grbda.ClusterTreeModel.from_robot(robot)
    try:
        o = binding.OracleModel(robot)
    except Exception:
        o = mirror_to_oracle(m, binding)
    q, yd, aux, _ = m.generateStates(512, seed=4)
    q32, yd32, aux32 = q.float(), yd.float(), aux.float()
    qn, ydn, auxn = (x.double().cpu().numpy() for x in (q32, yd32, aux32))
    ydd32 = m.forwardDynamics(q32, yd32, aux32).double().cpu().numpy()
    ydd = o.forward_dynamics(qn, ydn, auxn)
    err = np.abs(ydd32 - ydd).max(1) / np.abs(ydd).max(1)
    cond = np.linalg.cond(o.mass_matrix(qn))
    scaled = err / (EPS * cond)
    print(json.dumps({"model": robot, "median_err": float(np.median(err)), "max_err": float(err.max()),
                      "median_cond_H": float(np.median(cond)), "max_cond_H": float(cond.max()),
                      "median_err_over_eps_cond": float(np.median(scaled)), "max_err_over_eps_cond": float(scaled.max())}), flush=True)
