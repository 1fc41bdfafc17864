"""Kernel times of the derivative programs and, beside them, the finite-difference route they replace
(2 nv + 1 evaluations of the plain kernel per state). Usage: python tools/deriv_timing.py [log2 batch] > profiles/r2_derivatives.jsonl"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda  # noqa: E402

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for robot in ("revolute_chain_with_rotor_8", "mini_cheetah", "mit_humanoid", "tello_with_arms"):
    m = grbda.ClusterTreeModel.from_robot(robot)
    B = 1 << LOG2
    q, yd, aux, _ = m.generateStates(B)
    out = torch.empty_like(aux)
    rec = {"model": robot, "nv": m.nv, "states": B}
    t0 = time.time()
    m.inverseDynamicsDerivatives(q[:128], yd[:128], aux[:128])
    torch.cuda.synchronize()
    rec["id_deriv_first_call_s"] = round(time.time() - t0, 1)  # NVRTC (or the disk cache)
    t0 = time.time()
    m.forwardDynamicsDerivatives(q[:128], yd[:128], aux[:128])
    torch.cuda.synchronize()
    rec["fd_deriv_first_call_s"] = round(time.time() - t0, 1)
    rec["id_ms"] = round(timeit(lambda: m.inverseDynamics(q, yd, aux, out=out)), 4)
    rec["fd_ms"] = round(timeit(lambda: m.forwardDynamics(q, yd, aux, out=out)), 4)
    rec["id_deriv_ms"] = round(timeit(lambda: m.inverseDynamicsDerivatives(q, yd, aux)), 4)
    rec["fd_deriv_ms"] = round(timeit(lambda: m.forwardDynamicsDerivatives(q, yd, aux)), 4)
    rec["id_deriv_flops"] = m.dump_program(grbda.ALGO_ID_DERIV)["flops"]
    rec["fd_deriv_flops"] = m.dump_program(grbda.ALGO_FD_DERIV)["flops"]
    # central finite differences need 2 * (2 nv) evaluations (dq and yd); one-sided 2 nv + 1
    rec["id_deriv_vs_central_differences"] = round(4 * m.nv * rec["id_ms"] / rec["id_deriv_ms"], 1)
    rec["fd_deriv_vs_central_differences"] = round(6 * m.nv * rec["fd_ms"] / rec["fd_deriv_ms"], 1)
    print(json.dumps(rec), flush=True)
