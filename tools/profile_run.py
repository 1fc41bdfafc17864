"""Tiny driver for ncu captures: one launch of each hot kernel on generated states.
Usage: python tools/profile_run.py [model] [log2 batch] [algos]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "tello_with_arms"
B = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 18)
algos = (sys.argv[3] if len(sys.argv) > 3 else "fd,id").split(",")
m = grbda.ClusterTreeModel.from_robot(model)
q, yd, tau, _ = m.generateStates(B)
out = torch.empty_like(tau)
for _ in range(2):
    if "fd" in algos:
        m.forwardDynamics(q, yd, tau, out=out)
    if "id" in algos:
        m.inverseDynamics(q, yd, tau, out=out)
    if "h" in algos:
        m.getMassMatrix(q)
    if "fk" in algos:
        m.forwardKinematics(q, yd)
torch.cuda.synchronize()
