"""Digest of the four entry points' outputs on generated states (ragged batch sizes included): two library builds
that run the same programs must print the same lines. Usage: [GRBDA_LIB_PATH=...] python tools/digest_run.py [model]"""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda

model = sys.argv[1] if len(sys.argv) > 1 else "tello_with_arms"
m = grbda.ClusterTreeModel.from_robot(model)
for B in (1, 37, 128, 65536 + 37):
    q, yd, tau, _ = m.generateStates(B)
    a = m.forwardDynamics(q, yd, tau)
    b = m.inverseDynamics(q, yd, tau)
    fk = m.forwardKinematics(q, yd)
    H = m.getMassMatrix(q)
    h = hashlib.sha256()
    for x in (a, b, H) + tuple(fk):
        h.update(x.cpu().numpy().tobytes())
    print(model, B, h.hexdigest()[:16])
