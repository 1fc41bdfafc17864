"""Small run of every entry point for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda
for robot in ("tello_with_arms", "four_bar", "jvrc1_humanoid"):
    m = grbda.ClusterTreeModel.from_robot(robot)
    for B in (1, 37, 300):
        q, yd, tau, _ = m.generateStates(B, seed=3)
        for dt in (torch.float64, torch.float32):
            qq, yy, tt = q.to(dt), yd.to(dt), tau.to(dt)
            try:
                m.forwardDynamics(qq, yy, tt); m.inverseDynamics(qq, yy, tt); m.getMassMatrix(qq); m.forwardKinematics(qq, yy)
            except grbda.GrbdaError:
                assert dt == torch.float32  # FP32 kernels are only compiled for some models
        f = torch.ones((B, len(m.externalForceBodies()), 6), dtype=torch.float64, device="cuda")
        m.forwardDynamics(q, yd, tau, f_ext=f); m.inverseDynamics(q, yd, tau, f_ext=f)
        q2 = q.clone(); q2[0, m.clusters()[-1]["position_index"]] = 7.0e13
        m.forwardDynamics(q2, yd, tau); m.getMassMatrix(q2); m.forwardKinematics(q2, yd)
        m.constraintViolation(q)
    torch.cuda.synchronize()
print("sanitize_run: done")
