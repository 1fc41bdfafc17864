"""Per-model, per-algorithm kernel times of the default kernels (FP64), SURVEY 8(d) "per-config numbers".
Usage: python tools/per_model.py [log2 batch] > profiles/r1_per_model.jsonl"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
HBM = 6555.8e9


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


peak = grbda.measure_fma_peak(0, False, 0.5)
for robot in ("tello_with_arms", "tello", "mit_humanoid", "mini_cheetah", "jvrc1_humanoid", "mit_humanoid_leg",
              "revolute_rotor_chain", "revolute_chain_with_rotor_2", "revolute_chain_with_rotor_4",
              "revolute_chain_with_rotor_8", "revolute_chain_with_rotor_16", "four_bar", "six_bar",
              "planar_leg_linkage"):
    m = grbda.ClusterTreeModel.from_robot(robot)
    B = 1 << LOG2
    Bk = B >> 2 if m.nv > 12 else B
    q, yd, tau, _ = m.generateStates(B)
    out = torch.empty_like(tau)
    H = torch.empty((Bk, m.nv, m.nv), dtype=torch.float64, device="cuda")
    fk = m.forwardKinematics(q[:Bk], yd[:Bk])
    t = {"id": timeit(lambda: m.inverseDynamics(q, yd, tau, out=out)),
         "fd": timeit(lambda: m.forwardDynamics(q, yd, tau, out=out)),
         "h": timeit(lambda: m.getMassMatrix(q[:Bk], out=H), 5) * (B / Bk),
         "fk": timeit(lambda: m.forwardKinematics(q[:Bk], yd[:Bk], out=fk), 5) * (B / Bk)}
    bytes_ = {"id": 8 * (m.nq + 3 * m.nv), "fd": 8 * (m.nq + 3 * m.nv), "fk": 8 * (m.nq + m.nv + 18 * m.nb),
              "h": 8 * (m.nq + m.nv * m.nv)}
    rec = {"model": robot, "nq": m.nq, "nv": m.nv, "bodies": m.nb, "clusters": m.nc, "states": B}
    for a, idx in (("id", 0), ("fd", 1), ("fk", 2), ("h", 3)):
        flops = m.kernel_counts(idx)["flops"]
        rec[a] = {"ms_per_2^%d" % LOG2: round(t[a] * 1e3, 4), "evals_per_s": round(B / t[a]),
                  "executed_flops": flops, "fp64_frac": round(flops * B / t[a] / peak, 3),
                  "hbm_frac": round(bytes_[a] * B / t[a] / HBM, 3)}
    print(json.dumps(rec), flush=True)
    del q, yd, tau, out, H, fk
    torch.cuda.empty_cache()
