"""Time every compiled launch-shape variant of the Tello kernels (GRBDA_KERNEL_VARIANT) on one GPU.
Usage: python tools/sweep_variants.py [model] [log2 batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "tello_with_arms"
B = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
m = grbda.ClusterTreeModel.from_robot(model)
q, yd, tau, _ = m.generateStates(B)
peak64 = grbda.measure_fma_peak(0, False, 0.5)
peak32 = grbda.measure_fma_peak(0, True, 0.5)
print(json.dumps({"fp64_fma_peak_tflops": peak64 / 1e12, "fp32_fma_peak_tflops": peak32 / 1e12}))
progs = {a: m.dump_program(i) for i, a in enumerate(grbda.ALGO_NAMES)}
print(json.dumps({"program_counts": progs}))


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dtype in (torch.float64, torch.float32):
    qq, yy, tt = q.to(dtype), yd.to(dtype), tau.to(dtype)
    out = torch.empty_like(tt)
    Hn = min(B, 1 << 18)
    Hout = torch.empty((Hn, m.nv, m.nv), dtype=dtype, device="cuda")
    for variant in range(4):
        os.environ["GRBDA_KERNEL_VARIANT"] = str(variant)
        try:
            r = {"dtype": str(dtype), "variant": variant, "batch": B,
                 "fd_ms": timeit(lambda: m.forwardDynamics(qq, yy, tt, out=out)),
                 "id_ms": timeit(lambda: m.inverseDynamics(qq, yy, tt, out=out)),
                 "h_ms_per_2^18": timeit(lambda: m.getMassMatrix(qq[:Hn], out=Hout), 5),
                 "fk_ms_per_2^18": timeit(lambda: m.forwardKinematics(qq[:Hn], yy[:Hn]), 5)}
            r["fd_Mstates_s"] = B / r["fd_ms"] / 1e3
            r["id_Mstates_s"] = B / r["id_ms"] / 1e3
            peak = peak64 if dtype == torch.float64 else peak32
            r["fd_frac_executed_flops"] = progs["fd"]["flops"] * B / (r["fd_ms"] * 1e-3) / peak
            r["id_frac_executed_flops"] = progs["id"]["flops"] * B / (r["id_ms"] * 1e-3) / peak
            print(json.dumps(r), flush=True)
        except grbda.GrbdaError as e:
            print(json.dumps({"dtype": str(dtype), "variant": variant, "error": str(e)}))
