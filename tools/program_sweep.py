"""Forward dynamics program choice per model: the articulated-body sweep (O(depth)) against H^-1 (tau - C)
(cluster CRBA + bias + branch-sparse L^T D L), both compiled at run time (GRBDA_JIT=force) and timed on the
same states, next to the operation counts of the two programs. Calibrates the rule runtime/jit.cpp uses and
the ahead-of-time table of build.py. Usage: python tools/program_sweep.py [log2 batch] > profiles/r2_fd_program_sweep.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GRBDA_JIT"] = "force"
import generalized_rbda_b200 as grbda  # noqa: E402

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 18
CORPUS = os.path.join(ROOT, "tests", "urdf_corpus")


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


models = [("robot", r) for r in ("tello_with_arms", "mit_humanoid", "mini_cheetah", "jvrc1_humanoid")]
models += [("robot", "revolute_chain_with_rotor_%d" % d) for d in (2, 4, 6, 8, 10, 12, 16, 20, 24)]
models += [("urdf", os.path.join(CORPUS, f)) for f in ("explicit_parallel_chains_depth20_loop_size6.urdf",
                                                       "explicit_parallel_chains_depth40_loop_size8.urdf")]
for kind, name in models:
    rec = {"model": os.path.basename(name), "states": 1 << LOG2}
    for prog in ("ltl", "aba"):
        os.environ["GRBDA_FD_PROGRAM"] = prog
        m = grbda.ClusterTreeModel.from_robot(name) if kind == "robot" else grbda.ClusterTreeModel.from_urdf(name)
        B = 1 << LOG2
        q, yd, tau, _ = m.generateStates(B)
        out = torch.empty_like(tau)
        m.forwardDynamics(q, yd, tau, out=out)
        info = m.kernel_info(grbda.ALGO_FD)
        rec.update({"nv": m.nv, "bodies": m.nb, "clusters": m.nc})
        rec[prog] = {"fd_ms": round(timeit(lambda: m.forwardDynamics(q, yd, tau, out=out)), 4),
                     "flops": m.kernel_counts(grbda.ALGO_FD)["flops"], "block": info["block"],
                     "ctas_per_sm": info["min_blocks"], "parked": info["parked"], "nvrtc_ms": info["compile_ms"]}
        if prog == "ltl":
            rec["id_ms"] = round(timeit(lambda: m.inverseDynamics(q, yd, tau, out=out)), 4)
        del m, q, yd, tau, out
        torch.cuda.empty_cache()
    rec["faster"] = "aba" if rec["aba"]["fd_ms"] < rec["ltl"]["fd_ms"] else "ltl"
    rec["flops_ratio_ltl_over_aba"] = round(rec["ltl"]["flops"] / rec["aba"]["flops"], 2)
    print(json.dumps(rec), flush=True)
