"""Launch-shape sweep of run-time compiled kernels for the large models (GRBDA_JIT_SHAPE): bulk-copy staged tiles
at the CTA sizes that fit against direct global I/O. Usage: python tools/shape_sweep.py [log2 batch]"""
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


models = [("robot", "jvrc1_humanoid"), ("urdf", os.path.join(CORPUS, "explicit_parallel_chains_depth20_loop_size6.urdf")),
          ("urdf", os.path.join(CORPUS, "explicit_parallel_chains_depth40_loop_size8.urdf")), ("robot", "tello_with_arms")]
for kind, name in models:
    for shape in (None, "D", "T,64,2", "T,128,1", "T,64,1", "T,32,2", "T,32,4"):
        if shape is None:
            os.environ.pop("GRBDA_JIT_SHAPE", None)
        else:
            os.environ["GRBDA_JIT_SHAPE"] = shape
        rec = {"model": os.path.basename(name), "shape": shape or "auto", "states": 1 << LOG2}
        try:
            m = grbda.ClusterTreeModel.from_robot(name) if kind == "robot" else grbda.ClusterTreeModel.from_urdf(name)
            q, yd, tau, _ = m.generateStates(1 << LOG2)
            out = torch.empty_like(tau)
            m.forwardDynamics(q, yd, tau, out=out)
            m.inverseDynamics(q, yd, tau, out=out)
            i = m.kernel_info(grbda.ALGO_FD)
            rec.update({"fd_ms": round(timeit(lambda: m.forwardDynamics(q, yd, tau, out=out)), 4),
                        "id_ms": round(timeit(lambda: m.inverseDynamics(q, yd, tau, out=out)), 4),
                        "block": i["block"], "ctas": i["min_blocks"], "smem": i["smem"], "direct": i["direct"], "parked": i["parked"]})
            del m, q, yd, tau, out
        except Exception as e:  # a shape that does not fit
            rec["error"] = str(e)[:120]
        torch.cuda.empty_cache()
        print(json.dumps(rec), flush=True)
