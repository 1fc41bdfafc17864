"""A/B timing of library builds in ONE GPU session (box-to-box differences of 5-10 % make numbers from different
sessions incomparable). Every build is loaded in its own subprocess (GRBDA_LIB_PATH), rounds alternate.
Usage: python tools/ab.py [--model M] [--log2 B] [--rounds R] [--algos fd,id,fk,h] label=path/to/libgrbda_cuda.so ...
       (label=product is the product library)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
import generalized_rbda_b200 as grbda
model, B, algos = sys.argv[1], 1 << int(sys.argv[2]), sys.argv[3].split(",")
m = grbda.ClusterTreeModel.from_robot(model)
q, yd, tau, _ = m.generateStates(B)
out = torch.empty_like(tau)
Bk = max(1, B // 4)
def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
r = {}
if "fd" in algos: r["fd"] = timeit(lambda: m.forwardDynamics(q, yd, tau, out=out))
if "id" in algos: r["id"] = timeit(lambda: m.inverseDynamics(q, yd, tau, out=out))
if "h" in algos:
    H = torch.empty((Bk, m.nv, m.nv), dtype=torch.float64, device="cuda")
    r["h"] = timeit(lambda: m.getMassMatrix(q[:Bk], out=H), 10)
if "fk" in algos:
    o = m.forwardKinematics(q[:Bk], yd[:Bk])
    r["fk"] = timeit(lambda: m.forwardKinematics(q[:Bk], yd[:Bk], out=o), 10)
print(json.dumps(r))
''' % ROOT

args = sys.argv[1:]
model, log2, rounds, algos = "tello_with_arms", 20, 3, "fd,id"
libs = []
while args:
    a = args.pop(0)
    if a == "--model":
        model = args.pop(0)
    elif a == "--log2":
        log2 = int(args.pop(0))
    elif a == "--rounds":
        rounds = int(args.pop(0))
    elif a == "--algos":
        algos = args.pop(0)
    else:
        label, path = a.split("=", 1)
        libs.append((label, path))
results = {label: [] for label, _ in libs}
for rnd in range(rounds):
    for label, path in libs:
        env = dict(os.environ)
        if path != "product":
            env["GRBDA_LIB_PATH"] = os.path.abspath(path)
        r = subprocess.run([sys.executable, "-c", WORKER, model, str(log2), algos], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(label, "FAILED", r.stderr[-800:])
            continue
        results[label].append(json.loads(r.stdout.strip().splitlines()[-1]))
print("model %s, 2^%d states (h / fk on a quarter), ms, best of %d rounds [all rounds]" % (model, log2, rounds))
for label, rs in results.items():
    if rs:
        print("%-24s" % label, "  ".join("%s %.4f %s" % (k, min(r[k] for r in rs), [round(r[k], 4) for r in rs]) for k in rs[0]))
print(json.dumps(results))
