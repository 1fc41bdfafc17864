"""SASS opcode histogram of the default kernels of one model (evidence for profiles/: DFMA / DMUL / DADD mix,
spill instructions LDL / STL, bulk-copy shell UBLKCP + SYNCS, 256-bit stores), read with cuobjdump from the
objects build.py left in generalized_rbda_b200/_build.
Usage: python tools/sass_histogram.py [model] > profiles/r2_sass_histogram_<model>.json"""
import collections
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
model = sys.argv[1] if len(sys.argv) > 1 else "tello_with_arms"
KEEP = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "LDL", "STL", "LDS", "STS", "LDG", "STG", "UBLKCP", "SYNCS", "LDC", "LDCU",
        "MOV", "IMAD.MOV", "BRA", "BAR", "WARPSYNC", "SHFL", "LOP3", "SEL", "FSEL", "I2F", "F2I", "DSEL")
out = {"model": model, "kernels": []}
for algo in ("id", "fd", "fk", "h"):
    objs = sorted(glob.glob(os.path.join(ROOT, "generalized_rbda_b200", "_build", "%s_%s.cu.*.o" % (model, algo))),
                  key=os.path.getmtime)
    if not objs:
        continue
    sass = subprocess.run(["cuobjdump", "-sass", objs[-1]], capture_output=True, text=True).stdout
    fn, hist = None, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn:
                out["kernels"].append({"algo": algo, "function": fn, **hist})
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:160]
            hist = {"instructions": 0, "opcodes": collections.Counter()}
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and hist is not None:
            op = m.group(1)
            hist["instructions"] += 1
            key = next((k for k in sorted(KEEP, key=len, reverse=True) if op == k or op.startswith(k + ".")), None)
            if op.startswith("STG") and ".256" in op:
                key = "STG.256"
            hist["opcodes"][key or "other"] += 1
    if fn:
        out["kernels"].append({"algo": algo, "function": fn, **hist})
for k in out["kernels"]:
    k["opcodes"] = dict(sorted(k["opcodes"].items(), key=lambda kv: -kv[1]))
    k["code_bytes"] = 16 * k["instructions"]
print(json.dumps(out, indent=1))
