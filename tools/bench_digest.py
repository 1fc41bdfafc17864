"""Short digest of a bench.py JSON line. Usage: python tools/bench_digest.py line.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
print("value %.4g %s | ms/step %.4f | FD %.4f ms (frac %.3f) | ID %.4f ms (frac %.3f)" % (
    d["value"], d["scaling"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["inverse_dynamics"]["kernel_ms"],
    r["inverse_dynamics"]["frac"]))
print("weak %s | strong %s" % (d.get("weak_scaling"), d.get("strong_scaling")))
if "e2e" in d:
    print("e2e %.4g | host placement %s" % (d["e2e"]["value"], d["e2e"].get("host_placement")))
for k, v in d.get("other_kernels", {}).items():
    print(" ", k, json.dumps(v)[:400])
