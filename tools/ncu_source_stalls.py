"""Aggregate the source page of an ncu capture (ncu -i rep --page source --csv > file): stall samples per
SASS opcode and per stall reason, for each kernel. Usage: python tools/ncu_source_stalls.py source.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for k in kernels:
    h = k["hdr"]
    ix = {n: i for i, n in enumerate(h)}
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot, byop, cnt, samples = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    for r in k["rows"]:
        src = r[ix["Source"]].strip()
        parts = src.split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        cnt[op] += 1
        for s in stall_cols:
            v = int(r[ix[s]] or 0)
            tot[s] += v
            byop[op][s] += v
            samples[op] += v
    T = sum(tot.values()) or 1
    print(k["name"][:100], "| SASS instructions", len(k["rows"]), "| samples", T)
    print("  stall totals %:", {s[6:]: round(100 * v / T, 1) for s, v in tot.most_common(9)})
    print("  opcode histogram:", dict(cnt.most_common(14)))
    for op, v in samples.most_common(12):
        top = ", ".join("%s %.0f%%" % (s[6:], 100 * x / v) for s, x in byop[op].most_common(4))
        print("  %-8s n=%5d  samples %5.1f%%  per-instr %6.2f | %s" % (op, cnt[op], 100 * v / T, v / cnt[op], top))
