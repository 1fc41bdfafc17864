"""Summarise an .ncu-rep (read with `ncu -i`): per kernel the figures DESIGN.md / profiles/ quote.
Usage: python tools/ncu_summary.py report.ncu-rep [--json]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__stack_size",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.avg.peak_sustained", "sm__cycles_elapsed.max"]


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for w in WANT:
            if w in hdr:
                d[w] = (r[hdr.index(w)] + " " + units[hdr.index(w)]).strip()
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        d["stall_samples_pct"] = {h: round(100 * v / tot, 1) for v, h in sorted(stalls, reverse=True)[:8]}
        out.append(d)
    return out


if __name__ == "__main__":
    res = summarise(sys.argv[1])
    if "--json" in sys.argv:
        print(json.dumps(res, indent=1))
    else:
        for d in res:
            print("---")
            for k, v in d.items():
                print("%-66s %s" % (k, v))
