"""Copy one round's measurement artefacts from gpurun_out/ into profiles/ and print the numbers the
README tables quote. Usage: python tools/refresh_profiles.py <bench.json> <launches.csv> <fd_id.ncu-rep> [round prefix, default r2]"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bench, launches, rep = sys.argv[1:4]
RND = sys.argv[4] if len(sys.argv) > 4 else "r2"
P = os.path.join(ROOT, "profiles")
shutil.copy(launches, os.path.join(P, RND + "_launches_bench_steps2.csv"))
open(os.path.join(P, RND + "_bench_1gpu.json"), "w").write([l for l in open(bench) if l.startswith("{")][0])
summ = os.path.join(ROOT, "tools", "ncu_summary.py")
js = subprocess.run([sys.executable, summ, rep, "--json"], stdout=subprocess.PIPE, text=True).stdout
open(os.path.join(P, RND + "_ncu_full_fd_id.json"), "w").write(js)
open(os.path.join(P, RND + "_ncu_full_fd_id.txt"), "w").write(
    subprocess.run([sys.executable, summ, rep], stdout=subprocess.PIPE, text=True).stdout)
d = json.loads(js)
def b(x):
    v, u = x.split()
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
out = {}
for k, name in zip(d, ["forward_dynamics", "inverse_dynamics"]):
    out[name] = {"kernel": k["kernel"], "states": 1 << 20,
                 "dram_bytes_per_launch": b(k["dram__bytes_read.sum"]) + b(k["dram__bytes_write.sum"]),
                 "dram_read": b(k["dram__bytes_read.sum"]), "dram_write": b(k["dram__bytes_write.sum"]),
                 "duration_us_under_ncu": float(k["gpu__time_duration.sum"].split()[0]),
                 "fp64_pipe_active_pct": float(k["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"].split()[0]),
                 "source": "profiles/" + RND + "_ncu_full_fd_id.json (ncu --set full --clock-control none, tools/profile_run.py tello_with_arms 20 fd,id)"}
    print(k["kernel"][:72])
    for a in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__sass_inst_executed_op_local_ld.sum", "lts__t_sector_hit_rate.pct",
              "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed",
              "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed",
              "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed", "stall_samples_pct"]:
        print("   ", a, k.get(a))
json.dump(out, open(os.path.join(P, RND + "_roofline_traffic.json"), "w"), indent=1)
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) / 1e3
    if "grbda_batched" in r[ki] and (v > 250 or ", 0>" in r[ki]):
        agg[r[ki][20:92]].append(v)
for k, v in agg.items():
    print(k, len(v), round(min(v), 1), round(max(v), 1))
bj = json.load(open(os.path.join(P, RND + "_bench_1gpu.json"))); r = bj["roofline"]
print({k: r[k] for k in ["achieved", "peak", "frac", "kernel_ms", "traffic", "flops_per_state_executed"]}, r["algorithmic"])
print(r["inverse_dynamics"]); print(bj["other_kernels"]); print(bj["value"], bj["ms_per_step"], bj["e2e"]["value"], bj["cpu_baseline"]["value"])
