"""Repeat the FD / ID kernel timing several times in fresh allocations to look at run-to-run spread."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda
m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
B = 1 << 20
res = []
keep = []
for trial in range(3):
    q, yd, tau, _ = m.generateStates(B, first_index=trial * B)
    out = torch.empty_like(tau)
    def t(fn, reps=10):
        for _ in range(3): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    row = []
    for variant in range(3):
        os.environ["GRBDA_KERNEL_VARIANT"] = str(variant)
        row.append(round(t(lambda: m.forwardDynamics(q, yd, tau, out=out)), 3))
    os.environ["GRBDA_KERNEL_VARIANT"] = "0"
    row.append(round(t(lambda: m.inverseDynamics(q, yd, tau, out=out)), 3))
    res.append(row)
    keep.append(torch.empty(int(1e8 * (trial + 1)), dtype=torch.uint8, device="cuda"))  # perturb the allocator
print(json.dumps({"device": torch.cuda.get_device_name(0), "visible": os.environ.get("CUDA_VISIBLE_DEVICES"), "fd_v0_v1_v2_id_ms": res}))
