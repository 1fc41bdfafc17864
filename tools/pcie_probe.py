"""Host-link ceiling of the end-to-end path, one process per GPU (torchrun): pinned-memory copy bandwidth per GPU,
(a) one rank at a time, (b) all ranks at once, H2D alone, D2H alone and both directions together, next to the
end-to-end rate of grbda_cuda_forward_inverse_host_f64 measured the same two ways. If (b) collapses for plain
copies exactly as the end-to-end rate does, the limit is the host side of the box (PCIe root / IOMMU / memory of
the VM), not this library. Usage: python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import generalized_rbda_b200 as grbda  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = grbda.bind_host_to_device(local)

GB = 1 << 30
h_in = torch.empty(GB // 8, dtype=torch.float64).pin_memory()
h_out = torch.empty(GB // 8, dtype=torch.float64).pin_memory()
d_in = torch.empty(GB // 8, dtype=torch.float64, device=dev)
d_out = torch.ones(GB // 8, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def copy_rate(mode, reps=4):
    """GB/s of this rank for `mode` in {"h2d", "d2h", "both"}"""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return (2 if mode == "both" else 1) * reps * GB / dt / 1e9


m = grbda.ClusterTreeModel.from_robot("tello_with_arms", device=local)
B = 1 << 20
q, yd, tau, _ = m.generateStates(B, first_index=rank * B)
qh, ydh, tauh = (x.cpu().pin_memory() for x in (q, yd, tau))
yddh = torch.empty((B, m.nv), dtype=torch.float64).pin_memory()
tbh = torch.empty((B, m.nv), dtype=torch.float64).pin_memory()


def e2e_rate(reps=3):
    m.forward_inverse_host(qh, ydh, tauh, yddh, tbh)
    t0 = time.perf_counter()
    for _ in range(reps):
        m.forward_inverse_host(qh, ydh, tauh, yddh, tbh)
    return reps * B / (time.perf_counter() - t0)


res = {"rank": rank, "world": world, "host_placement": numa}
copy_rate("both", 1)
# (a) one rank at a time
for r in range(world):
    barrier()
    if r == rank:
        res["alone"] = {k: round(copy_rate(k), 1) for k in ("h2d", "d2h", "both")}
        res["alone"]["e2e_Mpairs_s"] = round(e2e_rate() / 1e6, 1)
    barrier()
# (b) all ranks at once
barrier()
res["together"] = {}
for k in ("h2d", "d2h", "both"):
    barrier()
    res["together"][k] = round(copy_rate(k), 1)
barrier()
res["together"]["e2e_Mpairs_s"] = round(e2e_rate() / 1e6, 1)
barrier()
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    for r in out:
        print(json.dumps(r))
    tot = lambda sect, k: round(sum(r[sect][k] for r in out), 1)
    print(json.dumps({"sum_over_ranks": {s: {k: tot(s, k) for k in ("h2d", "d2h", "both", "e2e_Mpairs_s")}
                                         for s in ("alone", "together")},
                      "note": "alone = each rank measured while the others idle (sum = what perfect scaling would give)"}))
if world > 1:
    dist.destroy_process_group()
