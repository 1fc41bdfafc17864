"""End-to-end (pinned host buffers -> results in host memory) throughput of the fused forward+inverse
dynamics call for several pipeline chunk sizes (GRBDA_HOST_CHUNK). Usage: python tools/e2e_sweep.py"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import generalized_rbda_b200 as grbda
m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
B = 1 << 20
q, yd, tau, _ = m.generateStates(B)
qh, ydh, tauh = (x.cpu().pin_memory() for x in (q, yd, tau))
yddh = torch.empty((B, m.nv), dtype=torch.float64).pin_memory()
tbh = torch.empty((B, m.nv), dtype=torch.float64).pin_memory()
for chunk in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18):
    os.environ["GRBDA_HOST_CHUNK"] = str(chunk)
    m.forward_inverse_host(qh, ydh, tauh, yddh, tbh)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        m.forward_inverse_host(qh, ydh, tauh, yddh, tbh)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps({"chunk": chunk, "ms": dt * 1e3, "pairs_per_s": B / dt, "h2d_GBs": B * (m.nq + 2 * m.nv) * 8 / dt / 1e9,
                      "d2h_GBs": 2 * B * m.nv * 8 / dt / 1e9}))
