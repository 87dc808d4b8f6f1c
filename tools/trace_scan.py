"""Per-step timeline of the tcgen05 forward scan (dev tool): where does a cell step spend its time?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-r2d2-dpg_b200")]
import numpy as np, torch
from r2d2_b200 import native as nv
lib = nv.lib()
H, B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 256, int(sys.argv[3]) if len(sys.argv) > 3 else 64
C = H // 32
NB = 16 if (B + 15) // 16 <= 15 * 8 // C else 32
grid = C * (((B + 15) // 16) if NB == 16 else min(15 * 8 // C, (B + 31) // 32 if B > 32 * (15 * 8 // C) else 15 * 8 // C))
gin = torch.randn(S, B, 4 * H, device="cuda") * 0.5
whh = (torch.rand(4 * H, H, device="cuda") * 2 - 1) / np.sqrt(4 * H)
gates = torch.empty_like(gin); hs = torch.empty(S + 1, B, H, device="cuda"); cs = torch.empty_like(hs)
trace = torch.zeros(grid, S, 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    nv.check(lib.r2d2_debug_scan_forward_trace(nv.dptr(gin), nv.dptr(whh), nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), S, B, H,
                                               nv.dptr(trace, torch.int64), nv.current_stream()))
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.float64)          # [grid][S][8] ns
names = ["top", "h_arrived", "mma_issued", "mma_done", "tmem_ld+sync", "-", "-", "cells+copies_issued"]
cl = t[:C]                                           # cluster 0
steps = slice(8, S - 1)
print(f"H={H} B={B} NB={NB} grid={grid}; mean ns per phase (cluster 0, rank 0), steps 8..{S-2}")
r0 = cl[0]
for a, b in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 7)):
    print(f"  {names[a]:>14s} -> {names[b]:<20s} {np.mean(r0[steps, b] - r0[steps, a]):8.1f} ns")
print(f"  step period {np.mean(np.diff(r0[steps, 0])):8.1f} ns")
# copy latency: my h_arrived(s+1) minus the latest copies_issued(s) among the ranks of the cluster
lat = cl[0][9:S-1, 1] - cl[:, 8:S-2, 7].max(axis=0)
print(f"  last copy issued (any rank) -> h_arrived at rank 0: {lat.mean():8.1f} ns (min {lat.min():.0f} max {lat.max():.0f})")
skew = cl[:, steps, 7].max(axis=0) - cl[:, steps, 7].min(axis=0)
print(f"  skew of copies_issued across the {C} ranks: {skew.mean():8.1f} ns")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"trace_scan_H{H}_B{B}.npy"), t[:C])
# ---- whole-kernel view: launch-to-first-step, per-cluster step period, end-to-end
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
nv.check(lib.r2d2_debug_scan_forward_trace(nv.dptr(gin), nv.dptr(whh), nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), S, B, H,
                                           nv.dptr(trace, torch.int64), nv.current_stream()))
ev1.record(); torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.float64)
first, last = t[:, 0, 0].min(), t[:, S - 1, 7].max()
print(f"kernel {ev0.elapsed_time(ev1) * 1e3:.1f} us by events; first top -> last stamp {(last - first) / 1e3:.1f} us")
print("spread of first 'top' across CTAs (us):", np.round((t[:, 0, 0].max() - first) / 1e3, 1))
per = np.diff(t[::C, :, 0], axis=1)          # rank 0 of every cluster
print("step period per cluster (ns): mean", np.round(per[:, 8:].mean(axis=1)), " first 8 steps mean", np.round(per[:, :8].mean()))
print("slowest steps of cluster 0:", np.round(np.sort(per[0])[-6:]))
