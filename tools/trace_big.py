"""Timeline of the H = 512 forward scan (dev tool).  Stamps per (step, sub-tile): MMA warp 0 top / 1 h arrived /
2 MMAs committed; cell warp 0: 3 mma_done seen / 4 first row group staged; exchange warp 0: 5 row group handed over /
6 store performed / 7 multicast load issued (cell warp 0 does the exchange of row group 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-r2d2-dpg_b200")]
import numpy as np, torch
from r2d2_b200 import native as nv
lib = nv.lib()
H, B, S = 512, int(os.environ.get("B", 512)), 125
gin = torch.randn(S, B, 4 * H, device="cuda") * 0.5
whh = (torch.rand(4 * H, H, device="cuda") * 2 - 1) / np.sqrt(4 * H)
gates = torch.empty_like(gin); hs = torch.empty(S + 1, B, H, device="cuda"); cs = torch.empty_like(hs)
grid = 16 * 16
trace = torch.zeros(grid, S, 2, 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    nv.check(lib.r2d2_debug_scan_forward_trace(nv.dptr(gin), nv.dptr(whh), nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), S, B, H,
                                               nv.dptr(trace, torch.int64), nv.current_stream()))
torch.cuda.synchronize()
for cta in (0, 5, 16 * 3 + 7):
    t = trace.cpu().numpy().astype(np.float64)[cta]      # [S][2][8]
    k = slice(20, S - 3)
    print(f"--- CTA {cta}: step period {float(np.mean(np.diff(t[k, 0, 0]))):.0f} ns")
    for sub in (0, 1):
        d = lambda a, b: float(np.mean(t[k, sub, b] - t[k, sub, a]))
        print(f" sub {sub}: MMA wait h {d(0,1):6.0f} | issue {d(1,2):6.0f} | commit->cells see done {d(2,3):6.0f} | ld+act+cells {d(3,4):6.0f} |"
              f" fence+bar+store+wait {d(4,6):6.0f} | buf_free+load issue {d(6,7):6.0f} |"
              f" load issued -> h arrived (next step) {float(np.mean(t[21:S-2, sub, 1] - t[20:S-3, sub, 7])):6.0f} ns")
