"""Timeline of the ping-pong forward scan (dev tool).  Stamps per iteration k (sub-tile k%2 of step k/2):
MMA warp: 0 top, 1 h arrived, 2 MMAs issued+committed.  Cell warps: 5 top, 3 mma_done seen, 4 accumulator in smem, 7 end."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-r2d2-dpg_b200")]
import numpy as np, torch
from r2d2_b200 import native as nv
lib = nv.lib()
H, B, S = 256, 256, 125
C = H // 32
grid = C * 15
gin = torch.randn(S, B, 4 * H, device="cuda") * 0.5
whh = (torch.rand(4 * H, H, device="cuda") * 2 - 1) / np.sqrt(4 * H)
gates = torch.empty_like(gin); hs = torch.empty(S + 1, B, H, device="cuda"); cs = torch.empty_like(hs)
trace = torch.zeros(grid, 2 * S, 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    nv.check(lib.r2d2_debug_scan_forward_trace(nv.dptr(gin), nv.dptr(whh), nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), S, B, H,
                                               nv.dptr(trace, torch.int64), nv.current_stream()))
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.float64)[0]      # CTA 0: [2S][8]
k = slice(20, 2 * S - 2)
d = lambda a, b: float(np.mean(t[k, b] - t[k, a]))
print(f"MMA warp : wait h {d(0,1):7.1f}  issue+commit {d(1,2):7.1f}  iteration period {float(np.mean(np.diff(t[k,0]))):7.1f} ns")
print(f"cell warps: wait mma_done {d(5,3):7.1f}  ld+sync {d(3,4):7.1f}  cells+copies {d(4,7):7.1f}  iteration period {float(np.mean(np.diff(t[k,5]))):7.1f} ns")
print(f"mma committed -> cell warps see it {float(np.mean(t[k,3]-t[k,2])):7.1f} ns;  cells end(k) -> MMA warp has h(k+2) {float(np.mean(t[22:2*S-2,1]-t[20:2*S-4,7])):7.1f} ns")
print("per-step (2 iterations) period:", float(np.mean(np.diff(t[k, 5]))) * 2, "ns")
