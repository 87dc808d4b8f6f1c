"""Timeline of the H = 512 BPTT scan (dev tool).  Stamps per step: 0 top, 1 partial sums arrived, 2 cell math done,
3 block barrier passed, 4 MMAs issued (warp 0), 5 warp 15 sees its tile done, 6 warp 15 staged, 7 warp 15 load issued."""
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
dh = torch.randn(S, B, H, device="cuda") * 0.01
scratch = torch.empty(B * 4 * H + 64, device="cuda")
st = nv.current_stream()
nv.check(lib.r2d2_lstm_scan_forward(nv.dptr(gin), nv.dptr(whh), None, None, nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), None,
                                    S, B, H, 1, nv.dptr(scratch), st))
grid = 16 * 16
trace = torch.zeros(grid, S, 8, dtype=torch.int64, device="cuda")
dgates = torch.empty_like(gates)
for _ in range(2):
    nv.check(lib.r2d2_debug_scan_backward_trace(nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), nv.dptr(whh), nv.dptr(dh), nv.dptr(dgates),
                                                S, B, H, nv.dptr(trace, torch.int64), st))
torch.cuda.synchronize()
for cta in (0, 5, 16 * 5 + 7, 16 * 9 + 3, 16 * 13):
    t = trace.cpu().numpy().astype(np.float64)[cta]
    k = slice(20, S - 3)
    d = lambda a, b: float(np.mean(t[k, b] - t[k, a]))
    print(f"--- CTA {cta}: step period {float(np.mean(np.diff(t[k, 0]))):.0f} ns | wait sums {d(0,1):5.0f} | cell math {d(1,2):5.0f} | fence+barrier {d(2,3):5.0f} |"
          f" barrier->MMAs issued {d(3,4):5.0f} | issued->tile 3 done {d(4,5):5.0f} | ld+stage {d(5,6):5.0f} | store+wait+load issue {d(6,7):5.0f} |"
          f" load issued -> next step's sums arrived {float(np.mean(t[21:S-2, 1] - t[20:S-3, 7])):5.0f}")
