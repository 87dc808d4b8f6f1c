"""Device time of the forward / backward scan kernels alone (dev tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-r2d2-dpg_b200")]
import numpy as np, torch
from r2d2_b200 import native as nv
lib = nv.lib()
H, B, S = int(os.environ.get("H", 512)), int(os.environ.get("B", 512)), int(os.environ.get("S", 125))
rep = int(os.environ.get("REP", 1))
T = S // rep
gin = torch.randn(T, B, 4 * H, device="cuda") * 0.5
whh = (torch.rand(4 * H, H, device="cuda") * 2 - 1) / np.sqrt(4 * H)
gates = torch.empty(S, B, 4 * H, device="cuda"); hs = torch.empty(S + 1, B, H, device="cuda"); cs = torch.empty_like(hs)
dgin = torch.empty(T, B, 4 * H, device="cuda")
dh = torch.randn(T, B, H, device="cuda") * 0.01
scratch = torch.empty(B * 4 * H + 64, device="cuda")
st = nv.current_stream()
def fwd():
    nv.check(lib.r2d2_lstm_scan_forward(nv.dptr(gin), nv.dptr(whh), None, None, nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), None,
                                        T, B, H, rep, nv.dptr(scratch), st))
def bwd():
    nv.check(lib.r2d2_lstm_scan_backward(nv.dptr(gates), nv.dptr(hs), nv.dptr(cs), nv.dptr(whh), nv.dptr(dh), 0, nv.dptr(gates),
                                         nv.dptr(dgin), T, B, H, rep, nv.dptr(scratch), st))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    if name == "bwd": fwd()
    fn(); torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        if name == "bwd": pass
        fn()
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    print(f"{name}: H={H} B={B} S={S} rep={rep}: {ms*1e3:.0f} us = {ms*1e3/S:.2f} us/step")
import ctypes
status = ctypes.c_int(0)
nv.check(lib.r2d2_scan_status(ctypes.byref(status), st)); print("scan status", status.value)
