"""Run a few cfg-2 learner steps (for ncu captures: ncu -k regex:<kernel> -s <skip> -c 1 python tools/prof_step.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-r2d2-dpg_b200")]
import torch
from r2d2_b200 import engine
from oracle import ref_port
c = dict(obs=17, act=6, hidden=256, batch=256, burn_in=40, learning=80, n_step=5)
eng = engine.LearnerEngine(engine.PathConfig(**c))
eng.set_batch(ref_port.synthetic_batch(ref_port.PathConfig(**c), 0))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    eng.step()
torch.cuda.synchronize()
print("done")
