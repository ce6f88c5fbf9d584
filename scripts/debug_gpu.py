"""Scratch diagnostics run on the GPU box (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepof_b200.flownet import FlowNetS
from deepof_b200.synth import make_pairs
from deepof_b200 import ops
import deepof_b200.flownet as F
orig_k = F.FlowNetS._k
def k(self, tag, fn, *a, **kw):
    r = fn(*a, **kw)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED at", tag, type(self).__name__, "B", self.B, "math", self.math, str(e)[:100]); raise
    return r
F.FlowNetS._k = k
H, W = 384, 512
for (B, mode) in [(4, "fp32"), (4, "bf16"), (32, "bf16"), (32, "fp32")]:
    s, t, _ = make_pairs(B, H, W, seed=1)
    e = FlowNetS(B, H, W, math_mode=mode, tc_wgrad=(mode != "fp32"))
    for i in range(2):
        e.train_step(s.cuda(), t.cuda(), lr=1.6e-5)
    print("ok", B, mode, float(e.total_loss()))
    del e
    torch.cuda.empty_cache()
