"""Scratch diagnostics run on the GPU box (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from oracle import flownet_s as fs, synth
from deepof_b200.flownet import FlowNetS

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))

B, H, W = 2, 192, 256
src, tgt, gt = synth.make_pairs(B, H, W, seed=5)
params = fs.init_params(seed=1)
total, grads, losses, flows_all, prev1 = fs.loss_and_grads(params, src, tgt)
p64 = {k: v.double() for k, v in params.items()}
t64, g64, *_ = fs.loss_and_grads(p64, src.double(), tgt.double())
eng = FlowNetS(B, H, W, seed=None); eng.load_params(params)
eng.forward(src.cuda(), tgt.cuda(), fs.LOSS_WEIGHTS, with_grad=True); eng.backward(); torch.cuda.synchronize()
print("total", total.item(), t64.item(), eng.total_loss().item())
for name in grads:
    r_dev = rel(eng.grads[name], grads[name]); r_dev64 = rel(eng.grads[name], g64[name]); r_cpu64 = rel(grads[name], g64[name])
    flag = " <<<" if r_dev > 2e-3 else ""
    print(f"{name:22s} dev-vs-cpu32 {r_dev:.2e}  dev-vs-cpu64 {r_dev64:.2e}  cpu32-vs-cpu64 {r_cpu64:.2e}{flag}")
for s in range(1, 7):
    d = (eng.pr[s].cpu() - (flows_all[s-1] / fs.FLOW_SCALES[s]).detach()).abs().max().item()
    print("pr", s, "max abs diff", d)
