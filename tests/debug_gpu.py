"""Scratch diagnostics run on the GPU box (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from oracle import loss_interp as li, flownet_s as fs, adam as oadam, synth
from deepof_b200 import flyingChairsWrapFlow as W
from deepof_b200.flownet import FlowNetS
EPS, AC, AS = 1e-4, 0.25, 0.37
g = torch.Generator().manual_seed(2)
B, h, w = 2, 12, 16
flows = torch.randn(B, h, w, 2, generator=g) * 0.5
src = torch.rand(B, h, w, 3, generator=g); tgt = torch.rand(B, h, w, 3, generator=g)
for coef in [(1.0, 0.0), (2.0, 0.5), (0.0, 1.0)]:
    fc = flows.cuda().requires_grad_(True)
    ld, _ = W.loss_interp(fc, src.cuda(), tgt.cuda(), EPS, AC, AS, 1.0, 2.5, None)
    (coef[0] * ld["total"] + coef[1] * ld["U_loss"]).backward()
    f = flows.clone().requires_grad_(True)
    ldr, _ = li.loss_interp(f, src, tgt, EPS, AC, AS, 1.0, 2.5, variant="A")
    (coef[0] * ldr["total"] + coef[1] * ldr["U_loss"]).backward()
    d = (fc.grad.cpu() - f.grad).abs()
    i = d.argmax().item()
    print("coef", coef, "max abs diff", d.max().item(), "ref max", f.grad.abs().max().item(), "at", np.unravel_index(i, d.shape),
          "ours", fc.grad.cpu().flatten()[i].item(), "ref", f.grad.flatten()[i].item())
# adam tracking stats
B, H, Wd = 2, 192, 256
s, t, _ = synth.make_pairs(B, H, Wd, seed=5)
params = fs.init_params(1); opt = oadam.TFAdam(params)
eng = FlowNetS(B, H, Wd, seed=None); eng.load_params(params)
for it in range(2):
    _t, grads, *_ = fs.loss_and_grads(params, s, t)
    opt.step(grads, 1.6e-5)
    eng.train_step(s.cuda(), t.cuda(), fs.LOSS_WEIGHTS, 1.6e-5)
lr = 1.6e-5
for name in list(params)[:6] + list(params)[-4:]:
    d = (eng.params[name].cpu() - params[name]).abs()
    print(name, "max/lr", d.max().item() / lr, "frac>0.1lr", (d > 0.1 * lr).float().mean().item(), "mean/lr", d.mean().item() / lr)
