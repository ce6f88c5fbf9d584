"""Gradient agreement of the bf16 engines (lean / classic, with and without split-K / phase-in-N) with the fp32 engine on one small batch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepof_b200.flownet import FlowNetS
from deepof_b200.synth import make_pairs
B, H, W = 2, 192, 256
src, tgt, _ = make_pairs(B, H, W, seed=21)
def grads(mode, lean, splitk, pin):
    os.environ["DOFB_LEAN"] = lean; os.environ["DOFB_SPLITK"] = splitk; os.environ["DOFB_PIN"] = pin
    e = FlowNetS(B, H, W, math_mode=mode, seed=1, tc_wgrad=mode != "fp32")
    e.forward(src.cuda(), tgt.cuda(), with_grad=True); e.backward(); torch.cuda.synchronize()
    return {k: v.double().flatten().clone() for k, v in e.grads.items()}, e.loss4.clone()
ref, l0 = grads("fp32", "0", "0", "0")
def cos(a, b): return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
for lean, sk, pin in (("0", "0", "0"), ("1", "0", "0"), ("1", "1", "0"), ("1", "1", "2"), ("0", "1", "2"), ("1", "0", "2")):
    g, l = grads("bf16", lean, sk, pin)
    cs = {k: cos(g[k], ref[k]) for k in ref}
    worst = sorted(cs.items(), key=lambda kv: kv[1])[:3]
    print(f"lean={lean} splitk={sk} pin={pin}: loss rel {float((l - l0).abs().max() / l0.abs().max()):.2e}; worst cos vs fp32:", [(k, round(v, 4)) for k, v in worst])
