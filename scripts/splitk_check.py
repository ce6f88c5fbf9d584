"""Engine-level A/B of split-K: every activation, flow and parameter gradient with DOFB_SPLITK=0 vs 1 (tf32 and bf16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepof_b200.flownet import FlowNetS
from deepof_b200.synth import make_pairs
B, H, W = 2, 384, 512
src, tgt, _ = make_pairs(B, H, W, seed=1234)
def run(mode, sk):
    os.environ["DOFB_SPLITK"] = sk
    e = FlowNetS(B, H, W, math_mode=mode, seed=1, tc_wgrad=True)
    e.forward(src.cuda(), tgt.cuda(), with_grad=True); e.backward(); torch.cuda.synchronize()
    return e
def rel(a, b): return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
for mode in ("tf32", "bf16"):
    e0, e1, e2 = run(mode, "0"), run(mode, "1"), run(mode, "1")
    print(mode, "pr:", [f"{rel(e1.pr[s], e0.pr[s]):.1e}" for s in range(1, 7)], "dpr:", [f"{rel(e1.dpr[s], e0.dpr[s]):.1e}" for s in range(1, 7)])
    rows = sorted(((rel(e1.grads[n], e0.grads[n]), rel(e2.grads[n], e1.grads[n]), n) for n in e0.grads), reverse=True)
    for r in rows[:12]: print(f"   {r[2]:22s} splitk1 vs 0: {r[0]:.2e}   splitk1 run-to-run: {r[1]:.2e}")
    for name in ("act", "dact"):
        d0, d1 = getattr(e0, name, None), getattr(e1, name, None)
        if not isinstance(d0, dict): continue
        out = []
        for k in d0:
            a, b = d0[k], d1[k]
            ta, tb = (a.t if a.t is not None else a.t16.float()), (b.t if b.t is not None else b.t16.float())
            out.append((rel(tb, ta), k))
        print("  ", name, [(k, f"{v:.1e}") for v, k in sorted(out, reverse=True)[:8]])
