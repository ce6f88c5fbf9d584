"""Prints the oracle-anchored precision numbers of the three math modes (used to set deepof_b200/precision.py TOLERANCE)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from oracle import flownet_s as ofs
from deepof_b200.flownet import FlowNetS, FLOW_SCALES
from deepof_b200 import precision
from deepof_b200.synth import make_pairs

H, W, B = 192, 256, 4
data = [tuple(x.cuda() for x in make_pairs(B, H, W, seed=100 + i)[:2]) for i in range(4)]
eng = FlowNetS(B, H, W, math_mode="bf16", seed=1, tc_wgrad=True)
for i in range(300):
    eng.train_step(*data[i % 4], lr=1.6e-4)
trained = eng.export_params()
src, tgt, _ = make_pairs(2, H, W, seed=1234)
for which, params in (("init", ofs.init_params(1)), ("trained", trained)):
    with torch.no_grad():
        _l, ref, _p, _t = ofs.forward(params, src, tgt)
    print(which, "oracle |flow| per scale", [round(v, 4) for v in precision.flow_magnitude(ref)])
    for mode in ("fp32", "tf32", "bf16"):
        e = FlowNetS(2, H, W, math_mode=mode, seed=None, tc_wgrad=mode != "fp32")
        e.load_params(params)
        e.forward(src.cuda(), tgt.cuda(), with_grad=False)
        st = precision.end_point_distance([e.pr[s] * FLOW_SCALES[s] for s in range(1, 7)], ref)
        print(f"  {mode}: mean", [f"{m:.2e}" for m, _ in st], "max", [f"{x:.2e}" for _, x in st])
data = [tuple(x.cuda() for x in make_pairs(B, H, W, seed=300 + i)[:2]) for i in range(5)]
for lr in (1.6e-5, 1.6e-4):
    traj = {}
    for mode in ("fp32", "tf32", "bf16"):
        e = FlowNetS(B, H, W, math_mode=mode, seed=1, tc_wgrad=mode != "fp32")
        ls = []
        for i in range(50):
            e.train_step(*data[i % 5], lr=lr)
            ls.append(e.total_loss().reshape(1).clone())
        traj[mode] = torch.cat(ls).cpu().double()
    ref = traj["fp32"]
    print("lr", lr, "fp32 first/last", float(ref[:5].mean()), float(ref[-5:].mean()))
    for mode in ("tf32", "bf16"):
        g = ((traj[mode] - ref).abs() / ref.abs())
        print(f"  {mode}: max rel gap {float(g.max()):.3e} mean {float(g.mean()):.3e}; descent {float(traj[mode][:5].mean() - traj[mode][-5:].mean()):.4f} vs {float(ref[:5].mean() - ref[-5:].mean()):.4f}")
