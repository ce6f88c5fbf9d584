"""Generates tests/golden/*.npz from the CPU oracle (run here, committed with its outputs).

The reference itself cannot be imported (Python 2 / TensorFlow 0.1x, neither available), so these
vectors pin the ORACLE (oracle/ -- the hand-restated reference semantics) and are what the CUDA
path is compared with on the GPU box, where /root/reference does not exist.
Usage:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import flownet_s as fs, loss_interp as li, synth, adam as oadam  # noqa: E402

OUT = Path(__file__).resolve().parent
torch.set_num_threads(8)


def loss_case():
    g = torch.Generator().manual_seed(7)
    B, h, w = 2, 12, 16
    flows = (torch.randn(B, h, w, 2, generator=g) * 0.6)
    src = torch.rand(B, h, w, 3, generator=g) - 0.4
    tgt = torch.rand(B, h, w, 3, generator=g) - 0.4
    out = dict(flows=flows.numpy(), src=src.numpy(), tgt=tgt.numpy())
    for variant in ("A", "B"):
        f = flows.clone().requires_grad_(True)
        ld, recon = li.loss_interp(f, src, tgt, 1e-4, 0.25, 0.37, 1.0, 2.5, variant=variant)
        ld["total"].backward()
        out[f"loss4_{variant}"] = np.array([ld[k].item() for k in ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")], np.float32)
        out[f"recon_{variant}"] = recon.detach().numpy()
        out[f"dflow_{variant}"] = f.grad.numpy()
    np.savez_compressed(OUT / "loss_interp_small.npz", **out)


def flownet_case():
    H, W = 192, 256
    src, tgt, gt = synth.make_pairs(1, H, W, seed=11)
    src, tgt = src.round().clamp(0, 255), tgt.round().clamp(0, 255)
    params = fs.init_params(seed=1)
    total, grads, losses, flows_all, prev1 = fs.loss_and_grads(params, src, tgt)
    keys = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
    out = dict(src_u8=src.numpy().astype(np.uint8), tgt_u8=tgt.numpy().astype(np.uint8),
               loss4=np.array([[l[k].item() for k in keys] for l in losses], np.float32),
               total=np.float32(total.item()),
               flow1=flows_all[0].detach().numpy(), flow3=flows_all[2].detach().numpy(), flow6=flows_all[5].detach().numpy(),
               prev1_sub=prev1.detach().numpy()[:, ::8, ::8],
               grad_names=np.array(list(grads.keys())),
               grad_norm=np.array([g.norm().item() for g in grads.values()], np.float64),
               grad_sum=np.array([g.double().sum().item() for g in grads.values()], np.float64))
    # one TF-Adam step and the parameter movement it produces
    opt = oadam.TFAdam(params)
    before = {k: v.clone() for k, v in params.items()}
    opt.step(grads, 1.6e-5)
    out["delta_norm"] = np.array([(params[k] - before[k]).norm().item() for k in params], np.float64)
    np.savez_compressed(OUT / "flownet_s_192x256.npz", **out)


if __name__ == "__main__":
    loss_case()
    flownet_case()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
