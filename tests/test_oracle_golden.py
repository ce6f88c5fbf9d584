"""The oracle reproduces its committed golden vectors (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import loss_interp as li, flownet_s as fs

KEYS = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")


def test_loss_interp_golden(golden_dir):
    z = np.load(golden_dir / "loss_interp_small.npz")
    flows, src, tgt = (torch.from_numpy(z[k]) for k in ("flows", "src", "tgt"))
    for variant in ("A", "B"):
        f = flows.clone().requires_grad_(True)
        ld, recon = li.loss_interp(f, src, tgt, 1e-4, 0.25, 0.37, 1.0, 2.5, variant=variant)
        ld["total"].backward()
        got = np.array([ld[k].item() for k in KEYS], np.float32)
        assert np.allclose(got, z[f"loss4_{variant}"], rtol=1e-6, atol=1e-7)
        assert np.allclose(recon.detach().numpy(), z[f"recon_{variant}"], atol=1e-7)
        assert np.allclose(f.grad.numpy(), z[f"dflow_{variant}"], rtol=1e-5, atol=1e-7)


def test_flownet_golden_forward(golden_dir):
    z = np.load(golden_dir / "flownet_s_192x256.npz")
    src = torch.from_numpy(z["src_u8"].astype(np.float32))
    tgt = torch.from_numpy(z["tgt_u8"].astype(np.float32))
    params = fs.init_params(seed=1)
    with torch.no_grad():
        losses, flows_all, prev1, total = fs.forward(params, src, tgt)
    got = np.array([[l[k].item() for k in KEYS] for l in losses], np.float32)
    assert np.allclose(got, z["loss4"], rtol=2e-5, atol=1e-6)
    assert np.allclose(flows_all[0].numpy(), z["flow1"], atol=2e-5)
    assert np.allclose(flows_all[5].numpy(), z["flow6"], atol=2e-5)
