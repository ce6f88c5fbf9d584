"""Multi-frame warp/loss (sintelWrapFlow.loss_interp_multi, sintelWrapFlow.py:492-630) and edge-aware smoothness
(version1/model/warpflow.py:91-116,148-157) on the device against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import loss_interp as li

pytestmark = pytest.mark.gpu
KEYS = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("shape", [(2, 12, 16), (1, 33, 21), (3, 24, 40)])
def test_edge_weights_match_oracle(shape):
    from deepof_b200 import ops
    B, h, w = shape
    g = torch.Generator().manual_seed(h + w)
    img = torch.rand(B, h, w, 3, generator=g) - 0.4
    img[0, :, : w // 2] *= 0.2                       # a strong vertical edge
    want = li.edge_weights(img)
    got = ops.edge_weights(img.cuda()).cpu()
    # integer quantisation + exact small-integer Sobel + one IEEE division: bit-identical up to the final 1 - |.|
    assert float((got - want).abs().max()) < 1e-6
    assert float(got.min()) >= 0.0 and float(got.max()) <= 1.0


@pytest.mark.parametrize("shape,scale", [((2, 12, 16), 2.5), ((1, 33, 21), 5.0), ((2, 24, 32), 0.625)])
def test_edge_aware_loss_matches_oracle(shape, scale):
    from deepof_b200 import warpflow
    B, h, w = shape
    g = torch.Generator().manual_seed(7 * h + w)
    flows = torch.randn(B, h, w, 2, generator=g) * (2.0 / scale)
    src = torch.rand(B, h, w, 3, generator=g) - 0.4
    tgt = torch.rand(B, h, w, 3, generator=g) - 0.4
    f = flows.clone().requires_grad_(True)
    ld, recon_ref = li.loss_interp_B_edge(f, src, tgt, 1e-4, 0.25, 0.37, 1.0, scale)
    (2.0 * ld["total"]).backward()
    fc = flows.cuda().requires_grad_(True)
    got, recon = warpflow.loss_interp(fc, src.cuda(), tgt.cuda(), 1e-4, 0.25, 0.37, 1.0, scale,
                                      {"needMask": True, "needImageGradients": True})
    (2.0 * got["total"]).backward()
    for k in KEYS:
        assert abs(float(got[k]) - float(ld[k])) <= 3e-6 * max(1.0, abs(float(ld[k]))), k
    assert float((recon.cpu() - recon_ref.detach()).abs().max()) < 1e-6
    assert rel(fc.grad, f.grad) < 3e-5
    # and it differs from the un-weighted loss
    plain, _ = warpflow.loss_interp(flows.cuda(), src.cuda(), tgt.cuda(), 1e-4, 0.25, 0.37, 1.0, scale, {"needMask": True})
    assert abs(float(plain["U_loss"]) - float(got["U_loss"])) > 1e-4


@pytest.mark.parametrize("B,h,w,T,scale,lam", [(2, 12, 16, 4, 1.25, 0.5), (1, 33, 21, 3, 5.0, 1.0), (2, 24, 32, 10, 2.5, 0.0),
                                               (1, 12, 16, 2, 0.625, 1.0)])
def test_loss_interp_multi_matches_oracle(B, h, w, T, scale, lam):
    from deepof_b200 import sintelWrapFlow as sw
    g = torch.Generator().manual_seed(B + h + w + T)
    frames = torch.rand(B, h, w, 3 * T, generator=g) - 0.4
    flows = torch.randn(B, h, w, 2 * (T - 1), generator=g) * (2.0 / scale)
    f = flows.clone().requires_grad_(True)
    ld, recon_ref = li.loss_interp_multi(f, frames, 1e-4, 0.3, 0.3, lam, scale)
    (3.0 * ld["total"]).backward()
    fc = flows.cuda().requires_grad_(True)
    got, recon = sw.loss_interp_multi(fc, frames.cuda(), 1e-4, 0.3, 0.3, lam, scale, None)
    (3.0 * got["total"]).backward()
    for k in KEYS:
        assert abs(float(got[k]) - float(ld[k])) <= 3e-6 * max(1.0, abs(float(ld[k]))), k
    assert tuple(recon.shape) == (B, h, w, 3 * (T - 1))
    assert float((recon.cpu() - recon_ref.detach()).abs().max()) < 1e-6
    assert rel(fc.grad, f.grad) < 3e-5
    # the constant the default stands for is the reference's short-list fill
    assert torch.equal(sw.flow_delta_weights(2 * (T - 1)), li.flow_delta_weights_multi(2 * (T - 1)))


def test_loss_interp_multi_two_frames_is_the_two_frame_photometric_term():
    """T = 2 degenerates to one pair: the photometric term equals the two-frame loss_interp's (sintelWrapFlow.py:632-766)."""
    from deepof_b200 import sintelWrapFlow as sw
    g = torch.Generator().manual_seed(3)
    B, h, w = 2, 12, 16
    a, b = torch.rand(B, h, w, 3, generator=g) - 0.4, torch.rand(B, h, w, 3, generator=g) - 0.4
    flows = torch.randn(B, h, w, 2, generator=g)
    m, recon_m = sw.loss_interp_multi(flows.cuda(), torch.cat([a, b], dim=3).cuda(), 1e-4, 0.3, 0.3, 0.0, 1.25, None)
    t, recon_t = sw.loss_interp(flows.cuda(), a.cuda(), b.cuda(), 1e-4, 0.3, 0.3, 0.0, 1.25, None)
    assert abs(float(m["Charbonnier_reconstruct"]) - float(t["Charbonnier_reconstruct"])) < 1e-6
    assert torch.equal(recon_m, recon_t)


def test_shape_errors_are_loud():
    from deepof_b200 import sintelWrapFlow as sw, DeepOFError
    with pytest.raises(DeepOFError):
        sw.loss_interp_multi(torch.zeros(1, 4, 4, 3, device="cuda"), torch.zeros(1, 4, 4, 9, device="cuda"), 1e-4, .3, .3, 0., 1., None)
