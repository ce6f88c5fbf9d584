"""Analytic known-answer tests that pin the CPU oracle (the reference ships no tests)."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops, loss_interp as li, flownet_s as fs, adam as oadam, metrics
from oracle.warp_literal import warp_literal

EPS, AC, AS = 1e-4, 0.25, 0.37


def test_same_pad_is_asymmetric_after():
    # conv1 7x7/2 on 384 -> (2,3); conv2 5x5/2 -> (1,2); 3x3/2 -> (0,1); 3x3/1 -> (1,1)
    assert tf_ops.same_pad(384, 7, 2) == (192, 2, 3)
    assert tf_ops.same_pad(192, 5, 2) == (96, 1, 2)
    assert tf_ops.same_pad(48, 3, 2) == (24, 0, 1)
    assert tf_ops.same_pad(48, 3, 1) == (48, 1, 1)
    assert tf_ops.same_pad(7, 3, 2) == (4, 1, 1)


def test_conv2d_same_matches_direct_sum():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 6, 8, 3, generator=g)
    w = torch.randn(3, 3, 3, 2, generator=g)
    b = torch.randn(2, generator=g)
    y = tf_ops.conv2d_same(x, w, b, stride=2)
    assert y.shape == (1, 3, 4, 2)
    oy, ox, co = 1, 2, 1
    acc = b[co].item()
    for kh in range(3):
        for kw in range(3):
            iy, ix = oy * 2 + kh - 0, ox * 2 + kw - 0      # pad_before = 0 for 3x3/2 on even sizes
            if 0 <= iy < 6 and 0 <= ix < 8:
                acc += (x[0, iy, ix] * w[kh, kw, :, co]).sum().item()
    assert abs(acc - y[0, oy, ox, co].item()) < 1e-5


def test_conv2d_transpose_is_gradient_of_conv():
    g = torch.Generator().manual_seed(1)
    small = torch.randn(2, 3, 4, 5, generator=g)
    w = torch.randn(4, 4, 7, 5, generator=g)              # [kh,kw,cout,cin]
    up = tf_ops.conv2d_transpose_same(small, w, None, 2)
    assert up.shape == (2, 6, 8, 7)
    big = torch.randn(2, 6, 8, 7, generator=g, requires_grad=True)
    y = tf_ops.conv2d_same(big, w, None, stride=2)         # w read as [kh,kw,cin=7,cout=5]
    (y * small).sum().backward()
    assert torch.allclose(big.grad, up, atol=1e-5)


def test_resize_bilinear_legacy_is_decimation_at_integer_ratio():
    x = torch.arange(2 * 8 * 12 * 3, dtype=torch.float32).reshape(2, 8, 12, 3)
    assert torch.equal(tf_ops.resize_bilinear_legacy(x, 4, 6), x[:, ::2, ::2])
    assert torch.equal(tf_ops.resize_bilinear_legacy(x, 2, 3), x[:, ::4, ::4])
    # non-integer ratio: src = dst * in/out, no half-pixel offset
    y = tf_ops.resize_bilinear_legacy(x[:, :, :, :1], 8, 8)
    sx = 5 * 12 / 8
    x0 = int(math.floor(sx))
    ref = x[0, 3, x0, 0] * (1 - (sx - x0)) + x[0, 3, x0 + 1, 0] * (sx - x0)
    assert abs(y[0, 3, 5, 0].item() - ref.item()) < 1e-4


def test_lrn_three_channels():
    x = torch.tensor([[[[0.3, -0.2, 0.5]]]])
    out = tf_ops.lrn(x)
    den = (1 + 0.09 + 0.04 + 0.25) ** 0.7
    assert torch.allclose(out, x / den, atol=1e-7)


def test_flow_delta_constant_fill_rule():
    w = li.flow_delta_weights()
    assert w.shape == (3, 3, 2, 2)
    nz = {(0, 1, 0, 0): 1.0, (0, 1, 0, 1): -1.0, (1, 0, 0, 1): 1.0, (1, 1, 0, 0): -1.0}
    for idx in np.ndindex(3, 3, 2, 2):
        assert w[idx].item() == nz.get(idx, 0.0)
    assert float(w[:, :, 1].abs().sum()) == 0.0             # V never enters (author bug, reproduced)


def test_bilinear_deconv_weights():
    w = tf_ops.bilinear_deconv_weights((4, 4, 2, 2))
    ax = torch.tensor([0.25, 0.75, 0.75, 0.25])
    assert torch.allclose(w[:, :, 0, 0], torch.outer(ax, ax))
    assert torch.allclose(w[:, :, 1, 1], torch.outer(ax, ax))
    assert float(w[:, :, 0, 1].abs().sum()) == 0.0


# ---------------------------------------------------------------- warp
def _imgs(B=2, h=6, w=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, h, w, 3, generator=g), torch.rand(B, h, w, 3, generator=g)


def test_warp_zero_flow_is_identity():
    _, tgt = _imgs()
    assert torch.equal(li.warp(torch.zeros(2, 6, 8, 2), tgt), tgt)


def test_warp_integer_flow_is_shift_with_edge_clamp():
    _, tgt = _imgs()
    fl = torch.zeros(2, 6, 8, 2)
    fl[..., 0] = 2.0      # U = +2 columns
    fl[..., 1] = -1.0     # V = -1 row
    out = li.warp(fl, tgt)
    for y in range(6):
        for x in range(8):
            assert torch.equal(out[:, y, x], tgt[:, max(y - 1, 0), min(x + 2, 7)])


def test_warp_half_pixel_is_neighbour_average():
    _, tgt = _imgs()
    fl = torch.zeros(2, 6, 8, 2)
    fl[..., 0] = 0.5
    out = li.warp(fl, tgt)
    assert torch.allclose(out[:, :, :7], 0.5 * (tgt[:, :, :7] + tgt[:, :, 1:]), atol=1e-7)
    assert torch.allclose(out[:, :, 7], tgt[:, :, 7], atol=1e-7)      # x1 clamps onto x0 at the border


def test_warp_vectorised_equals_literal_loop():
    g = torch.Generator().manual_seed(3)
    tgt = torch.rand(2, 7, 9, 3, generator=g)
    fl = torch.randn(2, 7, 9, 2, generator=g) * 3.0
    a = li.warp(fl, tgt).numpy()
    b = warp_literal(fl.numpy(), tgt.numpy())
    assert np.max(np.abs(a - b)) < 1e-6


# ---------------------------------------------------------------- losses
def test_loss_A_constant_flow_closed_form():
    B, h, w = 2, 20, 30
    src, _ = _imgs(B, h, w, 1)
    fl = torch.full((B, h, w, 2), 0.0)
    fl[..., 0] = 0.7           # constant U, un-scaled; scale 2 -> 1.4
    ld, recon = li.loss_interp_A(fl, src, src, EPS, AC, AS, 1.0, 2.0)
    bw = 2
    n = B * 3 * (h - 2 * bw) * (w - 2 * bw)
    u = 1.4
    e0 = (EPS ** 2) ** AS
    # out0 = U[y-1,x]-U[y,x]: row 0 sees the zero pad -> -u; last column masked
    u_sum = B * ((w - 1) * (u * u + EPS ** 2) ** AS + (h - 1) * (w - 1) * e0 + h * e0)
    # out1 = U[y,x-1]-U[y-1,x]: (0,0) -> 0 ; row 0, x>0 -> +u ; col 0, 0<y<h-1 -> -u ; last row masked
    v_sum = B * ((w - 1) * (u * u + EPS ** 2) ** AS + (h - 2) * (u * u + EPS ** 2) ** AS
                 + e0 + (h - 2) * (w - 1) * e0 + w * e0)
    assert abs(ld["U_loss"].item() - u_sum / n) < 1e-6 * max(1, u_sum / n)
    assert abs(ld["V_loss"].item() - v_sum / n) < 1e-6 * max(1, v_sum / n)
    assert abs(ld["total"].item() - (ld["Charbonnier_reconstruct"] + ld["U_loss"] + ld["V_loss"]).item()) < 1e-7


def test_loss_B_constant_flow_gives_eps_floor():
    B, h, w = 2, 20, 30
    src, tgt = _imgs(B, h, w, 2)
    fl = torch.full((B, h, w, 2), 0.37)
    ld, _ = li.loss_interp_B(fl, src, tgt, EPS, AC, AS, 1.0, 5.0)
    floor = (EPS ** 2) ** AS
    assert abs(ld["U_loss"].item() - floor) < 1e-9 + 1e-5 * floor
    assert abs(ld["V_loss"].item() - floor) < 1e-9 + 1e-5 * floor


def test_photometric_identical_images_zero_flow():
    B, h, w = 1, 10, 10
    src, _ = _imgs(B, h, w, 4)
    ld, recon = li.loss_interp_A(torch.zeros(B, h, w, 2), src, src, EPS, AC, AS, 0.0, 1.0)
    assert torch.equal(recon, src)
    assert abs(ld["Charbonnier_reconstruct"].item() - (EPS ** 2) ** AC) < 1e-8


def test_border_mask_counts():
    assert li.border_width(192) == 20 and li.border_width(6) == 1 and li.border_width(12) == 2
    m = li.border_mask(12, 16)
    assert m.sum().item() == (12 - 4) * (16 - 4)


@pytest.mark.parametrize("variant", ["A", "B"])
def test_flow_gradient_matches_hand_derivation(variant):
    """SURVEY.md 8a row W'/L': d recon/du = (Ic-Ia)(1-yw)+(Id-Ib)yw etc.; checked through autograd
    against finite differences of the oracle loss (away from the floor discontinuities)."""
    g = torch.Generator().manual_seed(5)
    B, h, w = 1, 12, 14
    src = torch.rand(B, h, w, 3, generator=g, dtype=torch.float64)
    tgt = torch.rand(B, h, w, 3, generator=g, dtype=torch.float64)
    fl = (torch.rand(B, h, w, 2, generator=g, dtype=torch.float64) * 0.8 + 0.1)     # frac in (0.1,0.9) at scale 1
    f = fl.clone().requires_grad_(True)
    ld, _ = li.loss_interp(f, src, tgt, 1e-2, AC, AS, 1.0, 1.0, variant=variant)
    ld["total"].backward()
    for (y, x, c) in [(5, 6, 0), (3, 3, 1), (0, 0, 0), (h - 1, w - 1, 1), (6, w - 1, 0)]:
        d = 1e-6
        fp, fm = fl.clone(), fl.clone()
        fp[0, y, x, c] += d
        fm[0, y, x, c] -= d
        lp, _ = li.loss_interp(fp, src, tgt, 1e-2, AC, AS, 1.0, 1.0, variant=variant)
        lm, _ = li.loss_interp(fm, src, tgt, 1e-2, AC, AS, 1.0, 1.0, variant=variant)
        fd = (lp["total"] - lm["total"]).item() / (2 * d)
        assert abs(fd - f.grad[0, y, x, c].item()) < 1e-5 * max(1.0, abs(fd))


# ---------------------------------------------------------------- model / optimiser / metric
def test_param_inventory():
    shapes = fs.param_shapes()
    assert len(shapes) == 52
    assert fs.num_params() == 38777706
    assert shapes["upconv4/weights"] == (4, 4, 256, 1026) and shapes["pr1/weights"] == (3, 3, 98, 2)


def test_adam_first_step_is_lr_sign():
    p = {"w": torch.tensor([1.0, -2.0, 3.0])}
    g = {"w": torch.tensor([0.5, -0.25, 2.0])}
    opt = oadam.TFAdam(p)
    opt.step(g, 0.01)
    assert torch.allclose(p["w"], torch.tensor([1.0 - 0.01, -2.0 + 0.01, 3.0 - 0.01]), atol=1e-6)
    assert abs(opt.lr_t(0.01) - 0.01 * math.sqrt(1 - 0.999) / (1 - 0.9)) < 1e-12


def test_flow_ee():
    a = torch.zeros(1, 2, 2, 2)
    b = torch.zeros(1, 2, 2, 2)
    b[0, 0, 0] = torch.tensor([3.0, 4.0])
    assert abs(metrics.flow_ee(a, b).item() - 5.0 / 4) < 1e-7


def test_forward_structure_small():
    p = fs.init_params(1)
    g = torch.Generator().manual_seed(0)
    src = torch.rand(1, 192, 256, 3, generator=g) * 255
    tgt = torch.rand(1, 192, 256, 3, generator=g) * 255
    losses, flows_all, prev1, total = fs.forward(p, src, tgt)
    assert [tuple(f.shape[1:3]) for f in flows_all] == [(96, 128), (48, 64), (24, 32), (12, 16), (6, 8), (3, 4)]
    assert prev1.shape == (1, 96, 128, 3) and len(losses) == 6
    want = sum(w * l["total"] for w, l in zip(fs.LOSS_WEIGHTS, losses))
    assert abs(total.item() - want.item()) < 1e-5


# ---- FlowNetC correlation (paper definition; the reference has no such layer) and VGG16 oracle pieces ----------------------------

def test_correlation_known_answers():
    """corr[b,y,x,(i,j)] = <f1[y,x], f2[y+dy_i, x+dx_j]> / C with zero outside: the centre channel of a map with itself is the mean square,
    a shifted copy moves the peak to the matching displacement, channels pointing outside the map are exactly zero."""
    from oracle import flownet_c as oc
    g = torch.Generator().manual_seed(0)
    B, h, w, c, md, s2 = 1, 9, 11, 8, 4, 2
    D = 2 * (md // s2) + 1
    f = torch.randn(B, h, w, c, generator=g)
    out = oc.correlation(f, f, md, s2)
    assert out.shape == (B, h, w, D * D)
    centre = (D // 2) * D + D // 2
    assert torch.allclose(out[..., centre], (f * f).mean(-1), atol=1e-6)
    # f2 = f1 shifted by (dy, dx) = (+2, -2): the response at channel (dy=+2, dx=-2) equals the mean square wherever the shift stays inside
    f2 = torch.zeros_like(f)
    f2[:, 2:, :w - 2] = f[:, :h - 2, 2:]
    out2 = oc.correlation(f, f2, md, s2)
    ch = ((2 + md) // s2) * D + ((-2 + md) // s2)
    assert torch.allclose(out2[:, :h - 2, 2:, ch], (f * f).mean(-1)[:, :h - 2, 2:], atol=1e-6)
    # top-left pixel, displacement (-4, -4) looks outside the map -> exactly zero
    assert float(out[0, 0, 0, 0].abs()) == 0.0
    # linear in each argument
    a = torch.randn(B, h, w, c, generator=g)
    assert torch.allclose(oc.correlation(2.0 * f + a, f2, md, s2), 2.0 * out2 + oc.correlation(a, f2, md, s2), atol=1e-5)


def test_flownetc_and_vgg16_oracle_structure():
    from oracle import flownet_c as oc, vgg16 as ov
    shapes = oc.param_shapes()
    assert shapes["conv3_1/weights"] == (3, 3, 473, 256) and shapes["conv_redir/weights"] == (1, 1, 256, 32)      # 32 + 21*21 channels
    assert "conv1/weights" in shapes and shapes["conv1/weights"] == (7, 7, 3, 64)                                # siamese: 3 input channels
    vs = ov.param_shapes()
    assert sum(1 for k in vs if k.startswith("conv") and k.endswith("weights")) == 13
    x = torch.arange(2 * 4 * 6 * 1, dtype=torch.float32).reshape(2, 4, 6, 1)
    y = ov.max_pool2(x)
    assert y.shape == (2, 2, 3, 1) and torch.equal(y[0, :, :, 0], torch.tensor([[7., 9., 11.], [19., 21., 23.]]))


def test_bench_host_cores_and_reference_contract(monkeypatch):
    """bench.py's CPU arm: a sane thread count (physical cores within the affinity mask, cgroup quota cap) and the reference-arm JSON keys."""
    import importlib
    import json
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    bench = importlib.import_module("bench")
    n = bench._host_cores()
    assert 1 <= n <= (len(__import__("os").sched_getaffinity(0)) if hasattr(__import__("os"), "sched_getaffinity") else 4096)
    lines = []
    monkeypatch.setattr("builtins.print", lambda *a, **k: lines.append(a[0]))
    monkeypatch.setattr(bench, "cpu_step_rate", lambda sample, steps, warm: (dict(value=2.0, unit=bench.UNIT, cores=n, kind="port", sample="stub"), 1.0))
    args = type("A", (), dict(steps=3, warmup=1, gpus=2))()
    bench.run_reference(args, 1, 2)                 # other ranks: no work, no line
    assert lines == []
    bench.run_reference(args, 0, 2)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["e2e"] == {"value": 2.0, "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["n_gpus"] == 2
