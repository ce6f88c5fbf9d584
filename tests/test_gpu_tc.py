"""tcgen05 (TF32) conv family vs the fp32 SIMT kernels on the same inputs (both on the device), then vs the oracle.

TF32 keeps 10 mantissa bits of each operand and accumulates in fp32: tolerance 3e-3 of the output max-norm."""
import math

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

TOL = 3e-3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _buf(B, h, w, ld, c, gen):
    t = torch.zeros(B, h, w, ld, device="cuda")
    t[..., :c] = torch.randn(B, h, w, c, generator=gen).cuda()
    return t


def _oracle_conv(x, w, b, stride, ci, co, elu=True):
    """CPU oracle (oracle/tf_ops.py, float64) of slim.conv2d on the dense channels of a pitched buffer."""
    from oracle import tf_ops
    y = tf_ops.conv2d_same(x.cpu()[..., :ci].double(), w.cpu().double(), b.cpu().double() if b is not None else None, stride)
    return tf_ops.elu(y) if elu else y


FWD_CASES = [
    # B, H, W, ci, x_ld, co, k, s
    (2, 24, 32, 64, 128, 128, 5, 2),      # conv2-like, parity gather, asymmetric pad (1,2)
    (2, 12, 16, 256, 256, 256, 3, 1),     # conv3_2-like
    (4, 12, 16, 256, 416, 512, 3, 2),     # conv4_1-like: slab of a concat buffer, pad (0,1)
    (8, 6, 8, 512, 512, 1024, 3, 2),      # conv6_1-like: 3x4 output map
    (8, 3, 4, 1024, 1024, 1024, 3, 1),    # conv6_2-like
    (1, 48, 64, 128, 224, 256, 5, 2),     # conv3_1-like
    (2, 10, 14, 32, 32, 32, 3, 1),        # ragged map, small N
    (2, 16, 16, 96, 96, 64, 3, 1),
]


@pytest.mark.parametrize("case", FWD_CASES)
def test_tc_conv_fwd_and_dgrad_match_simt(case):
    from deepof_b200 import ops
    B, H, W, ci, x_ld, co, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    x = _buf(B, H, W, x_ld, ci, g)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    y_ld = (co + 31) // 32 * 32 + 32
    y0 = torch.zeros(B, geom.oh, geom.ow, y_ld, device="cuda")
    y1 = torch.zeros_like(y0)
    ops.conv_fwd(geom, ops.Slab(x, 0, ci), w, b, ops.Slab(y0, 32, co), ops.ACT_ELU, ops.MATH_FP32)
    ops.conv_fwd(geom, ops.Slab(x, 0, ci), w, b, ops.Slab(y1, 32, co), ops.ACT_ELU, ops.MATH_TF32)
    torch.cuda.synchronize()
    assert rel(y1, y0) < TOL
    assert float(y1[..., :32].abs().max()) == 0.0
    # ... and against the CPU oracle (not only against our own SIMT kernel)
    want = _oracle_conv(x, w, b, s, ci, co)
    assert rel(y1[..., 32:32 + co], want) < TOL
    assert rel(y0[..., 32:32 + co], want) < 2e-5
    # input gradient
    dy = _buf(B, geom.oh, geom.ow, (co + 31) // 32 * 32, co, g)
    d0 = torch.full((B, H, W, x_ld), 0.25, device="cuda")
    d1 = torch.full((B, H, W, x_ld), 0.25, device="cuda")
    for acc in (True, False):
        ops.conv_dgrad(geom, ops.Slab(dy, 0, co), w, None, ops.Slab(d0, 0, ci), ops.ACT_NONE, acc, ops.MATH_FP32)
        ops.conv_dgrad(geom, ops.Slab(dy, 0, co), w, None, ops.Slab(d1, 0, ci), ops.ACT_NONE, acc, ops.MATH_TF32)
        torch.cuda.synchronize()
        assert rel(d1[..., :ci], d0[..., :ci]) < TOL, acc
    # oracle: autograd of the TF-SAME conv (the last call above wrote the plain, non-accumulated gradient)
    xd = torch.zeros(B, H, W, ci, dtype=torch.float64, requires_grad=True)
    _oracle_conv(xd, w, None, s, ci, co, elu=False).backward(dy.cpu()[..., :co].double())
    assert rel(d1[..., :ci], xd.grad) < TOL


DECONV_CASES = [
    # B, h, w, cfeat, feat_ld, upc
    (2, 6, 8, 1024, 1024, 512),
    (2, 12, 16, 1026, 1056, 256),
    (1, 24, 32, 770, 800, 128),
    (1, 48, 64, 386, 416, 64),
    (2, 24, 32, 194, 224, 32),
]


@pytest.mark.parametrize("case", DECONV_CASES)
def test_tc_deconv_fwd_and_dgrad_match_simt(case):
    from deepof_b200 import ops
    B, h, w, cfeat, fld, upc = case
    g = torch.Generator().manual_seed(sum(case))
    x = _buf(B, h, w, fld, cfeat, g)
    wt = (torch.randn(4, 4, upc, cfeat, generator=g) / math.sqrt(4 * cfeat)).cuda()
    b = (torch.randn(upc, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, 2 * h, 2 * w, upc, cfeat, 4, 2)
    y0 = torch.zeros(B, 2 * h, 2 * w, upc + 64, device="cuda")
    y1 = torch.zeros_like(y0)
    ops.conv_dgrad(geom, ops.Slab(x, 0, cfeat), wt, b, ops.Slab(y0, 32, upc), ops.ACT_ELU, False, ops.MATH_FP32)
    ops.conv_dgrad(geom, ops.Slab(x, 0, cfeat), wt, b, ops.Slab(y1, 32, upc), ops.ACT_ELU, False, ops.MATH_TF32)
    torch.cuda.synchronize()
    assert rel(y1, y0) < TOL
    from oracle import tf_ops
    want = tf_ops.elu(tf_ops.conv2d_transpose_same(x.cpu()[..., :cfeat].double(), wt.cpu().double(), b.cpu().double(), 2))
    assert rel(y1[..., 32:32 + upc], want) < TOL          # slim.conv2d_transpose restated by the CPU oracle
    dy = _buf(B, 2 * h, 2 * w, (upc + 31) // 32 * 32, upc, g)
    d0 = torch.zeros(B, h, w, fld, device="cuda")
    d1 = torch.zeros(B, h, w, fld, device="cuda")
    ops.conv_fwd(geom, ops.Slab(dy, 0, upc), wt, None, ops.Slab(d0, 0, cfeat), ops.ACT_NONE, ops.MATH_FP32)
    ops.conv_fwd(geom, ops.Slab(dy, 0, upc), wt, None, ops.Slab(d1, 0, cfeat), ops.ACT_NONE, ops.MATH_TF32)
    torch.cuda.synchronize()
    assert rel(d1[..., :cfeat], d0[..., :cfeat]) < TOL


WGRAD_CASES = [
    # B, H, W, ci, x_ld, co, dy_ld, k, s
    (2, 24, 32, 64, 128, 128, 224, 5, 2),       # conv2-like (swap: ci < 128 <= co)
    (2, 12, 16, 256, 256, 256, 416, 3, 1),      # conv3_2-like
    (4, 12, 16, 256, 416, 512, 512, 3, 2),      # conv4_1-like
    (8, 3, 4, 1024, 1024, 1024, 1024, 3, 1),    # conv6_2-like
    (2, 12, 16, 256, 800, 1026, 1056, 4, 2),    # upconv4-like transposed conv: x = large map (upc=256), dy = small (cfeat=1026)
    (2, 48, 64, 32, 128, 194, 224, 4, 2),       # upconv1-like (swap, ci=32)
    (2, 10, 14, 32, 32, 32, 32, 3, 1),          # ragged map
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_tc_wgrad_matches_simt(case):
    from deepof_b200 import ops
    B, H, W, ci, x_ld, co, dy_ld, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    x = _buf(B, H, W, x_ld, ci, g)
    dy = _buf(B, geom.oh, geom.ow, dy_ld, co, g)
    dw0 = torch.zeros(k, k, ci, co, device="cuda")
    dw1 = torch.zeros(k, k, ci, co, device="cuda")
    db0 = torch.zeros(co, device="cuda")
    db1 = torch.zeros(co, device="cuda")
    ops.conv_wgrad(geom, ops.Slab(x, 0, ci), ops.Slab(dy, 0, co), dw0, db0, ops.MATH_FP32)
    ops.conv_wgrad(geom, ops.Slab(x, 0, ci), ops.Slab(dy, 0, co), dw1, db1, ops.MATH_TF32)
    torch.cuda.synchronize()
    assert rel(dw1, dw0) < TOL
    assert rel(db1, db0) < 1e-4
    # oracle: autograd of the TF-SAME conv with respect to the weights
    wd = torch.zeros(k, k, ci, co, dtype=torch.float64, requires_grad=True)
    _oracle_conv(x, wd, None, s, ci, co, elu=False).backward(dy.cpu()[..., :co].double())
    assert rel(dw1, wd.grad) < TOL


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 384, 512), (3, 48, 80)])
def test_tc_conv1_fwd_and_wgrad_match_simt(B, H, W):
    """First-layer 7x7/2 conv (ci=6) on tensor cores from the zero-bordered buffer vs the SIMT kernel on the dense one."""
    from deepof_b200 import ops
    g = torch.Generator().manual_seed(B * H + W)
    ci, co, k = 6, 64, 7
    img = torch.randn(B, H, W, ci, generator=g).cuda()
    dense = torch.zeros(B, H, W, 8, device="cuda")
    dense[..., :ci] = img
    padded = torch.zeros(B, H + 6, W + 8, 8, device="cuda")
    padded[:, 2:2 + H, 2:2 + W, :ci] = img
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, 2)
    y0 = torch.zeros(B, geom.oh, geom.ow, 128, device="cuda")
    y1 = torch.zeros_like(y0)
    ops.conv_fwd(geom, ops.Slab(dense, 0, ci), w, b, ops.Slab(y0, 0, co), ops.ACT_ELU, ops.MATH_FP32)
    ops.conv1_fwd(geom, padded, (2, 2), w, b, ops.Slab(y1, 0, co), ops.ACT_ELU)
    torch.cuda.synchronize()
    assert rel(y1, y0) < TOL
    dy = _buf(B, geom.oh, geom.ow, 128, co, g)
    dw0 = torch.zeros(k, k, ci, co, device="cuda"); dw1 = torch.zeros_like(dw0)
    db0 = torch.zeros(co, device="cuda"); db1 = torch.zeros_like(db0)
    ops.conv_wgrad(geom, ops.Slab(dense, 0, ci), ops.Slab(dy, 0, co), dw0, db0, ops.MATH_FP32)
    ops.conv1_wgrad(geom, padded, (2, 2), ops.Slab(dy, 0, co), dw1, db1)
    torch.cuda.synchronize()
    assert rel(dw1, dw0) < TOL
    assert rel(db1, db0) < 1e-4


@pytest.mark.parametrize("B,h,w,md,s2", [(2, 12, 64, 20, 2), (1, 7, 40, 20, 2), (2, 6, 128, 20, 2), (1, 5, 64, 16, 4), (1, 48, 64, 20, 2)])
def test_tc_correlation_matches_simt(B, h, w, md, s2):
    from deepof_b200 import ops
    g = torch.Generator().manual_seed(B * h + w)
    c = 256
    f1 = _buf(B, h, w, c, c, g)
    f2 = _buf(B, h, w, c, c, g)
    D = 2 * (md // s2) + 1
    o0 = torch.zeros(B, h, w, D * D + 7, device="cuda")
    o1 = torch.full((B, h, w, D * D + 7), 5.0, device="cuda")
    ops.corr_fwd(ops.Slab(f1, 0, c), ops.Slab(f2, 0, c), ops.Slab(o0, 0, D * D), md, s2, ops.ACT_ELU, ops.MATH_FP32)
    ops.corr_fwd(ops.Slab(f1, 0, c), ops.Slab(f2, 0, c), ops.Slab(o1, 0, D * D), md, s2, ops.ACT_ELU, ops.MATH_TF32)
    torch.cuda.synchronize()
    assert rel(o1[..., :D * D], o0[..., :D * D]) < TOL
    assert float((o1[..., D * D:] - 5.0).abs().max()) == 0.0          # nothing written past the D*D channels
    from oracle import flownet_c, tf_ops
    want = tf_ops.elu(flownet_c.correlation(f1.cpu().double(), f2.cpu().double(), md, s2))
    assert rel(o1[..., :D * D], want) < TOL                            # the (paper-derived) CPU restatement of the cost volume


@pytest.mark.parametrize("B,h,w,md,s2", [(2, 12, 64, 20, 2), (1, 7, 40, 20, 2), (2, 6, 128, 20, 2), (1, 5, 64, 16, 4), (1, 48, 64, 20, 2)])
def test_tc_correlation_backward_matches_simt(B, h, w, md, s2):
    from deepof_b200 import ops
    g = torch.Generator().manual_seed(B * h + w + 1)
    c = 256
    f1 = _buf(B, h, w, c, c, g)
    f2 = _buf(B, h, w, c, c, g)
    D = 2 * (md // s2) + 1
    dout = _buf(B, h, w, D * D + 7, D * D, g)
    outs = []
    for math_mode in (ops.MATH_FP32, ops.MATH_TF32):
        d1 = torch.full((B, h, w, c), 3.0, device="cuda")
        d2 = torch.full((B, h, w, c), 3.0, device="cuda")
        ops.corr_bwd(ops.Slab(f1, 0, c), ops.Slab(f2, 0, c), ops.Slab(dout, 0, D * D), ops.Slab(d1, 0, c), ops.Slab(d2, 0, c), md, s2, math_mode)
        outs.append((d1, d2))
    torch.cuda.synchronize()
    assert rel(outs[1][0], outs[0][0]) < TOL
    assert rel(outs[1][1], outs[0][1]) < TOL


def _shadow(t):
    return t.to(torch.bfloat16).contiguous()


BF_TOL = 2e-2     # bf16 operands (8-bit mantissa), fp32 accumulate: max-norm relative


@pytest.mark.parametrize("B,H,W,ci", [(2, 64, 128, 6), (1, 384, 512, 6), (3, 48, 80, 3)])
def test_bf16_conv1_fwd_and_wgrad_match_simt(B, H, W, ci):
    """First layer with bf16 operands (bf16 copy of the zero-bordered input, one K block per filter row) vs the fp32 SIMT kernel;
    ci = 3 is FlowNetC's siamese first layer."""
    from deepof_b200 import ops
    g = torch.Generator().manual_seed(B * H + W + ci)
    co, k = 64, 7
    img = torch.randn(B, H, W, ci, generator=g).cuda()
    dense = torch.zeros(B, H, W, 8, device="cuda")
    dense[..., :ci] = img
    padded = torch.zeros(B, H + 6, W + 8, 8, device="cuda")
    padded[:, 2:2 + H, 2:2 + W, :ci] = img
    padded16 = torch.zeros(padded.shape, dtype=torch.bfloat16, device="cuda")
    ops.cast_bf16_raw(padded, padded16, 8)
    torch.cuda.synchronize()
    assert torch.equal(padded16, padded.to(torch.bfloat16))
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, 2)
    y0 = torch.zeros(B, geom.oh, geom.ow, 128, device="cuda")
    y1 = torch.zeros_like(y0)
    y1s = torch.zeros(y1.shape, dtype=torch.bfloat16, device="cuda")
    ops.conv_fwd(geom, ops.Slab(dense, 0, ci), w, b, ops.Slab(y0, 0, co), ops.ACT_ELU, ops.MATH_FP32)
    ops.conv1_fwd(geom, padded, (2, 2), w, b, ops.Slab(y1, 0, co, y1s), ops.ACT_ELU, padded16)
    torch.cuda.synchronize()
    assert rel(y1, y0) < BF_TOL
    assert torch.equal(y1s[..., :co], y1[..., :co].to(torch.bfloat16))
    dy = _buf(B, geom.oh, geom.ow, 128, co, g)
    dw0 = torch.zeros(k, k, ci, co, device="cuda"); dw1 = torch.zeros_like(dw0)
    ops.conv_wgrad(geom, ops.Slab(dense, 0, ci), ops.Slab(dy, 0, co), dw0, None, ops.MATH_FP32)
    ops.conv1_wgrad(geom, padded, (2, 2), ops.Slab(dy, 0, co, _shadow(dy)), dw1, None, padded16)
    torch.cuda.synchronize()
    assert rel(dw1, dw0) < BF_TOL


@pytest.mark.parametrize("B,h,w,c,ld", [(4, 6, 8, 1024, 1024), (2, 12, 16, 1026, 1088), (3, 24, 32, 770, 832)])
def test_head_fwd_through_the_gather_gemm(B, h, w, c, ld):
    """Coarse-scale flow heads (3x3 conv to 2 channels, output pitch 2) through the tensor-core gather-GEMM vs the SIMT head kernel."""
    from deepof_b200 import ops
    g = torch.Generator().manual_seed(B + h + c)
    x = _buf(B, h, w, ld, c, g)
    wt = (torch.randn(3, 3, c, 2, generator=g) / math.sqrt(9 * c)).cuda()
    b = (torch.randn(2, generator=g) * 0.1).cuda()
    pr0 = torch.zeros(B, h, w, 2, device="cuda")
    ops.head_fwd(ops.Slab(x, 0, c), wt, b, pr0)
    geom = ops.conv_geom(B, h, w, c, 2, 3, 1)
    for mth, tol in ((ops.MATH_TF32, TOL), (ops.MATH_BF16, BF_TOL)):
        pr1 = torch.full_like(pr0, 7.0)
        ops.conv_fwd(geom, ops.Slab(x, 0, c, _shadow(x)), wt, b, ops.full(pr1), ops.ACT_NONE, mth)
        torch.cuda.synchronize()
        assert rel(pr1, pr0) < tol, mth


def test_conv_fwd_accumulate_flag():
    """ACT_ACCUMULATE: y += conv(x) + bias in all three math modes (used where a transposed conv's input gradient lands second)."""
    from deepof_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, H, W, ci, co, k, s = 2, 12, 16, 64, 128, 3, 1
    x = _buf(B, H, W, 64, ci, g)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    ref = torch.zeros(B, geom.oh, geom.ow, 128, device="cuda")
    ops.conv_fwd(geom, ops.Slab(x, 0, ci), w, None, ops.Slab(ref, 0, co), ops.ACT_NONE, ops.MATH_FP32)
    for mth, tol in ((ops.MATH_FP32, 1e-5), (ops.MATH_TF32, TOL), (ops.MATH_BF16, BF_TOL)):
        y = torch.full_like(ref, 0.5)
        ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x)), w, None, ops.Slab(y, 0, co), ops.ACT_NONE | ops.ACT_ACCUMULATE, mth)
        torch.cuda.synchronize()
        assert rel(y[..., :co] - 0.5, ref[..., :co]) < tol, mth


@pytest.mark.parametrize("case", [
    (2, 24, 32, 64, 128, 128, 5, 2), (2, 12, 16, 256, 256, 256, 3, 1), (4, 12, 16, 256, 448, 512, 3, 2), (8, 3, 4, 1024, 1024, 1024, 3, 1),
    (1, 48, 64, 128, 256, 256, 5, 2), (2, 10, 14, 32, 64, 32, 3, 1), (2, 16, 16, 96, 128, 64, 3, 1)])
def test_bf16_conv_fwd_dgrad_wgrad_match_simt(case):
    from deepof_b200 import ops
    B, H, W, ci, x_ld, co, k, s = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = _buf(B, H, W, x_ld, ci, g)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    y_ld = (co + 63) // 64 * 64 + 64
    y0 = torch.zeros(B, geom.oh, geom.ow, y_ld, device="cuda")
    y1 = torch.zeros_like(y0)
    y1s = torch.zeros(y1.shape, dtype=torch.bfloat16, device="cuda")
    ops.conv_fwd(geom, ops.Slab(x, 0, ci), w, b, ops.Slab(y0, 64, co), ops.ACT_ELU, ops.MATH_FP32)
    ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x)), w, b, ops.Slab(y1, 64, co, y1s), ops.ACT_ELU, ops.MATH_BF16)
    torch.cuda.synchronize()
    assert rel(y1, y0) < BF_TOL
    assert torch.equal(y1s[..., 64:64 + co], y1[..., 64:64 + co].to(torch.bfloat16))          # fused shadow == rounded fp32 output
    assert float(y1s[..., :64].abs().max()) == 0.0
    dyl = (co + 63) // 64 * 64
    dy = _buf(B, geom.oh, geom.ow, dyl, co, g)
    d0 = torch.full((B, H, W, x_ld), 0.25, device="cuda")
    d1 = torch.full((B, H, W, x_ld), 0.25, device="cuda")
    for acc in (True, False):
        ops.conv_dgrad(geom, ops.Slab(dy, 0, co), w, None, ops.Slab(d0, 0, ci), ops.ACT_NONE, acc, ops.MATH_FP32)
        ops.conv_dgrad(geom, ops.Slab(dy, 0, co, _shadow(dy)), w, None, ops.Slab(d1, 0, ci), ops.ACT_NONE, acc, ops.MATH_BF16)
        torch.cuda.synchronize()
        assert rel(d1[..., :ci], d0[..., :ci]) < BF_TOL, acc
    dw0 = torch.zeros(k, k, ci, co, device="cuda"); dw1 = torch.zeros_like(dw0)
    ops.conv_wgrad(geom, ops.Slab(x, 0, ci), ops.Slab(dy, 0, co), dw0, None, ops.MATH_FP32)
    ops.conv_wgrad(geom, ops.Slab(x, 0, ci, _shadow(x)), ops.Slab(dy, 0, co, _shadow(dy)), dw1, None, ops.MATH_BF16)
    torch.cuda.synchronize()
    assert rel(dw1, dw0) < BF_TOL


@pytest.mark.parametrize("case", [(2, 6, 8, 1024, 1024, 512), (2, 12, 16, 1026, 1088, 256), (1, 48, 64, 386, 448, 64), (2, 24, 32, 194, 256, 32)])
def test_bf16_deconv_match_simt(case):
    from deepof_b200 import ops
    B, h, w, cfeat, fld, upc = case
    g = torch.Generator().manual_seed(sum(case) + 2)
    x = _buf(B, h, w, fld, cfeat, g)
    wt = (torch.randn(4, 4, upc, cfeat, generator=g) / math.sqrt(4 * cfeat)).cuda()
    b = (torch.randn(upc, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, 2 * h, 2 * w, upc, cfeat, 4, 2)
    y0 = torch.zeros(B, 2 * h, 2 * w, 128, device="cuda")
    y1 = torch.zeros_like(y0)
    y1s = torch.zeros(y1.shape, dtype=torch.bfloat16, device="cuda")
    ld_out = 64 + (upc + 63) // 64 * 64
    y0 = torch.zeros(B, 2 * h, 2 * w, ld_out, device="cuda"); y1 = torch.zeros_like(y0)
    y1s = torch.zeros(y1.shape, dtype=torch.bfloat16, device="cuda")
    ops.conv_dgrad(geom, ops.Slab(x, 0, cfeat), wt, b, ops.Slab(y0, 64, upc), ops.ACT_ELU, False, ops.MATH_FP32)
    ops.conv_dgrad(geom, ops.Slab(x, 0, cfeat, _shadow(x)), wt, b, ops.Slab(y1, 64, upc, y1s), ops.ACT_ELU, False, ops.MATH_BF16)
    torch.cuda.synchronize()
    assert rel(y1, y0) < BF_TOL
    assert torch.equal(y1s[..., 64:64 + upc], y1[..., 64:64 + upc].to(torch.bfloat16))
    dy = _buf(B, 2 * h, 2 * w, (upc + 63) // 64 * 64, upc, g)
    d0 = torch.zeros(B, h, w, fld, device="cuda"); d1 = torch.zeros_like(d0)
    ops.conv_fwd(geom, ops.Slab(dy, 0, upc), wt, None, ops.Slab(d0, 0, cfeat), ops.ACT_NONE, ops.MATH_FP32)
    ops.conv_fwd(geom, ops.Slab(dy, 0, upc, _shadow(dy)), wt, None, ops.Slab(d1, 0, cfeat), ops.ACT_NONE, ops.MATH_BF16)
    dw0 = torch.zeros(4, 4, upc, cfeat, device="cuda"); dw1 = torch.zeros_like(dw0)
    ops.conv_wgrad(geom, ops.Slab(dy, 0, upc), ops.Slab(x, 0, cfeat), dw0, None, ops.MATH_FP32)
    ops.conv_wgrad(geom, ops.Slab(dy, 0, upc, _shadow(dy)), ops.Slab(x, 0, cfeat, _shadow(x)), dw1, None, ops.MATH_BF16)
    torch.cuda.synchronize()
    assert rel(d1[..., :cfeat], d0[..., :cfeat]) < BF_TOL
    assert rel(dw1, dw0) < BF_TOL


def test_bf16_engine_epe_within_tolerance():
    """BF16 tensor-core math (bf16 operands, fp32 accumulate/epilogue): EPE on the held-out batch within 1e-3 of the fp32 path."""
    from deepof_b200.flownet import FlowNetS
    from oracle import synth, metrics
    B, H, W = 2, 384, 512
    src, tgt, gt = synth.make_pairs(B, H, W, seed=1234)
    e32 = FlowNetS(B, H, W, seed=1, math_mode="fp32")
    ebf = FlowNetS(B, H, W, seed=1, math_mode="bf16")
    for e in (e32, ebf):
        e.forward(src.cuda(), tgt.cuda())
        e.backward()
    torch.cuda.synchronize()
    d = (ebf.pr[1] - e32.pr[1]).abs() * 10.0
    epe32 = metrics.flow_ee(metrics.eval_flow(e32.pr[1].cpu() * 10.0, H, W), gt).item()
    epebf = metrics.flow_ee(metrics.eval_flow(ebf.pr[1].cpu() * 10.0, H, W), gt).item()
    cos = torch.nn.functional.cosine_similarity(ebf.grad.double(), e32.grad.double(), dim=0).item()
    print(f"bf16 vs fp32: mean|dflow1|={d.mean().item():.3e} max={d.max().item():.3e} EPE fp32={epe32:.6f} bf16={epebf:.6f} grad cosine={cos:.4f}")
    assert abs(epebf - epe32) < 1e-3
    assert d.mean().item() < 1e-2
    assert torch.allclose(ebf.loss4, e32.loss4, rtol=2e-2, atol=1e-3)
    assert cos > 0.9
    for _ in range(2):
        ebf.train_step(src.cuda(), tgt.cuda(), lr=1.6e-5)
    assert torch.isfinite(ebf.theta).all()


def test_tf32_engine_tracks_fp32_engine_and_oracle_epe():
    """Whole step in TF32 mode: flows within the stated tolerance of the fp32 device path and EPE within 1e-3 of the CPU oracle."""
    from deepof_b200.flownet import FlowNetS
    from oracle import flownet_s as fs, synth, metrics
    B, H, W = 2, 384, 512
    src, tgt, gt = synth.make_pairs(B, H, W, seed=1234)
    e32 = FlowNetS(B, H, W, seed=1, math_mode="fp32")
    etf = FlowNetS(B, H, W, seed=1, math_mode="tf32", tc_wgrad=True)
    for e in (e32, etf):
        e.forward(src.cuda(), tgt.cuda())
        e.backward()
    torch.cuda.synchronize()
    l1 = (etf.pr[1] - e32.pr[1]).abs().mean().item() * 10.0
    mx = (etf.pr[1] - e32.pr[1]).abs().max().item() * 10.0
    epe32 = metrics.flow_ee(metrics.eval_flow(e32.pr[1].cpu() * 10.0, H, W), gt).item()
    epetf = metrics.flow_ee(metrics.eval_flow(etf.pr[1].cpu() * 10.0, H, W), gt).item()
    print(f"tf32 vs fp32: mean|dflow1|={l1:.3e} max={mx:.3e} EPE fp32={epe32:.6f} tf32={epetf:.6f}")
    assert abs(epetf - epe32) < 1e-3
    assert torch.allclose(etf.loss4, e32.loss4, rtol=5e-3, atol=1e-4)
    # Gradients: every kernel is within 3e-3 of fp32 (tests above), but the Charbonnier loss (alpha=0.25, eps=1e-4) is
    # extremely ill-conditioned around zero residuals, so a 1e-4 px change in the flow moves individual gradient entries
    # a lot.  The step direction must still agree: cosine similarity of the whole 38.8M-element gradient.
    cos = torch.nn.functional.cosine_similarity(etf.grad.double(), e32.grad.double(), dim=0).item()
    worst = max(rel(etf.grads[n], e32.grads[n]) for n in e32.grads)
    print(f"tf32 vs fp32 gradient: cosine={cos:.6f} worst per-tensor max-norm deviation={worst:.3e}")
    # (0.9866 with every K loop in one piece; at this tiny batch every layer is "coarse" and runs split-K, whose partial sums meet in a
    # run-dependent order: measured 0.973 .. 0.982 -- scripts/splitk_check.py shows the forward flows moving by 2e-5 .. 1e-3 relative, the
    # same order as TF32's own truncation, and the Charbonnier gradient amplifying it)
    assert cos > 0.95

# ---- CTA pairs (cta_group::2) vs single-CTA tiles: same K order per tile, so forward / input gradients must be bit-identical --------

@pytest.fixture
def cta_pairs():
    from deepof_b200 import _lib
    lib = _lib.load()
    yield lib
    lib.dofb_enable_cta_pairs(1)        # (the engine's default)


PAIR_CASES = [(32, 48, 64, 256, 256, 3, 1), (31, 12, 16, 512, 512, 3, 1),      # odd tile count: phantom second tile
              (32, 48, 64, 256, 512, 3, 2), (16, 24, 32, 512, 256, 4, 2), (32, 6, 8, 1024, 1024, 3, 1)]


@pytest.mark.parametrize("case", PAIR_CASES)
@pytest.mark.parametrize("mth", [1, 2])
def test_cta_pairs_fwd_and_dgrad_bit_identical(case, mth, cta_pairs):
    from deepof_b200 import ops
    B, H, W, ci, co, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    x = _buf(B, H, W, (ci + 63) // 64 * 64, ci, g)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    yl = (co + 63) // 64 * 64
    dy = _buf(B, geom.oh, geom.ow, yl, co, g)
    ys, ds = [], []
    cta_pairs.dofb_enable_split_k(0)          # (split-K partial sums meet through atomics in a run-dependent order: not bit-stable)
    try:
        for pairs in (0, 1):
            cta_pairs.dofb_enable_cta_pairs(pairs)
            y = torch.zeros(B, geom.oh, geom.ow, yl, device="cuda")
            ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x)), w, b, ops.Slab(y, 0, co), ops.ACT_ELU, mth)
            d = torch.full((B, H, W, x.shape[3]), 0.5, device="cuda")
            ops.conv_dgrad(geom, ops.Slab(dy, 0, co, _shadow(dy)), w, None, ops.Slab(d, 0, ci), ops.ACT_NONE, True, mth)   # (strided: all phases, one launch)
            torch.cuda.synchronize()
            ys.append(y); ds.append(d)
    finally:
        cta_pairs.dofb_enable_split_k(1)
    assert torch.equal(ys[0], ys[1])
    assert torch.equal(ds[0], ds[1])
    ref = torch.zeros_like(ys[0])
    ops.conv_fwd(geom, ops.Slab(x, 0, ci), w, b, ops.Slab(ref, 0, co), ops.ACT_ELU, ops.MATH_FP32)
    assert rel(ys[1], ref) < (BF_TOL if mth == 2 else TOL)


@pytest.mark.parametrize("case", [(8, 48, 64, 256, 256, 3, 1), (8, 48, 64, 128, 256, 5, 2), (8, 96, 128, 64, 128, 5, 2), (8, 96, 128, 32, 194, 4, 2),
                                  (8, 48, 64, 256, 130, 3, 1), (3, 6, 8, 1024, 1024, 3, 1)])
@pytest.mark.parametrize("mth", [1, 2])
def test_cta_pairs_wgrad_matches_single_cta(case, mth, cta_pairs):
    """Weight gradient: pairs of A tiles sharing a dy tile (odd tap counts leave a phantom partner); fp32 atomics -> order-level differences."""
    from deepof_b200 import ops
    B, H, W, ci, co, k, s = case
    g = torch.Generator().manual_seed(sum(case) + 7)
    x = _buf(B, H, W, (ci + 63) // 64 * 64, ci, g)
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    dy = _buf(B, geom.oh, geom.ow, (co + 63) // 64 * 64, co, g)
    outs = []
    for pairs in (0, 1):
        cta_pairs.dofb_enable_cta_pairs(pairs)
        dw = torch.zeros(k, k, ci, co, device="cuda")
        ops.conv_wgrad(geom, ops.Slab(x, 0, ci, _shadow(x)), ops.Slab(dy, 0, co, _shadow(dy)), dw, None, mth)
        torch.cuda.synchronize()
        outs.append(dw)
    assert rel(outs[1], outs[0]) < 1e-5
    ref = torch.zeros_like(outs[0])
    ops.conv_wgrad(geom, ops.Slab(x, 0, ci), ops.Slab(dy, 0, co), ref, None, ops.MATH_FP32)
    assert rel(outs[1], ref) < (BF_TOL if mth == 2 else TOL)


def test_pack_weights_batch_matches_lazy_packing():
    """dofb_pack_weights_batch fills the cache the convolutions read: results must equal the lazily packed path bit for bit."""
    from deepof_b200 import ops, _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 12, 16
    specs = [(64, 128, 3, 1), (130, 96, 5, 2), (256, 256, 3, 1), (32, 194, 4, 2)]
    ws = [(torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda() for ci, co, k, s in specs]
    xs = [_buf(B, H, W, (ci + 63) // 64 * 64, ci, g) for ci, co, k, s in specs]

    def run_all(mth):
        outs = []
        for (ci, co, k, s), w, x in zip(specs, ws, xs):
            geom = ops.conv_geom(B, H, W, ci, co, k, s)
            y = torch.zeros(B, geom.oh, geom.ow, (co + 63) // 64 * 64, device="cuda")
            ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x)), w, None, ops.Slab(y, 0, co), ops.ACT_NONE, mth)
            d = torch.zeros_like(x)
            ops.conv_dgrad(geom, ops.Slab(y, 0, co, _shadow(y)), w, None, ops.Slab(d, 0, ci), ops.ACT_NONE, False, mth)
            outs += [y, d]
        torch.cuda.synchronize()
        return outs
    lib.dofb_enable_split_k(0)                # bit-for-bit comparison: no atomically summed K ranges
    for mth in (ops.MATH_TF32, ops.MATH_BF16):
        lib.dofb_enable_weight_cache(0)
        lazy = run_all(mth)
        lib.dofb_enable_weight_cache(1)                 # (bumps the epoch: nothing is fresh)
        jobs = ops.make_pack_jobs([(w, c) for w in ws for c in (1, 0)])
        ops.pack_weights_batch(jobs, mth == ops.MATH_BF16)
        before = lib.dofb_launch_count()
        batched = run_all(mth)
        assert lib.dofb_launch_count() - before == 2 * len(specs)          # no pack launches: one fwd + one (merged-phase) dgrad per layer
        lib.dofb_enable_weight_cache(0)
        for a, b_ in zip(lazy, batched):
            assert torch.equal(a, b_)
    lib.dofb_enable_split_k(1)


@pytest.mark.parametrize("which,case", [("dgrad", (4, 96, 128, 32, 194, 4, 2)), ("dgrad", (2, 64, 96, 64, 386, 4, 2)), ("dgrad", (2, 48, 64, 128, 130, 4, 2)),
                                        ("dgrad", (2, 64, 48, 64, 128, 3, 2)), ("dgrad", (3, 34, 22, 96, 64, 2, 1)), ("fwd", (2, 40, 24, 64, 128, 2, 1)),
                                        ("fwd", (4, 48, 64, 64, 128, 3, 1)), ("dgrad", (2, 96, 128, 64, 128, 5, 2))])   # (last two: > 4 taps, per-tap path)
@pytest.mark.parametrize("mth", [1, 2])
def test_halo_tiles_match_per_tap_gather(which, case, mth):
    """Opt-in halo path (one TMA box per tile and channel block, MMA descriptors offset by whole rows of the halo; phases of <= 4 taps:
    the 4x4 / stride-2 transposed convs and 3x3 / stride-2 input gradients) vs the per-tap gather;
    only the K order differs (channel block outer), so the results agree to fp32 summation order."""
    from deepof_b200 import ops, _lib
    lib = _lib.load()
    B, H, W, ci, co, k, s = case
    g = torch.Generator().manual_seed(sum(case) + 3)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    xl, yl = (ci + 63) // 64 * 64, (co + 63) // 64 * 64
    outs = []
    try:
        for halo in (0, 1):
            lib.dofb_enable_halo_tiles(halo)
            if which == "fwd":
                x = _buf(B, H, W, xl, ci, torch.Generator().manual_seed(5))
                y = torch.zeros(B, geom.oh, geom.ow, yl, device="cuda")
                ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x)), w, None, ops.Slab(y, 0, co), ops.ACT_ELU, mth)
                outs.append(y)
            else:
                dy = _buf(B, geom.oh, geom.ow, yl, co, torch.Generator().manual_seed(6))
                d = torch.zeros(B, H, W, xl, device="cuda")
                ops.conv_dgrad(geom, ops.Slab(dy, 0, co, _shadow(dy)), w, None, ops.Slab(d, 0, ci), ops.ACT_NONE, False, mth)
                outs.append(d)
            torch.cuda.synchronize()
    finally:
        lib.dofb_enable_halo_tiles(0)
    assert rel(outs[1], outs[0]) < 1e-5


@pytest.mark.parametrize("B,h,w,md,s2", [(2, 12, 64, 20, 2), (1, 7, 40, 20, 2), (2, 6, 128, 20, 2), (1, 5, 64, 16, 4), (1, 48, 64, 20, 2), (3, 9, 70, 8, 1)])
def test_bf16_correlation_backward_matches_oracle(B, h, w, md, s2):
    """tc_corr_bwd16_kernel (bf16 band-GEMMs, eight generator warps) against autograd of the CPU restatement of the cost volume."""
    from deepof_b200 import ops
    from oracle import flownet_c
    g = torch.Generator().manual_seed(B * h + w + 5)
    c = 256
    f1 = _buf(B, h, w, c, c, g).to(torch.bfloat16)
    f2 = _buf(B, h, w, c, c, g).to(torch.bfloat16)
    D = 2 * (md // s2) + 1
    dout = _buf(B, h, w, D * D + 7, D * D, g)
    d1 = torch.full((B, h, w, c), 3.0, device="cuda")
    d2 = torch.full((B, h, w, c), 3.0, device="cuda")
    dummy = torch.zeros(B, h, w, c, device="cuda")
    ops.corr_bwd(ops.Slab(dummy, 0, c, f1), ops.Slab(dummy, 0, c, f2), ops.Slab(dout, 0, D * D), ops.Slab(d1, 0, c), ops.Slab(d2, 0, c), md, s2, ops.MATH_BF16)
    torch.cuda.synchronize()
    a = f1.float().cpu().double().requires_grad_(True)
    b = f2.float().cpu().double().requires_grad_(True)
    # the kernel multiplies the bf16 rounding of dout (the band matrix is generated in bf16)
    flownet_c.correlation(a, b, md, s2).backward(dout.cpu()[..., :D * D].to(torch.bfloat16).double())
    assert rel(d1, a.grad) < 2e-5
    assert rel(d2, b.grad) < 2e-5


@pytest.mark.parametrize("case", [(2, 24, 32, 64, 128, 128, 5, 2), (1, 48, 64, 64, 64, 128, 3, 1), (2, 40, 56, 48, 64, 96, 5, 2)])
def test_wgrad_tap_packing_on_n_matches_m_side_packing(case):
    """conv2-shaped bf16 weight gradients: 'dy on M, four taps of x on N' (full 128 x 256 MMAs) against the default 'two taps on M' form
    and the CPU oracle."""
    from deepof_b200 import ops, _lib
    from oracle import tf_ops
    B, H, W, ci, x_ld, co, k, s = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(sum(case))
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    x = _buf(B, H, W, x_ld, ci, g).to(torch.bfloat16)
    dy = _buf(B, geom.oh, geom.ow, (co + 63) // 64 * 64, co, g).to(torch.bfloat16)
    dummy_x, dummy_dy = torch.zeros(x.shape, device="cuda"), torch.zeros(dy.shape, device="cuda")
    outs = []
    for on in (0, 1):
        lib.dofb_enable_wgrad_npack(on)
        dw = torch.zeros(k, k, ci, co, device="cuda")
        ops.conv_wgrad(geom, ops.Slab(dummy_x, 0, ci, x), ops.Slab(dummy_dy, 0, co, dy), dw, None, ops.MATH_BF16)
        torch.cuda.synchronize()
        outs.append(dw)
    lib.dofb_enable_wgrad_npack(1)
    assert rel(outs[1], outs[0]) < 2e-5
    wd = torch.zeros(k, k, ci, co, dtype=torch.float64, requires_grad=True)
    tf_ops.conv2d_same(x.float().cpu()[..., :ci].double(), wd, None, s).backward(dy.float().cpu()[..., :co].double())
    assert rel(outs[1], wd.grad) < 2e-5


@pytest.mark.parametrize("mth", ["tf32", "bf16"])
@pytest.mark.parametrize("case", [
    # (kind, B, h, w, c_contract, ld, c_out, k): deconv = 4x4/2 conv2d_transpose forward on an h x w map; dgrad = k x k / 2 conv input gradient
    ("deconv", 2, 24, 32, 194, 256, 32, 4), ("deconv", 1, 48, 64, 386, 448, 64, 4), ("deconv", 2, 12, 16, 770, 832, 128, 4), ("deconv", 1, 9, 11, 100, 128, 64, 4),
    ("dgrad", 2, 24, 32, 128, 128, 64, 5), ("dgrad", 1, 24, 32, 256, 256, 128, 5), ("dgrad", 2, 10, 12, 96, 128, 32, 3), ("dgrad", 1, 13, 9, 64, 64, 64, 7)])
def test_phase_in_n_matches_per_phase_gather(case, mth):
    """Stride-2 transposed gathers with 32 / 64 / 128 output channels: the four output phases side by side on N over the distinct source offsets
    (dofb_enable_phase_in_n) against the per-phase / per-tap form -- same products, so equal up to fp32 summation order -- for
    overwrite + bias + ELU + bf16 shadow (deconv forward) and for accumulation into an existing gradient (conv dgrad)."""
    from deepof_b200 import ops, _lib
    kind, B, h, w, cc, ld, cout, k = case
    lib = _lib.load()
    m = ops.MATH_BF16 if mth == "bf16" else ops.MATH_TF32
    g = torch.Generator().manual_seed(sum(case[1:]) + 9)
    x = _buf(B, h, w, ld, cc, g)
    xs = _shadow(x) if mth == "bf16" else None
    wt = (torch.randn(k, k, cout, cc, generator=g) / math.sqrt(k * k * cc / 4)).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, 2 * h, 2 * w, cout, cc, k, 2)
    assert (geom.oh, geom.ow) == (h, w)
    ld_out = 128 + (cout + 63) // 64 * 64         # the slab sits between 64 foreign channels on either side
    outs = []
    try:
        for on in (0, 2):
            lib.dofb_enable_phase_in_n(on)
            if kind == "deconv":
                y = torch.zeros(B, 2 * h, 2 * w, ld_out, device="cuda")
                ys = torch.zeros(y.shape, dtype=torch.bfloat16, device="cuda") if mth == "bf16" else None
                ops.conv_dgrad(geom, ops.Slab(x, 0, cc, xs), wt, b, ops.Slab(y, 64, cout, ys), ops.ACT_ELU, False, m)
                torch.cuda.synchronize()
                if ys is not None:
                    assert torch.equal(ys[..., 64:64 + cout], y[..., 64:64 + cout].to(torch.bfloat16))
                    assert float(ys[..., :64].abs().max()) == 0.0 and float(ys[..., 64 + cout:].abs().max()) == 0.0
                assert float(y[..., :64].abs().max()) == 0.0 and float(y[..., 64 + cout:].abs().max()) == 0.0
            else:
                y = torch.full((B, 2 * h, 2 * w, ld_out), 0.25, device="cuda")
                ops.conv_dgrad(geom, ops.Slab(x, 0, cc, xs), wt, None, ops.Slab(y, 64, cout), ops.ACT_NONE, True, m)
                torch.cuda.synchronize()
                assert float((y[..., :64] - 0.25).abs().max()) == 0.0 and float((y[..., 64 + cout:] - 0.25).abs().max()) == 0.0
                y = y - 0.25
            outs.append(y[..., 64:64 + cout].clone())
    finally:
        lib.dofb_enable_phase_in_n(1)
    assert rel(outs[1], outs[0]) < 1e-5
    # and against the fp32 SIMT path (itself checked against the oracle elsewhere)
    r = torch.zeros(B, 2 * h, 2 * w, ld_out, device="cuda")
    ops.conv_dgrad(geom, ops.Slab(x, 0, cc), wt, b if kind == "deconv" else None, ops.Slab(r, 64, cout), ops.ACT_ELU if kind == "deconv" else ops.ACT_NONE, False, ops.MATH_FP32)
    torch.cuda.synchronize()
    assert rel(outs[1], r[..., 64:64 + cout]) < (BF_TOL if mth == "bf16" else TOL)


@pytest.mark.parametrize("mth", ["tf32", "bf16"])
@pytest.mark.parametrize("pairs", [0, 1])
@pytest.mark.parametrize("case", [(4, 6, 8, 512, 512, 1024, 3, 2), (8, 3, 4, 1024, 1024, 1024, 3, 1), (2, 12, 16, 256, 320, 512, 3, 1), (3, 6, 8, 300, 320, 260, 3, 1)])
@pytest.mark.parametrize("ks", [2, 3, 5])
def test_split_k_matches_unsplit(case, ks, pairs, mth):
    """Coarse 256-column layers: K loop cut into ks ranges with atomic partial sums + finish pass (dofb_enable_split_k) against the unsplit
    launch, for forward (bias + ELU, fp32 and bf16-only outputs), input gradient overwriting and accumulating (all stride phases)."""
    from deepof_b200 import ops, _lib
    B, H, W, ci, x_ld, co, k, s = case
    lib = _lib.load()
    m = ops.MATH_BF16 if mth == "bf16" else ops.MATH_TF32
    bf = mth == "bf16"
    g = torch.Generator().manual_seed(sum(case) + 17)
    x = _buf(B, H, W, x_ld, ci, g)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    y_ld = (co + 63) // 64 * 64 + 128
    dy = _buf(B, geom.oh, geom.ow, (co + 63) // 64 * 64, co, g)
    res = []
    lib.dofb_enable_cta_pairs(pairs)
    try:
        for on in (0, ks):
            lib.dofb_enable_split_k(on)
            y = torch.zeros(B, geom.oh, geom.ow, y_ld, device="cuda")
            ys = torch.zeros(y.shape, dtype=torch.bfloat16, device="cuda") if bf else None
            ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x) if bf else None), w, b, ops.Slab(y, 64, co, ys), ops.ACT_ELU, m)
            out = [y[..., 64:64 + co].clone()]
            torch.cuda.synchronize()
            assert float(y[..., :64].abs().max()) == 0.0 and float(y[..., 64 + co:].abs().max()) == 0.0
            if bf:
                assert torch.equal(ys[..., 64:64 + co], y[..., 64:64 + co].to(torch.bfloat16))
                y2s = torch.zeros(y.shape, dtype=torch.bfloat16, device="cuda")       # bf16-only output (lean schedule)
                ops.conv_fwd(geom, ops.Slab(x, 0, ci, _shadow(x)), w, b, ops.Slab(None, 64, co, y2s), ops.ACT_ELU, m)
                torch.cuda.synchronize()
                if on == 0:
                    assert torch.equal(y2s[..., 64:64 + co], ys[..., 64:64 + co])
                else:           # atomically summed: the fp32 sums differ in the last bits, the bf16 roundings by at most one ulp
                    assert rel(y2s[..., 64:64 + co].float(), ys[..., 64:64 + co].float()) < 8e-3
                assert float(y2s[..., :64].abs().max()) == 0.0
            for acc in (False, True):
                d = torch.full((B, H, W, x_ld), 0.25, device="cuda")
                ops.conv_dgrad(geom, ops.Slab(dy, 0, co, _shadow(dy) if bf else None), w, None, ops.Slab(d, 0, ci), ops.ACT_NONE, acc, m)
                torch.cuda.synchronize()
                if ci < x_ld:
                    assert float((d[..., ci:] - 0.25).abs().max()) == 0.0
                out.append(d[..., :ci] - (0.25 if acc else 0.0))
            res.append(out)
    finally:
        lib.dofb_enable_split_k(1)
        lib.dofb_enable_cta_pairs(1)
    for a, r in zip(res[1], res[0]):
        assert rel(a, r) < 5e-5           # same products, different fp32 summation order (K ranges meet through atomics)


@pytest.mark.parametrize("B,H,W,ci,co", [(2, 64, 128, 6, 64), (1, 384, 512, 6, 64), (3, 48, 80, 3, 64), (2, 40, 72, 6, 32), (5, 16, 24, 6, 64)])
def test_conv1_four_pixels_per_row_matches_one_pixel_form(B, H, W, ci, co):
    """conv1 forward (bf16) with four x-adjacent output pixels per GEMM row (N = 4 x co, K = the 16-pixel window of a filter row) against
    the one-pixel-per-row form: same products -> equal up to fp32 summation order; bias + ELU + bf16 shadow, pad channels untouched."""
    from deepof_b200 import ops, _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * H + W + ci + co)
    k = 7
    img = torch.randn(B, H, W, ci, generator=g).cuda()
    padded = torch.zeros(B, H + 6, W + 8, 8, device="cuda")
    padded[:, 2:2 + H, 2:2 + W, :ci] = img
    padded16 = padded.to(torch.bfloat16)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, 2)
    outs = []
    try:
        for on in (0, 2):
            lib.dofb_enable_phase_in_n(on)
            y = torch.zeros(B, geom.oh, geom.ow, 128, device="cuda")
            ys = torch.zeros(y.shape, dtype=torch.bfloat16, device="cuda")
            ops.conv1_fwd(geom, padded, (2, 2), w, b, ops.Slab(y, 32, co, ys), ops.ACT_ELU, padded16)
            torch.cuda.synchronize()
            assert torch.equal(ys[..., 32:32 + co], y[..., 32:32 + co].to(torch.bfloat16))
            assert float(y[..., :32].abs().max()) == 0.0 and float(y[..., 32 + co:].abs().max()) == 0.0
            assert float(ys[..., :32].abs().max()) == 0.0 and float(ys[..., 32 + co:].abs().max()) == 0.0
            outs.append(y[..., 32:32 + co].clone())
            if on == 2:                          # bf16-only output (lean schedule)
                y2 = torch.zeros(y.shape, dtype=torch.bfloat16, device="cuda")
                ops.conv1_fwd(geom, padded, (2, 2), w, b, ops.Slab(None, 32, co, y2), ops.ACT_ELU, padded16)
                torch.cuda.synchronize()
                assert torch.equal(y2, ys)
    finally:
        lib.dofb_enable_phase_in_n(1)
    assert rel(outs[1], outs[0]) < 1e-5
    assert float(outs[0].abs().max()) > 0.1


@pytest.mark.parametrize("B,h,w,md,s2", [(2, 12, 64, 20, 2), (1, 7, 40, 20, 2), (2, 6, 128, 20, 2), (1, 5, 64, 16, 4), (1, 48, 64, 20, 2)])
def test_bf16_correlation_forward_matches_oracle(B, h, w, md, s2):
    """tc_corr_fwd_kernel<true> (bf16 maps, kind::f16) against the CPU restatement of the cost volume evaluated on the same bf16-rounded
    maps (fp32 accumulation of exact bf16 products -> tight bound), with the fused bf16 shadow of the output."""
    from deepof_b200 import ops
    from oracle import flownet_c, tf_ops
    g = torch.Generator().manual_seed(B * h + w + 11)
    c = 256
    f1 = _buf(B, h, w, c, c, g).to(torch.bfloat16)
    f2 = _buf(B, h, w, c, c, g).to(torch.bfloat16)
    D = 2 * (md // s2) + 1
    ld = D * D + 7
    o = torch.full((B, h, w, ld), 5.0, device="cuda")
    o16 = torch.full((B, h, w, ld), 5.0, device="cuda").to(torch.bfloat16)
    ops.corr_fwd(ops.Slab(None, 0, c, f1), ops.Slab(None, 0, c, f2), ops.Slab(o, 0, D * D, o16), md, s2, ops.ACT_ELU, ops.MATH_BF16)
    torch.cuda.synchronize()
    want = tf_ops.elu(flownet_c.correlation(f1.float().cpu().double(), f2.float().cpu().double(), md, s2))
    assert rel(o[..., :D * D], want) < 2e-5
    assert float((o[..., D * D:] - 5.0).abs().max()) == 0.0 and float((o16[..., D * D:].float() - 5.0).abs().max()) == 0.0
    assert torch.equal(o16[..., :D * D], o[..., :D * D].to(torch.bfloat16))
