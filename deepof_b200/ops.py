"""Tensor-level wrappers over the C ABI: argument checking + pointer/pitch plumbing.

Every function launches hand-written CUDA from libdeepof_b200.so on the current torch stream.
Tensors are NHWC float32 CUDA tensors; ``Slab`` describes a channel slice of a pitched buffer.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ConvGeom, LossScale, DeepOFError, check

ACT_NONE, ACT_ELU = 0, 1
ACT_ACCUMULATE = 16          # conv_fwd: OR-ed into act, y += conv(x) instead of y = conv(x)
MATH_FP32, MATH_TF32, MATH_BF16 = 0, 1, 2


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DeepOFError(f"{name}: expected a CUDA tensor (the deepof_b200 product path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise DeepOFError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise DeepOFError(f"{name}: expected a contiguous tensor")


@dataclass
class Slab:
    """Channels [c0, c0+c) of a contiguous [B,H,W,ld] buffer (optionally with its bf16 shadow of the same shape).

    ``t`` may be None for a bf16-only buffer (lean bf16 engine: activations that only tensor-core kernels read keep no fp32 copy);
    ``ptr`` is then None and a kernel that needs the fp32 data fails loudly instead of reading zeros."""
    t: torch.Tensor | None
    c0: int
    c: int
    t16: torch.Tensor | None = None

    def __post_init__(self):
        if self.t is not None:
            _req(self.t, "Slab")
        elif self.t16 is None:
            raise DeepOFError("Slab: needs an fp32 buffer or a bf16 buffer")
        ref = self._ref
        assert ref.dim() == 4 and 0 <= self.c0 and self.c0 + self.c <= ref.shape[3]
        if self.t16 is not None:
            assert self.t16.dtype == torch.bfloat16 and self.t16.is_cuda and self.t16.is_contiguous()
            assert self.t is None or self.t16.shape == self.t.shape

    @property
    def _ref(self) -> torch.Tensor:
        return self.t if self.t is not None else self.t16

    @property
    def ptr(self):
        return None if self.t is None else self.t.data_ptr() + 4 * self.c0

    @property
    def ptr16(self):
        return None if self.t16 is None else self.t16.data_ptr() + 2 * self.c0

    @property
    def ld(self) -> int:
        return self._ref.shape[3]

    @property
    def B(self):
        return self._ref.shape[0]

    @property
    def h(self):
        return self._ref.shape[1]

    @property
    def w(self):
        return self._ref.shape[2]

    @property
    def n_pix(self) -> int:
        r = self._ref
        return r.shape[0] * r.shape[1] * r.shape[2]

    def dense(self) -> torch.Tensor:
        return self._ref[..., self.c0:self.c0 + self.c]

    def sub(self, c0: int, c: int) -> "Slab":
        """Channels [c0, c0+c) of this slab."""
        assert 0 <= c0 and c0 + c <= self.c
        return Slab(self.t, self.c0 + c0, c, self.t16)


def full(t: torch.Tensor, c: int | None = None) -> Slab:
    return Slab(t, 0, t.shape[3] if c is None else c)


def same_pad(in_size: int, k: int, stride: int):
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2


def conv_geom(B, ih, iw, ci, co, k, stride) -> ConvGeom:
    """Geometry of a TF-SAME conv (also used, as 'the conv it is the gradient of', for conv2d_transpose)."""
    oh, pt = same_pad(ih, k, stride)
    ow, pl = same_pad(iw, k, stride)
    return ConvGeom(B, ih, iw, ci, oh, ow, co, k, k, stride, pt, pl)


def preprocess(src, tgt, mean, x6, pyr_src, pyr_tgt, origin=(0, 0), x6b=None, divisor=255.0, x6_16=None, x6b_16=None):
    """x6 may be larger than the image (zero border for dofb_conv1_*); ``origin`` = (row, col) of the image in it.
    With ``x6b`` (siamese models) the source goes to x6[..., 0:3] and the target to x6b[..., 0:3] instead of 6 stacked channels."""
    u8 = isinstance(src, torch.Tensor) and src.dtype == torch.uint8
    if u8:          # 8-bit images as the reference's loader returns them (flyingChairsLoader.py:64-80): cast on the device, bit-identical
        if not (isinstance(tgt, torch.Tensor) and tgt.dtype == torch.uint8 and src.is_cuda and tgt.is_cuda and src.is_contiguous() and tgt.is_contiguous()):
            raise DeepOFError("preprocess: uint8 source needs a contiguous uint8 CUDA target as well")
    else:
        _req(src, "src"); _req(tgt, "tgt")
    if x6 is not None:
        _req(x6, "x6")
    B, H, W, _ = src.shape
    n = len(pyr_src)
    for t in list(pyr_src) + list(pyr_tgt):
        _req(t, "pyramid")
    m = (C.c_float * 3)(*[float(v) for v in mean])
    ps = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in pyr_src])
    pt = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in pyr_tgt])
    lib = _lib.load()
    if u8:
        if x6_16 is not None:
            assert x6_16.dtype == torch.bfloat16 and x6_16.shape[3] == 8 and x6_16.is_cuda and x6_16.is_contiguous()
        ref = x6_16 if x6_16 is not None else x6
        xs = ref.shape if ref is not None else (0, 0, 0, 0)
        check(lib.dofb_preprocess_u8(src.data_ptr(), tgt.data_ptr(), m, float(divisor), B, H, W,
                                     x6.data_ptr() if (x6 is not None and x6_16 is None) else None,
                                     x6b.data_ptr() if (x6b is not None and x6_16 is None) else None, xs[3],
                                     x6_16.data_ptr() if x6_16 is not None else None, x6b_16.data_ptr() if x6b_16 is not None else None,
                                     xs[1], xs[2], origin[0], origin[1], n, ps, pt, _stream()))
        return
    if x6_16 is not None:           # bf16 network input only (bf16 first-layer kernels): no fp32 x6 is written
        assert x6_16.dtype == torch.bfloat16 and x6_16.shape[3] == 8 and x6_16.is_cuda and x6_16.is_contiguous()
        check(lib.dofb_preprocess_bf16(src.data_ptr(), tgt.data_ptr(), m, float(divisor), B, H, W, x6_16.data_ptr(),
                                       x6b_16.data_ptr() if x6b_16 is not None else None, x6_16.shape[1], x6_16.shape[2], origin[0], origin[1], n, ps, pt,
                                       _stream()))
        return
    xs = x6.shape if x6 is not None else (0, 0, 0, 0)
    check(lib.dofb_preprocess(src.data_ptr(), tgt.data_ptr(), m, float(divisor), B, H, W, x6.data_ptr() if x6 is not None else None,
                              x6b.data_ptr() if x6b is not None else None, xs[3], xs[1], xs[2], origin[0], origin[1], n, ps, pt, _stream()))


def maxpool2_fwd(x: Slab, y: Slab):
    assert x.c == y.c and x.h == 2 * y.h and x.w == 2 * y.w
    check(_lib.load().dofb_maxpool2_fwd(x.ptr, x.ld, y.B, y.h, y.w, y.c, y.ptr, y.ld, _stream()))


def maxpool2_bwd(x: Slab, dy: Slab, dx: Slab):
    check(_lib.load().dofb_maxpool2_bwd(x.ptr, x.ld, dy.ptr, dy.ld, dy.B, dy.h, dy.w, dy.c, dx.ptr, dx.ld, _stream()))


def conv1_fwd(g: ConvGeom, xpad, origin, w, b, y: Slab, act=ACT_ELU, xpad16=None):
    """conv1 on tensor cores from the zero-bordered input buffer (see dofb_conv1_fwd); xpad16 = its bf16 copy -> bf16 math."""
    _req(xpad, "xpad"); _req(w, "w")
    lib = _lib.load()
    bp = b.data_ptr() if b is not None else None
    if xpad16 is not None:
        check(lib.dofb_conv1_fwd_bf16(C.byref(g), xpad16.data_ptr(), xpad.shape[1], xpad.shape[2], origin[0], origin[1], w.data_ptr(),
                                      bp, y.ptr, y.ptr16, y.ld, act, _stream()))
        return
    check(lib.dofb_conv1_fwd(C.byref(g), xpad.data_ptr(), xpad.shape[1], xpad.shape[2], origin[0], origin[1], w.data_ptr(),
                             bp, y.ptr, y.ptr16, y.ld, act, _stream()))


def conv1_wgrad(g: ConvGeom, xpad, origin, dy: Slab, dw, db, xpad16=None):
    _req(xpad, "xpad"); _req(dw, "dw")
    lib = _lib.load()
    if xpad16 is not None:
        if db is not None:
            raise DeepOFError("conv1_wgrad(bf16): the bias gradient comes from elu_bwd, pass db=None")
        check(lib.dofb_conv1_wgrad_bf16(C.byref(g), xpad16.data_ptr(), xpad.shape[1], xpad.shape[2], origin[0], origin[1],
                                        _need16(dy, "conv1_wgrad"), dy.ld, dw.data_ptr(), _stream()))
        return
    check(lib.dofb_conv1_wgrad(C.byref(g), xpad.data_ptr(), xpad.shape[1], xpad.shape[2], origin[0], origin[1], dy.ptr, dy.ld,
                               dw.data_ptr(), db.data_ptr() if db is not None else None, _stream()))


def cast_bf16_raw(src, dst16, c: int):
    """bf16 copy of a dense [..., c] fp32 tensor (pitch c)."""
    _req(src, "src")
    check(_lib.load().dofb_cast_bf16(src.data_ptr(), c, dst16.data_ptr(), c, src.numel() // c, c, _stream()))


def _need16(s: Slab, who: str):
    if s.ptr16 is None:
        raise DeepOFError(f"{who}: bf16 math needs a bf16 shadow of this operand (Slab.t16)")
    return s.ptr16


def conv_fwd(g: ConvGeom, x: Slab, w, b, y: Slab, act=ACT_ELU, math=MATH_FP32):
    _req(w, "w")
    lib = _lib.load()
    if math == MATH_BF16:
        check(lib.dofb_conv_fwd_bf16(C.byref(g), _need16(x, "conv_fwd"), x.ld, w.data_ptr(), b.data_ptr() if b is not None else None,
                                     y.ptr, y.ptr16, y.ld, act, _stream()))
        return
    check(lib.dofb_conv_fwd(C.byref(g), x.ptr, x.ld, w.data_ptr(), b.data_ptr() if b is not None else None,
                            y.ptr, y.ld, act, math, _stream()))


def conv_dgrad(g: ConvGeom, dy: Slab, w, bias, dx: Slab, act=ACT_NONE, accumulate=False, math=MATH_FP32):
    _req(w, "w")
    lib = _lib.load()
    if math == MATH_BF16:
        check(lib.dofb_conv_dgrad_bf16(C.byref(g), _need16(dy, "conv_dgrad"), dy.ld, w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       dx.ptr, dx.ptr16, dx.ld, act, int(accumulate), _stream()))
        return
    check(lib.dofb_conv_dgrad(C.byref(g), dy.ptr, dy.ld, w.data_ptr(), bias.data_ptr() if bias is not None else None,
                              dx.ptr, dx.ld, act, int(accumulate), math, _stream()))


def conv_wgrad(g: ConvGeom, x: Slab, dy: Slab, dw, db, math=MATH_FP32, bias_on_large=False):
    _req(dw, "dw")
    lib = _lib.load()
    if math == MATH_BF16:
        if db is not None:
            raise DeepOFError("conv_wgrad(bf16): the bias gradient comes from elu_bwd(db=...) on this path")
        check(lib.dofb_conv_wgrad_bf16(C.byref(g), _need16(x, "conv_wgrad"), x.ld, _need16(dy, "conv_wgrad"), dy.ld, dw.data_ptr(), _stream()))
        return
    fn = lib.dofb_conv_wgrad_tbias if bias_on_large else lib.dofb_conv_wgrad
    check(fn(C.byref(g), x.ptr, x.ld, dy.ptr, dy.ld, dw.data_ptr(), db.data_ptr() if db is not None else None, math, _stream()))


def elu_bwd(g: Slab, y: Slab, db=None, shadow_only=False):
    """g *= elu'(y); with db, also db += column sums of the result (the layer's bias gradient) in the same pass.
    shadow_only: write the result to the bf16 shadow of g only (g itself untouched) -- see dofb_elu_bwd_shadow."""
    assert g.c == y.c and g.n_pix == y.n_pix
    dbp = db.data_ptr() if db is not None else None
    if shadow_only:
        if y.ptr16 is not None and os.environ.get("DOFB_ELU_Y32", "0") != "1":      # ELU output from its bf16 shadow: 8 B per element
            check(_lib.load().dofb_elu_bwd_shadow16(g.ptr, g.ld, y.ptr16, y.ld, g.n_pix, g.c, dbp, _need16(g, "elu_bwd"), _stream()))
        else:
            check(_lib.load().dofb_elu_bwd_shadow(g.ptr, g.ld, y.ptr, y.ld, g.n_pix, g.c, dbp, _need16(g, "elu_bwd"), _stream()))
        return
    check(_lib.load().dofb_elu_bwd(g.ptr, g.ld, y.ptr, y.ld, g.n_pix, g.c, dbp, g.ptr16, _stream()))


def cast_bf16(s: Slab):
    """Refresh the bf16 shadow of a slab whose producer has no fused shadow output."""
    check(_lib.load().dofb_cast_bf16(s.ptr, s.ld, _need16(s, "cast_bf16"), s.ld, s.n_pix, s.c, _stream()))


def make_pack_jobs(entries):
    """entries: (weight tensor [kh,kw,ci,co], contract_ci) -> ctypes array for pack_weights_batch (build once, reuse every step)."""
    arr = (_lib.PackJob * len(entries))()
    for i, (w, contract_ci) in enumerate(entries):
        _req(w, "w")
        kh, kw, ci, co = w.shape
        arr[i] = _lib.PackJob(w.data_ptr(), kh * kw, ci, co, int(contract_ci))
    return arr


def pack_weights_batch(jobs, bf16: bool):
    """(Re)pack the tensor-core weight copies of many layers in one launch (see dofb_pack_weights_batch)."""
    check(_lib.load().dofb_pack_weights_batch(C.cast(jobs, C.c_void_p), len(jobs), int(bf16), _stream()))


def invalidate_weight_cache():
    _lib.load().dofb_invalidate_weight_cache()


def _need32(s: Slab, who: str):
    if s.ptr is None:
        raise DeepOFError(f"{who}: this kernel reads/writes the fp32 buffer, but the slab is bf16-only")
    return s.ptr


def head_wz_pack(ws, wzs):
    """wz[k] = [C,20] tap-in-N form of the head weights w[k] = [3,3,C,2] (dofb_head_wz_pack), all heads in one launch."""
    n = len(ws)
    a = (C.c_void_p * n)(*[t.data_ptr() for t in ws])
    b = (C.c_void_p * n)(*[t.data_ptr() for t in wzs])
    cs = (C.c_int * n)(*[int(t.shape[2]) for t in ws])
    check(_lib.load().dofb_head_wz_pack(n, a, b, cs, _stream()))


def head_wgrad_tc(x: Slab, d9, dw):
    """dw[3,3,C,2] += weight gradient of the flow head from the bf16 feature map ``x`` and the bf16 im2col ``d9`` of dpr (dofb_head_wgrad_bf16)."""
    _req(dw, "dw")
    assert d9.dtype == torch.bfloat16 and d9.is_cuda and d9.is_contiguous() and tuple(d9.shape[:3]) == (x.B, x.h, x.w)
    check(_lib.load().dofb_head_wgrad_bf16(_need16(x, "head_wgrad_tc"), x.ld, d9.data_ptr(), d9.shape[3], x.B, x.h, x.w, x.c, dw.data_ptr(), _stream()))


def head_tapsum(z, bias, pr):
    _req(z, "z"); _req(pr, "pr")
    B, h, w, ld = z.shape
    check(_lib.load().dofb_head_tapsum(z.data_ptr(), ld, B, h, w, bias.data_ptr(), pr.data_ptr(), _stream()))


def head_dpr9(dpr, d9, dbias=None):
    _req(dpr, "dpr")
    assert d9.dtype == torch.bfloat16 and d9.is_cuda and d9.is_contiguous() and tuple(d9.shape[:3]) == tuple(dpr.shape[:3])
    B, h, w, _ = dpr.shape
    check(_lib.load().dofb_head_dpr9(dpr.data_ptr(), B, h, w, d9.data_ptr(), d9.shape[3], dbias.data_ptr() if dbias is not None else None, _stream()))


def head_dgrad_elu(d9, wz, c0: int, g: Slab | None, y: Slab, out: Slab, c_elu: int, db=None, lin_out: torch.Tensor | None = None):
    """Finish the gradient of the channel slab ``out`` (= channels [c0, c0+out.c) of the head's input feat_s): adds the head's input
    gradient (D9 row . Wz rows) to ``g`` (or to zero), applies ELU' (from the bf16 ELU outputs ``y``) on the first c_elu channels -> bf16
    shadow of ``out`` and bias gradient ``db``; the remaining (linear) channels are written in fp32 to ``out`` -- or to the compact tensor
    ``lin_out`` -- (dofb_head_dgrad_elu_bf16)."""
    _req(wz, "wz")
    assert d9.dtype == torch.bfloat16 and d9.is_cuda and d9.is_contiguous() and d9.dim() == 4
    B, h, w, d9_ld = d9.shape
    c = out.c
    c_total = wz.numel() // 20
    assert (g is None or g.c == c) and y.c == c and out.n_pix == B * h * w
    lin = c_elu < c
    gout, gout_ld = None, out.ld
    if lin and lin_out is not None:         # compact [B,h,w,c-c_elu] destination of the linear channels (slab channel ch lands at ch - c_elu)
        _req(lin_out, "lin_out")
        assert tuple(lin_out.shape) == (B, h, w, c - c_elu)
        gout, gout_ld = lin_out.data_ptr() - 4 * int(c_elu), c - c_elu
    elif lin:
        gout = _need32(out, "head_dgrad_elu")
    check(_lib.load().dofb_head_dgrad_elu_bf16(
        d9.data_ptr(), d9_ld, B, h, w, wz.data_ptr(), c_total, int(c0), c, int(c_elu),
        _need32(g, "head_dgrad_elu") if g is not None else None, g.ld if g is not None else 0,
        _need16(y, "head_dgrad_elu") if c_elu else None, y.ld, _need16(out, "head_dgrad_elu") if c_elu else None, out.ld,
        gout, gout_ld, db.data_ptr() if db is not None else None, _stream()))


def head_fwd(x: Slab, w, b, pr):
    _req(pr, "pr")
    check(_lib.load().dofb_head_fwd(x.ptr, x.ld, x.B, x.h, x.w, x.c, w.data_ptr(), b.data_ptr(), pr.data_ptr(), _stream()))


def head_dgrad(dpr, w, dx: Slab, accumulate: bool):
    _req(dpr, "dpr")
    check(_lib.load().dofb_head_dgrad(dpr.data_ptr(), dx.B, dx.h, dx.w, dx.c, w.data_ptr(), dx.ptr, dx.ld, int(accumulate), _stream()))


def head_wgrad(x: Slab, dpr, dw, db):
    check(_lib.load().dofb_head_wgrad(x.ptr, x.ld, dpr.data_ptr(), x.B, x.h, x.w, x.c, dw.data_ptr(),
                                      db.data_ptr() if db is not None else None, _stream()))


def uppr_fwd(pr, w, b, y: Slab):
    B, h, wd, _ = pr.shape
    assert y.c == 2 and y.h == 2 * h and y.w == 2 * wd
    check(_lib.load().dofb_uppr_fwd(pr.data_ptr(), B, h, wd, w.data_ptr(), b.data_ptr(), y.ptr, y.ptr16, y.ld, _stream()))   # (+ bf16 shadow)


def uppr_bwd(pr, dy: Slab, w, dpr, dw, db):
    B, h, wd, _ = pr.shape
    check(_lib.load().dofb_uppr_bwd(pr.data_ptr(), dy.ptr, dy.ld, B, h, wd, w.data_ptr(), dpr.data_ptr(), dw.data_ptr(),
                                    db.data_ptr() if db is not None else None, _stream()))


def adam(theta, g, m, v, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    for t in (theta, g, m, v):
        _req(t, "adam arena")
    check(_lib.load().dofb_adam(theta.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), theta.numel(),
                                lr_t, beta1, beta2, eps, grad_scale, _stream()))


def epe_sum(flow, gt, out):
    _req(flow, "flow"); _req(gt, "gt")
    assert out.dtype == torch.float64 and out.is_cuda
    check(_lib.load().dofb_epe_sum(flow.data_ptr(), gt.data_ptr(), flow.numel() // 2, out.data_ptr(), _stream()))


def corr_fwd(f1: Slab, f2: Slab, out: Slab, max_disp=20, stride2=2, act=ACT_NONE, math=MATH_FP32):
    assert f1.ld == f2.ld and f1.c == f2.c
    if math == MATH_BF16:       # bf16 shadows of the maps; fp32 output + (when the slab has one) its bf16 shadow in the same pass
        check(_lib.load().dofb_corr_fwd_bf16(_need16(f1, "corr_fwd"), _need16(f2, "corr_fwd"), f1.ld, f1.B, f1.h, f1.w, f1.c, max_disp, stride2,
                                             out.ptr, out.ptr16, out.ld, act, _stream()))
        return
    check(_lib.load().dofb_corr_fwd(f1.ptr, f2.ptr, f1.ld, f1.B, f1.h, f1.w, f1.c, max_disp, stride2, out.ptr, out.ld, act, math,
                                    _stream()))


def corr_bwd(f1: Slab, f2: Slab, dout: Slab, df1: Slab, df2: Slab, max_disp=20, stride2=2, math=MATH_FP32):
    assert df1.ld == df2.ld
    if math == MATH_BF16:
        check(_lib.load().dofb_corr_bwd_bf16(_need16(f1, "corr_bwd"), _need16(f2, "corr_bwd"), f1.ld, f1.B, f1.h, f1.w, f1.c, max_disp, stride2,
                                             dout.ptr, dout.ld, df1.ptr, df2.ptr, df1.ld, _stream()))
        return
    check(_lib.load().dofb_corr_bwd(f1.ptr, f2.ptr, f1.ld, f1.B, f1.h, f1.w, f1.c, max_disp, stride2, dout.ptr, dout.ld,
                                    df1.ptr, df2.ptr, df1.ld, math, _stream()))


class WarpLoss:
    """All pyramid scales of loss_interp (forward + d/dflow) in one launch."""

    def __init__(self, device):
        self.device = device
        self.ws = None

    def __call__(self, scales: list[dict]):
        n = len(scales)
        arr = (LossScale * n)()
        keep = []
        for i, s in enumerate(scales):
            flow, src, tgt = s["flow"], s["src"], s["tgt"]
            for t, nm in ((flow, "flow"), (src, "src"), (tgt, "tgt"), (s["loss4"], "loss4")):
                _req(t, nm)
            B, h, w, _ = flow.shape
            if tuple(src.shape) != (B, h, w, 3) or tuple(tgt.shape) != (B, h, w, 3):
                raise DeepOFError(f"loss_interp: inputs/outputs must be [B,h,w,3] matching the flow {tuple(flow.shape)}")
            recon, dflow = s.get("recon"), s.get("dflow")
            ew = s.get("edge_w")
            if ew is not None:
                _req(ew, "edge_w")
                if tuple(ew.shape) != (B, h, w, 2) or int(s.get("variant", 0)) != 1:
                    raise DeepOFError("loss_interp: edge weights must be [B,h,w,2] and need the variant-B loss")
            arr[i] = LossScale(flow.data_ptr(), src.data_ptr(), tgt.data_ptr(),
                               recon.data_ptr() if recon is not None else None,
                               dflow.data_ptr() if dflow is not None else None,
                               s["loss4"].data_ptr(), B, h, w, float(s["flow_scale"]),
                               float(s["epsilon"]), float(s["alpha_c"]), float(s["alpha_s"]), float(s["lambda_smooth"]),
                               float(s.get("g_charb", 1.0)), float(s.get("g_u", 1.0)), float(s.get("g_v", 1.0)),
                               int(s.get("variant", 0)), ew.data_ptr() if ew is not None else None)
            keep.append((flow, src, tgt, recon, dflow, ew))
        lib = _lib.load()
        need = lib.dofb_warp_loss_workspace_bytes(n, arr)
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.zeros(int(need) + 256, dtype=torch.uint8, device=self.device)
        base = self.ws.data_ptr()
        off = (-base) % 256
        check(lib.dofb_warp_loss(n, arr, base + off, self.ws.numel() - off, _stream()))


def _workspace(cache: dict, key, nbytes: int, device) -> torch.Tensor:
    ws = cache.get(key)
    if ws is None or ws.numel() < nbytes + 256:
        ws = cache[key] = torch.zeros(int(nbytes) + 256, dtype=torch.uint8, device=device)
    return ws


_ws_cache: dict = {}


def edge_weights(img: torch.Tensor) -> torch.Tensor:
    """[B,h,w,2] edge weights of the edge-aware smoothness (version1/model/warpflow.py:91-116) for ``img`` [B,h,w,3] (dofb_edge_weights)."""
    _req(img, "img")
    B, h, w, c = img.shape
    if c != 3:
        raise DeepOFError("edge_weights: expected [B,h,w,3]")
    lib = _lib.load()
    need = lib.dofb_edge_weights_workspace_bytes(B, h, w)
    ws = _workspace(_ws_cache, ("edge", str(img.device)), need, img.device)
    base = ws.data_ptr()
    off = (-base) % 256
    out = torch.empty(B, h, w, 2, dtype=torch.float32, device=img.device)
    check(lib.dofb_edge_weights(img.data_ptr(), B, h, w, out.data_ptr(), base + off, ws.numel() - off, _stream()))
    return out


def flow_stencil(delta_weights) -> "_lib.FlowStencil":
    """Non-zero entries of a dense [3,3,Cf,Cf] smoothness constant (deltaWeights["FlowDeltaWeights"], sintelWrapFlow.py:378) as the
    sparse stencil dofb_warp_loss_multi takes: out[p, cout] += w * in[p + (kh-1, kw-1), cin]."""
    dw = torch.as_tensor(delta_weights, dtype=torch.float32).cpu()
    if dw.dim() != 4 or tuple(dw.shape[:2]) != (3, 3) or dw.shape[2] != dw.shape[3]:
        raise DeepOFError(f"FlowDeltaWeights must be [3,3,Cf,Cf], got {tuple(dw.shape)}")
    nz = dw.nonzero().tolist()
    if len(nz) > 64:
        raise DeepOFError(f"FlowDeltaWeights has {len(nz)} non-zero entries; the CUDA stencil holds at most 64")
    st = _lib.FlowStencil()
    st.n = len(nz)
    for i, (kh, kw, cin, cout) in enumerate(nz):
        st.e[i].dy, st.e[i].dx, st.e[i].cin, st.e[i].cout, st.e[i].w = kh - 1, kw - 1, cin, cout, float(dw[kh, kw, cin, cout])
    return st


def warp_loss_multi(flow, frames, stencil, flow_scale, epsilon, alpha_c, alpha_s, lambda_smooth, g=(1.0, 1.0, 1.0), want_recon=True,
                    want_grad=True):
    """sintelWrapFlow.loss_interp_multi forward + d/dflow (dofb_warp_loss_multi) -> (loss4 [4], recon or None, dflow or None)."""
    _req(flow, "flow"); _req(frames, "frames")
    B, h, w, cf = flow.shape
    if frames.shape[:3] != flow.shape[:3] or frames.shape[3] % 3 or cf != 2 * (frames.shape[3] // 3 - 1) or cf < 2:
        raise DeepOFError(f"loss_interp_multi: flows {tuple(flow.shape)} do not match inputs {tuple(frames.shape)} (3T image and 2(T-1) flow channels)")
    P = cf // 2
    lib = _lib.load()
    need = lib.dofb_warp_loss_multi_workspace_bytes(B, h, w)
    ws = _workspace(_ws_cache, ("multi", str(flow.device)), need, flow.device)
    base = ws.data_ptr()
    off = (-base) % 256
    loss4 = torch.empty(4, dtype=torch.float32, device=flow.device)
    recon = torch.empty(B, h, w, 3 * P, dtype=torch.float32, device=flow.device) if want_recon else None
    dflow = torch.empty_like(flow) if want_grad else None
    check(lib.dofb_warp_loss_multi(flow.data_ptr(), frames.data_ptr(), recon.data_ptr() if recon is not None else None,
                                   dflow.data_ptr() if dflow is not None else None, loss4.data_ptr(), B, h, w, P, float(flow_scale),
                                   float(epsilon), float(alpha_c), float(alpha_s), float(lambda_smooth), float(g[0]), float(g[1]), float(g[2]),
                                   C.byref(stencil), base + off, ws.numel() - off, _stream()))
    return loss4, recon, dflow


def decode_ppm(raw: torch.Tensor, data_off: torch.Tensor, src_hw, out_hw) -> torch.Tensor:
    """raw: uint8 CUDA tensor with the file bytes; data_off: int64 CUDA tensor [B] (first pixel byte of each image) -> [B,oh,ow,3] BGR float."""
    assert raw.dtype == torch.uint8 and raw.is_cuda and data_off.dtype == torch.int64 and data_off.is_cuda
    B = data_off.numel()
    out = torch.empty(B, out_hw[0], out_hw[1], 3, dtype=torch.float32, device=raw.device)
    check(_lib.load().dofb_decode_ppm(raw.data_ptr(), data_off.data_ptr(), B, int(src_hw[0]), int(src_hw[1]), out.data_ptr(), int(out_hw[0]),
                                      int(out_hw[1]), _stream()))
    return out


def decode_flo(raw: torch.Tensor, file_off: torch.Tensor, hw):
    """-> ([B,h,w,2] float flow, status int tensor: non-zero = a header was not a .flo of that size)."""
    assert raw.dtype == torch.uint8 and raw.is_cuda and file_off.dtype == torch.int64 and file_off.is_cuda
    B = file_off.numel()
    out = torch.empty(B, hw[0], hw[1], 2, dtype=torch.float32, device=raw.device)
    status = torch.zeros(1, dtype=torch.int32, device=raw.device)
    check(_lib.load().dofb_decode_flo(raw.data_ptr(), file_off.data_ptr(), B, int(hw[0]), int(hw[1]), out.data_ptr(), status.data_ptr(), _stream()))
    return out, status


def eval_flow_aee(flow1: torch.Tensor, gt: torch.Tensor, mult=2.0, clip=(-300.0, 250.0)) -> torch.Tensor:
    """Average end-point error of the reference's evaluation recipe (flyingChairsTrain.py:264-266,294-296): a float64 device scalar."""
    _req(flow1, "flow"); _req(gt, "gt")
    B, h, w, _ = flow1.shape
    Bg, H, W, _ = gt.shape
    if Bg != B:
        raise DeepOFError("eval_flow_aee: batch mismatch")
    out = torch.zeros(1, dtype=torch.float64, device=flow1.device)
    check(_lib.load().dofb_eval_flow_aee_sum(flow1.data_ptr(), B, h, w, gt.data_ptr(), H, W, float(mult), float(clip[0]), float(clip[1]),
                                             out.data_ptr(), _stream()))
    return out / float(B * H * W)
