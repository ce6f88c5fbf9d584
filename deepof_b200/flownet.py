"""FlowNetS guided-flow training engine: static memory plan + hand-scheduled forward/backward.

Replaces the TF graph built by ``flyingChairsWrapFlow.flowNet`` (flyingChairsWrapFlow.py:5-129)
plus ``AdamOptimizer.minimize`` (flyingChairsTrain.py:121-124).  There is no tracing compiler and
no autograd on this path: the layer list is static, so the backward schedule is written out once
(reverse topological order, accumulation flags fixed at plan time) and every launch is one of the
CUDA kernels behind include/deepof_b200.h.

Memory plan (NHWC fp32, sized for 180 GB HBM3e; B=32 @384x512 needs ~3 GB):
  * one flat parameter arena + identically laid out gradient / Adam-m / Adam-v arenas
    (TF layouts, every tensor 256-byte aligned) -> Adam and the gradient all-reduce are single
    passes over flat memory;
  * tf.concat is never materialised: conv / deconv / up_pr write into channel slices of the
    concat buffers (pitch rounded up to 32 floats: 128, 224, 416, 800, 1056);
  * gradient buffers mirror the activation buffers; a slice's gradient is accumulated in place by
    its consumers in a fixed order, then multiplied by ELU' once.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch

from . import ops
from .ops import Slab, full, conv_geom, ACT_ELU, ACT_NONE, MATH_FP32, MATH_TF32, MATH_BF16

FLYINGCHAIRS_MEAN = (97.533268117955444, 99.238235788550085, 97.055973199626948)   # flyingChairsWrapFlow.py:16
SINTEL_MEAN = (70.1433, 83.1915, 92.8827)                                          # sintelWrapFlow.py:773

# (name, k, stride, cin, cout)  flyingChairsWrapFlow.py:31-40
TOWER = [("conv1", 7, 2, 6, 64), ("conv2", 5, 2, 64, 128), ("conv3_1", 5, 2, 128, 256), ("conv3_2", 3, 1, 256, 256),
         ("conv4_1", 3, 2, 256, 512), ("conv4_2", 3, 1, 512, 512), ("conv5_1", 3, 2, 512, 512), ("conv5_2", 3, 1, 512, 512),
         ("conv6_1", 3, 2, 512, 1024), ("conv6_2", 3, 1, 1024, 1024)]
# (s, feat channels at s, upconv name, upconv cout, up_pr name, skip channels at s-1)  :58-111
REFINE = [(6, 1024, "upconv5", 512, "up_pr6to5", 512), (5, 1026, "upconv4", 256, "up_pr5to4", 512),
          (4, 770, "upconv3", 128, "up_pr4to3", 256), (3, 386, "upconv2", 64, "up_pr3to2", 128),
          (2, 194, "upconv1", 32, "up_pr2to1", 64)]
FEAT_C = {6: 1024, 5: 1026, 4: 770, 3: 386, 2: 194, 1: 98}
FLOW_SCALES = {1: 10.0, 2: 5.0, 3: 2.5, 4: 1.25, 5: 0.625, 6: 0.3125}            # :118,107,96,85,74,63
HYPER = dict(epsilon=1e-4, alpha_c=0.25, alpha_s=0.37, lambda_smooth=1.0)         # :43-46
LOSS_WEIGHTS = (16.0, 8.0, 4.0, 2.0, 1.0, 1.0)                                    # flyingChairsTrain.py:165


def param_shapes() -> "OrderedDict[str, tuple]":
    """The 52 trainable tensors in TF creation order and TF layouts."""
    sh: "OrderedDict[str, tuple]" = OrderedDict()
    for name, k, _s, cin, cout in TOWER:
        sh[name + "/weights"] = (k, k, cin, cout)
        sh[name + "/biases"] = (cout,)
    for s, cfeat, up, upc, uppr, _skip in REFINE:
        sh[f"pr{s}/weights"] = (3, 3, cfeat, 2)
        sh[f"pr{s}/biases"] = (2,)
        sh[up + "/weights"] = (4, 4, upc, cfeat)
        sh[up + "/biases"] = (upc,)
        sh[uppr + "/weights"] = (4, 4, 2, 2)
        sh[uppr + "/biases"] = (2,)
    sh["pr1/weights"] = (3, 3, 98, 2)
    sh["pr1/biases"] = (2,)
    return sh


def _round_up(x, m):
    return (x + m - 1) // m * m


class ParamArena:
    """Flat fp32 arena with named TF-layout views (256-byte aligned starts)."""

    def __init__(self, shapes: "OrderedDict[str, tuple]", device):
        self.shapes = shapes
        self.offsets = OrderedDict()
        off = 0
        for name, shape in shapes.items():
            self.offsets[name] = off
            off += _round_up(math.prod(shape), 64)
        self.numel = off
        self.n_true = sum(math.prod(s) for s in shapes.values())
        self.device = device

    def new(self):
        return torch.zeros(self.numel, dtype=torch.float32, device=self.device)

    def views(self, flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for name, shape in self.shapes.items():
            o = self.offsets[name]
            out[name] = flat[o:o + math.prod(shape)].view(shape)
        return out


def xavier_uniform(shape, gen):
    rf = math.prod(shape[:-2])
    lim = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).float()


def bilinear_deconv(shape):
    """flyingChairsTrain.py:78-92 (k=4 -> outer([.25,.75,.75,.25]) on the channel diagonal)."""
    k = shape[0]
    f = float(math.ceil(k / 2.0))
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    ax = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float64)
    w = torch.zeros(shape, dtype=torch.float64)
    for i in range(shape[2]):
        w[:, :, i, i] = torch.outer(ax, ax)
    return w.float()


class Buf:
    """One NHWC activation / gradient buffer: fp32 master (``t``, may be None in the lean bf16 engine) + bf16 shadow (``t16``, may be None)."""
    __slots__ = ("t", "t16", "shape")

    def __init__(self, shape, device, want32=True, want16=False):
        self.shape = tuple(shape)
        self.t = torch.zeros(shape, dtype=torch.float32, device=device) if want32 else None
        self.t16 = torch.zeros(shape, dtype=torch.bfloat16, device=device) if want16 else None


class FlowNetS:
    """Static-shape FlowNetS engine.  ``math``: 'fp32' (SIMT FFMA, parity grade), 'tf32' or 'bf16' (tcgen05).

    bf16 math runs the LEAN schedule unless DOFB_LEAN=0: activations that only tensor-core kernels read are kept in bf16
    alone, the flow heads run on the tensor pipe in tap-in-N form (csrc/heads_tc.cu) and the head input gradients are added on the fly
    by the ELU' pass of each slab instead of being written to memory."""

    def __init__(self, batch: int, height: int = 384, width: int = 512, device="cuda", variant: str = "A",
                 math_mode: str = "fp32", mean=FLYINGCHAIRS_MEAN, hyper=None, seed: int | None = 1, tc_wgrad: bool = False):
        mult = 32 if self.ARCH == "V" else 64
        if height % mult or width % mult:
            raise ValueError(f"H and W must be multiples of {mult} (the reference resizes/pads as well, SURVEY.md 0.5)")
        if not torch.cuda.is_available():
            raise ops.DeepOFError("deepof_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.B, self.H, self.W = batch, height, width
        self.device = torch.device(device)
        self.variant = {"A": 0, "B": 1}[variant]
        self.math = {"fp32": MATH_FP32, "tf32": MATH_TF32, "bf16": MATH_BF16}[math_mode]
        # per-op math: the tcgen05 path needs 32-float pitches (conv1 reads the 8-float pitched input) and has its own
        # weight-gradient kernel; anything it does not cover runs on the fp32 SIMT kernels (explicit, per layer).
        self.math_wgrad = self.math if (tc_wgrad or self.math == MATH_BF16) else MATH_FP32
        self.mean = tuple(float(m) for m in mean)
        self.hyper = dict(HYPER)
        if hyper:
            self.hyper.update(hyper)
        self.lean = self.math == MATH_BF16 and self.ARCH in ("S", "C") and os.environ.get("DOFB_LEAN", "1") != "0"
        self.arena = ParamArena(self.param_shapes(), self.device)
        self.theta = self.arena.new()
        self.grad = self.arena.new()
        self.m, self.v = self.arena.new(), self.arena.new()
        self.params = self.arena.views(self.theta)
        self.grads = self.arena.views(self.grad)
        self.t = 0
        self._alloc()
        self._plan()
        self.warp_loss = ops.WarpLoss(self.device)
        ops._lib.load().dofb_enable_weight_cache(1)     # this engine invalidates after every parameter change
        ops._lib.load().dofb_enable_cta_pairs(0 if os.environ.get("DOFB_CTA_PAIRS", "1") == "0" else 1)   # cta_group::2 tiles for the wide layers
        # halo-tile reuse of A across taps: correct (tests) but slower than the per-tap gather in its first form -> opt-in
        ops._lib.load().dofb_enable_halo_tiles(1 if os.environ.get("DOFB_HALO", "0") == "1" else 0)
        ops._lib.load().dofb_enable_phase_in_n(0 if os.environ.get("DOFB_PIN", "1") == "0" else 1)      # phase-in-N stride-2 transposed gathers
        ops._lib.load().dofb_enable_split_k(int(os.environ.get("DOFB_SPLITK", "1")))                    # split-K of the coarse layers
        # second stream for the flow-head / up_pr chains (small, latency-bound launches that only meet the big GEMMs at level boundaries)
        self._side = None
        if self.lean and os.environ.get("DOFB_SIDE_STREAM", "1") != "0":
            with torch.cuda.device(self.device):
                self._side = torch.cuda.Stream(device=self.device)
        self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        # ... and every weight gradient as well: the main stream keeps the critical chain (ELU' pass -> input gradient -> next ELU' pass), the
        # weight-gradient GEMMs queue up behind it on the side stream and share the SMs with the bandwidth-bound ELU' passes
        self._side_wgrad = self._side is not None and os.environ.get("DOFB_SIDE_WGRAD", "1") != "0"
        self._zero_ev = None
        self._side_extra = self._side_wgrad and os.environ.get("DOFB_SIDE_EXTRA", "1") != "0"   # gradient clear under the forward, pr1's weight gradient
        ops._lib.load().dofb_enable_wgrad_npack(0 if os.environ.get("DOFB_NPACK", "1") == "0" else 1)
        self.profile = None            # list of (tag, start_event, end_event) when per-launch timing is on
        self._nvtx = os.environ.get("DOFB_NVTX", "0") == "1"
        if seed is not None:
            self.init_params(seed)

    @staticmethod
    def param_shapes():
        return param_shapes()

    # ---- side stream: fork = the side stream waits for everything enqueued on the main stream so far; join = the reverse ----
    def _side_on(self) -> bool:
        return self._side is not None and self.profile is None and not self._nvtx

    def _fork(self):
        if self._side_on():
            self._ev_fork.record()
            self._side.wait_event(self._ev_fork)

    def _join(self):
        if self._side_on():
            self._ev_join.record(self._side)
            torch.cuda.current_stream(self.device).wait_event(self._ev_join)

    def _join_at(self, ev):
        """The main stream waits for an event recorded on the side stream (finer than _join: later side work keeps running)."""
        if self._side_on():
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _side_event(self):
        if not self._side_on():
            return None
        ev = torch.cuda.Event()
        ev.record(self._side)
        return ev

    def _ks(self, tag, fn, *args, **kw):
        """_k on the side stream (between _fork() and _join()); serial when profiling."""
        if not self._side_on():
            return self._k(tag, fn, *args, **kw)
        with torch.cuda.stream(self._side):
            return fn(*args, **kw)

    def _k(self, tag, fn, *args, **kw):
        """Launch one kernel; with self.profile set, bracket it with CUDA events on the launch stream."""
        if self._nvtx:                  # DOFB_NVTX=1: one NVTX range per launch tag (ncu --nvtx --nvtx-include "conv_dgrad:conv2/")
            torch.cuda.nvtx.range_push(tag)
            try:
                return fn(*args, **kw)
            finally:
                torch.cuda.nvtx.range_pop()
        if self.profile is None:
            return fn(*args, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*args, **kw)
        e1.record()
        self.profile.append((tag, e0, e1))

    # ------------------------------------------------------------------ parameters
    def init_params(self, seed: int = 1):
        """slim defaults (xavier-uniform, zero bias) + bilinear overwrite of every 'up*' weight."""
        gen = torch.Generator().manual_seed(seed)
        for name, shape in self.arena.shapes.items():
            if name.endswith("biases"):
                self.params[name].zero_()
                continue
            w = xavier_uniform(shape, gen)
            if name.startswith("up"):
                w = bilinear_deconv(shape)
            self.params[name].copy_(w)
        ops.invalidate_weight_cache()

    def load_params(self, params: dict):
        for name, p in params.items():
            self.params[name].copy_(torch.as_tensor(p, dtype=torch.float32).reshape(self.arena.shapes[name]))
        ops.invalidate_weight_cache()

    def export_params(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, v.detach().cpu().clone()) for k, v in self.params.items())

    # ------------------------------------------------------------------ buffers
    ARCH = "S"
    N_SCALES = 6
    REFINE_SPEC = REFINE

    TC_HEAD_MAX_PIX = 32 * 24 * 32          # heads with at most this many output pixels go through the tensor-core GEMM

    def _alloc(self):
        B, H, W, dev = self.B, self.H, self.W, self.device
        self._head_geom = {}
        z = lambda h, w, c, b=B: torch.zeros(b, h, w, c, dtype=torch.float32, device=dev)  # noqa: E731
        self._z = z
        # first-layer input: dense for the SIMT path; zero-bordered (2 rows/cols before, 4/6 after) for the tcgen05 first-layer
        # kernel, which reads the SAME padding (2,3) of the 7x7/2 conv straight from the border
        tensor = self.math != MATH_FP32
        self.x6_origin = (2, 2) if tensor else (0, 0)
        xshape = (H + 6, W + 8, 8) if tensor else (H, W, 8)
        if self.ARCH == "V":            # VGG16: 3x3/1 first layer, generic kernels, 6 channels padded to one K block (32 floats / 64 bf16)
            xshape, self.x6_origin = (H, W, 64 if self.math == MATH_BF16 else 32), (0, 0)
        self.x6 = z(*xshape)
        self.x6b = z(*xshape) if self.ARCH == "C" else None          # siamese: target image in its own buffer
        shp = self._buffer_shapes()
        if self.math == MATH_BF16:      # bf16 K blocks are 64 channels: pitches become multiples of 64
            shp = {k: (h, w, _round_up(c, 64)) for k, (h, w, c) in shp.items()}
        bf = self.math == MATH_BF16
        # lean bf16 engine: only buffers some fp32 kernel still reads keep their fp32 master (FlowNetC: the correlation runs on fp32 maps)
        keep32 = self.NEED32 if self.lean else set(shp)
        self.act = {k: Buf((B,) + v, dev, want32=k in keep32, want16=bf) for k, v in shp.items()}
        self.dact = {k: Buf((B,) + v, dev, want32=True, want16=bf) for k, v in shp.items()}
        # bf16 copies of the first-layer inputs (VGG: read by the generic kernels; S/C: by the bf16 first-layer kernels)
        self._sh = {}
        if bf:
            self._sh[id(self.x6)] = torch.zeros(self.x6.shape, dtype=torch.bfloat16, device=dev)
            if self.x6b is not None:
                self._sh[id(self.x6b)] = torch.zeros(self.x6b.shape, dtype=torch.bfloat16, device=dev)
        self.hw = {s: (H >> s, W >> s) for s in range(1, self.N_SCALES + 1)}
        self.pr = {s: z(*self.hw[s], 2) for s in range(1, self.N_SCALES + 1)}
        self.dpr = {s: z(*self.hw[s], 2) for s in range(1, self.N_SCALES + 1)}
        self.pyr_src = {s: z(*self.hw[s], 3) for s in range(1, self.N_SCALES + 1)}
        self.pyr_tgt = {s: z(*self.hw[s], 3) for s in range(1, self.N_SCALES + 1)}
        self.recon1 = z(*self.hw[1], 3)
        self.loss4 = torch.zeros(self.N_SCALES, 4, dtype=torch.float32, device=dev)
        if self.lean:
            # tap-in-N flow heads: Z maps share one flat buffer (forward is sequential), the bf16 im2col of dpr_s has one buffer per scale
            # (its 64-column rows keep columns 18.. at zero), [C,20] weights per head
            h1, w1 = self.hw[1]
            self._z_flat = torch.zeros(B * h1 * w1 * 20, dtype=torch.float32, device=dev)
            self.head_z = {s: self._z_flat[:B * hh * ww * 20].view(B, hh, ww, 20) for s, (hh, ww) in self.hw.items()}
            self.head_d9 = {s: torch.zeros(B, hh, ww, 64, dtype=torch.bfloat16, device=dev) for s, (hh, ww) in self.hw.items()}
            # gradient of the 2-channel up_pr outputs: compact [B,h,w,2] maps (inside the pitched concat gradient every pixel of them is its own
            # 32-byte sector for the 4x4 gather of uppr_bwd)
            self.dpr_up = {s: torch.zeros(B, hh, ww, 2, dtype=torch.float32, device=dev) for s, (hh, ww) in self.hw.items() if s < self.N_SCALES}
            self.head_wz = {s: torch.zeros(1, 1, self.arena.shapes[f"pr{s}/weights"][2], 20, dtype=torch.float32, device=dev)
                            for s in range(1, self.N_SCALES + 1)}

    NEED32 = frozenset()            # activation buffers that keep an fp32 master in the lean engine

    def _S(self, t, c0, c):
        if isinstance(t, Buf):
            return Slab(t.t, c0, c, t.t16)
        return Slab(t, c0, c, self._sh.get(id(t)))

    def _F(self, t):
        return self._S(t, 0, t.shape[3])

    def _buffer_shapes(self):
        H, W = self.H, self.W
        shp = {"concat1": (H // 2, W // 2, 128), "concat2": (H // 4, W // 4, 224), "c31": (H // 8, W // 8, 256),
               "concat3": (H // 8, W // 8, 416), "c41": (H // 16, W // 16, 512), "concat4": (H // 16, W // 16, 800),
               "c51": (H // 32, W // 32, 512), "concat5": (H // 32, W // 32, 1056), "c61": (H // 64, W // 64, 1024),
               "c62": (H // 64, W // 64, 1024)}
        if self.ARCH == "C":
            del shp["c31"]
            shp.update({"c1b": (H // 2, W // 2, 64), "c2b": (H // 4, W // 4, 128), "c3a": (H // 8, W // 8, 256),
                        "c3b": (H // 8, W // 8, 256), "cat3": (H // 8, W // 8, 480)})
        return shp

    def _conv_rec(self, name, k, stride, cin, cout, x, dx, y, dy, acc=False, wname=None, xpad=None, ih=None, iw=None):
        """One conv layer record: y = ELU(conv(x)); backward writes (or accumulates, acc=True) into dx."""
        ih = ih if ih is not None else (x.h if x is not None else self.H)
        iw = iw if iw is not None else (x.w if x is not None else self.W)
        return dict(op="conv", name=name, wname=wname or name, g=conv_geom(self.B, ih, iw, cin, cout, k, stride),
                    x=x, dx=dx, y=y, dy=dy, acc=acc, xpad=xpad)

    def _plan_tower(self):
        a, d, S, full = self.act, self.dact, self._S, self._F
        first_x = S(self.x6, 0, 6) if self.math == MATH_FP32 else None
        first_pad = self.x6 if self.math != MATH_FP32 else None
        R = self._conv_rec
        return [
            R("conv1", 7, 2, 6, 64, first_x, None, S(a["concat1"], 0, 64), S(d["concat1"], 0, 64), xpad=first_pad),
            R("conv2", 5, 2, 64, 128, S(a["concat1"], 0, 64), S(d["concat1"], 0, 64), S(a["concat2"], 0, 128), S(d["concat2"], 0, 128), acc=True),
            R("conv3_1", 5, 2, 128, 256, S(a["concat2"], 0, 128), S(d["concat2"], 0, 128), full(a["c31"]), full(d["c31"]), acc=True),
            R("conv3_2", 3, 1, 256, 256, full(a["c31"]), full(d["c31"]), S(a["concat3"], 0, 256), S(d["concat3"], 0, 256)),
        ] + self._plan_tower_top()

    def _plan_tower_top(self):
        a, d, S, R, full = self.act, self.dact, self._S, self._conv_rec, self._F
        return [
            R("conv4_1", 3, 2, 256, 512, S(a["concat3"], 0, 256), S(d["concat3"], 0, 256), full(a["c41"]), full(d["c41"]), acc=True),
            R("conv4_2", 3, 1, 512, 512, full(a["c41"]), full(d["c41"]), S(a["concat4"], 0, 512), S(d["concat4"], 0, 512)),
            R("conv5_1", 3, 2, 512, 512, S(a["concat4"], 0, 512), S(d["concat4"], 0, 512), full(a["c51"]), full(d["c51"]), acc=True),
            R("conv5_2", 3, 1, 512, 512, full(a["c51"]), full(d["c51"]), S(a["concat5"], 0, 512), S(d["concat5"], 0, 512)),
            R("conv6_1", 3, 2, 512, 1024, S(a["concat5"], 0, 512), S(d["concat5"], 0, 512), full(a["c61"]), full(d["c61"]), acc=True),
            R("conv6_2", 3, 1, 1024, 1024, full(a["c61"]), full(d["c61"]), full(a["c62"]), full(d["c62"])),
        ]

    def _plan_feat(self):
        a, d, S, full = self.act, self.dact, self._S, self._F
        return {6: (full(a["c62"]), full(d["c62"])), 5: (S(a["concat5"], 0, 1026), S(d["concat5"], 0, 1026)),
                4: (S(a["concat4"], 0, 770), S(d["concat4"], 0, 770)), 3: (S(a["concat3"], 0, 386), S(d["concat3"], 0, 386)),
                2: (S(a["concat2"], 0, 194), S(d["concat2"], 0, 194)), 1: (S(a["concat1"], 0, 98), S(d["concat1"], 0, 98))}

    def _plan(self):
        B = self.B
        a, d = self.act, self.dact
        S = self._S
        self.tower = self._plan_tower()
        self.feat = self._plan_feat()
        cat = {5: "concat5", 4: "concat4", 3: "concat3", 2: "concat2", 1: "concat1"}
        self.refine = []
        for s, cfeat, up, upc, uppr, skipc in self.REFINE_SPEC:
            hs, ws = self.hw[s]
            g = conv_geom(B, 2 * hs, 2 * ws, upc, cfeat, 4, 2)     # the conv whose input-gradient is this deconv
            assert g.oh == hs and g.ow == ws and g.pad_t == 1
            tgt = cat[s - 1]
            self.refine.append(dict(s=s, g=g, up=up, uppr=uppr, skipc=skipc, upc=upc,
                                    up_y=S(a[tgt], skipc, upc), up_dy=S(d[tgt], skipc, upc),
                                    pr_y=S(a[tgt], skipc + upc, 2), pr_dy=S(d[tgt], skipc + upc, 2)))
        if self.lean:
            # which conv outputs are the skip part (channel 0..) of a flow head's input feat_s: their ELU' pass adds that head's input gradient
            for L in self.tower:
                if L["op"] != "conv":
                    continue
                for s, (fx, _fd) in self.feat.items():
                    if fx.t16 is L["y"].t16 and L["y"].c0 == 0:
                        L["head_s"] = s
                # the head of scale 1 no longer writes d concat1 first: the conv reading concat1 becomes the first writer of its gradient
                if L["dx"] is not None and L["dx"].t is self.dact["concat1"].t:
                    L["acc"] = False
            self._head_geom1 = {s: conv_geom(B, self.hw[s][0], self.hw[s][1], self.feat[s][0].c, 20, 1, 1) for s in self.feat}

    def _pack_jobs(self):
        """(weight, orientation) of every tensor-core gather-GEMM the step runs: conv fwd / transposed-conv dgrad read the contract-ci copy,
        conv dgrad / transposed-conv fwd the contract-co copy; built once (the parameter views never move)."""
        if getattr(self, "_jobs", None) is None:
            P, seen, entries = self.params, set(), []

            def add(w, contract_ci):
                key = (w.data_ptr(), contract_ci)
                if key not in seen:
                    seen.add(key)
                    entries.append((w, contract_ci))
            for L in self.tower:
                if L["op"] != "conv" or L["xpad"] is not None:
                    continue
                w = P[L["wname"] + "/weights"]
                add(w, 1)
                if L["dx"] is not None:
                    add(w, 0)
            for R in self.refine:
                w = P[R["up"] + "/weights"]
                add(w, 1)               # (contract-ci first: the batch kernel writes an adjacent (1, 0) pair in one pass over the weights)
                add(w, 0)
            for s in range(1, self.N_SCALES + 1):
                h, wd = self.hw[s]
                if self.lean:
                    add(self.head_wz[s], 1)             # tap-in-N form: a 1x1 convolution with the [1,1,C,20] weights
                elif self.B * h * wd <= self.TC_HEAD_MAX_PIX:
                    add(P[f"pr{s}/weights"], 1)
            self._jobs = ops.make_pack_jobs(entries)
        return self._jobs

    def _head_fwd(self, s, x, side=False):
        """pr_s = 3x3 conv to 2 channels.  Coarse scales (small maps, 386..1026 input channels) are GEMM-shaped with a long K and too few
        pixels to fill the GPU with the streaming SIMT kernel, so in the tensor-core math modes they run through the gather-GEMM
        (N padded to 32); the fine scales are bandwidth-bound and stay on the strip kernel."""
        P = self.params
        h, w = self.hw[s]
        if self.lean:       # tap-in-N: Z = x (1x1) Wz on the tensor pipe (x crosses the chip once), pr = bias + 9-tap sum over the 20-float map
            z = self.head_z[s]
            k = self._ks if side else self._k
            k(f"head_fwd:pr{s}", ops.conv_fwd, self._head_geom1[s], x, self.head_wz[s], None, full(z), ACT_NONE, MATH_BF16)
            k(f"head_tapsum:pr{s}", ops.head_tapsum, z, P[f"pr{s}/biases"], self.pr[s])
        elif self.math != MATH_FP32 and self.B * h * w <= self.TC_HEAD_MAX_PIX:
            g = self._head_geom.get(s)
            if g is None:
                g = self._head_geom[s] = conv_geom(self.B, h, w, x.c, 2, 3, 1)
            self._k(f"head_fwd:pr{s}", ops.conv_fwd, g, x, P[f"pr{s}/weights"], P[f"pr{s}/biases"], full(self.pr[s]), ACT_NONE, self.math)
        else:
            self._k(f"head_fwd:pr{s}", ops.head_fwd, x, P[f"pr{s}/weights"], P[f"pr{s}/biases"], self.pr[s])

    # ------------------------------------------------------------------ forward
    def _preprocess(self, source, target):
        bf = self.math == MATH_BF16 and self.x6.shape[3] == 8      # bf16 first-layer kernels: the network input is produced in bf16 only
        self._k("preprocess", ops.preprocess, source, target, self.mean, self.x6, [self.pyr_src[s] for s in range(1, self.N_SCALES + 1)],
                [self.pyr_tgt[s] for s in range(1, self.N_SCALES + 1)], self.x6_origin, self.x6b, 255.0,
                self._sh[id(self.x6)] if bf else None, self._sh[id(self.x6b)] if bf and self.x6b is not None else None)

    def _fwd_layer(self, L):
        P, mth = self.params, self.math
        if L["op"] == "conv":
            w, b = P[L["wname"] + "/weights"], P[L["wname"] + "/biases"]
            if L["xpad"] is not None:       # first layer on tensor cores from the zero-bordered buffer
                self._k("conv_fwd:" + L["name"], ops.conv1_fwd, L["g"], L["xpad"], self.x6_origin, w, b, L["y"], ACT_ELU,
                        self._sh.get(id(L["xpad"])))
            else:
                self._k("conv_fwd:" + L["name"], ops.conv_fwd, L["g"], L["x"], w, b, L["y"], ACT_ELU, mth)
        elif L["op"] == "pool":
            self._k("pool_fwd:" + L["name"], ops.maxpool2_fwd, L["x"], L["y"])
            if mth == MATH_BF16:
                self._k("cast:" + L["name"], ops.cast_bf16, L["y"])
        elif L["op"] == "corr":
            if mth == MATH_BF16 and L["f1"].t16 is not None and L["y"].t16 is not None and os.environ.get("DOFB_CORR_FWD16", "0") == "1":
                # opt-in: bf16 shadows of conv3a / conv3b on the tensor pipe with the bf16 shadow of the volume written by the same epilogue.
                # Measured SLOWER than TF32 maps + cast pass (0.69 vs 0.41 + 0.11 ms at B = 32): the kernel is bound by its band-extraction
                # epilogue (21 scattered stores per pixel and displacement row), which the extra two-byte stores make worse
                self._k("corr_fwd", ops.corr_fwd, L["f1"], L["f2"], L["y"], L["max_disp"], L["stride2"], ACT_ELU, MATH_BF16)
            else:
                self._k("corr_fwd", ops.corr_fwd, L["f1"], L["f2"], L["y"], L["max_disp"], L["stride2"], ACT_ELU,
                        MATH_TF32 if mth == MATH_BF16 else mth)          # the correlation band-GEMMs read the fp32 maps (TF32)
                if mth == MATH_BF16:
                    self._k("cast:corr", ops.cast_bf16, L["y"])

    def _bwd_layer(self, L):
        P, G, mth, mthw = self.params, self.grads, self.math, self.math_wgrad
        if L["op"] == "conv":
            w, dw, db = P[L["wname"] + "/weights"], G[L["wname"] + "/weights"], G[L["wname"] + "/biases"]
            # + bias gradient; bf16 math: only the bf16 shadow of the finished gradient is read again (by this layer's wgrad / dgrad)
            if self.lean and "head_s" in L:     # ... and the input gradient of the flow head reading this slab is added on the fly
                hs = L["head_s"]
                self._k("elu_bwd:" + L["name"], ops.head_dgrad_elu, self.head_d9[hs], self.head_wz[hs], 0, L["dy"], L["y"], L["dy"], L["y"].c, db)
            else:
                self._k("elu_bwd:" + L["name"], ops.elu_bwd, L["dy"], L["y"], db, mth == MATH_BF16 and mthw == MATH_BF16)
            wk = self._k
            if self.lean and self._side_wgrad:      # weight gradient behind the side stream's queue; the input gradient below does not wait for it
                self._fork()
                wk = self._ks
            if L["xpad"] is not None:
                wk("conv_wgrad:" + L["name"], ops.conv1_wgrad, L["g"], L["xpad"], self.x6_origin, L["dy"], dw, None,
                   self._sh.get(id(L["xpad"])) if mthw == MATH_BF16 else None)
            else:
                wk("conv_wgrad:" + L["name"], ops.conv_wgrad, L["g"], L["x"], L["dy"], dw, None, mthw)
            if L["dx"] is not None:
                dx = Slab(L["dx"].t, L["dx"].c0, L["dx"].c) if self.lean else L["dx"]      # (lean: no bf16 shadow of an unfinished gradient)
                self._k("conv_dgrad:" + L["name"], ops.conv_dgrad, L["g"], L["dy"], w, None, dx, ACT_NONE, L["acc"], mth)
        elif L["op"] == "pool":
            self._k("pool_bwd:" + L["name"], ops.maxpool2_bwd, L["x"], L["dy"], L["dx"])
        elif L["op"] == "corr":
            self._k("elu_bwd:corr", ops.elu_bwd, L["dy4"], L["y4"], None)
            self._k("corr_bwd", ops.corr_bwd, L["f1"], L["f2"], L["dy"], L["df1"], L["df2"], L["max_disp"], L["stride2"],
                    mth if (mth != MATH_BF16 or os.environ.get("DOFB_CORR_BWD16", "1") != "0") else MATH_TF32)   # bf16: band-GEMMs on the bf16 shadows

    def forward(self, source: torch.Tensor, target: torch.Tensor, loss_weight=LOSS_WEIGHTS, with_grad: bool = True):
        """flowNet(inputs, outputs, loss_weight): runs the whole forward; when ``with_grad`` the fused
        warp/loss kernel also leaves d(total)/d(pr_s) in self.dpr (the start of the backward)."""
        if tuple(source.shape) != (self.B, self.H, self.W, 3) or tuple(target.shape) != (self.B, self.H, self.W, 3):
            raise ops.DeepOFError(f"expected [B={self.B},{self.H},{self.W},3] NHWC inputs, got {tuple(source.shape)} / {tuple(target.shape)}")
        P, mth = self.params, self.math
        if with_grad and self.lean and self._side_extra and self._side_on():
            # clear the gradient arena (weight gradients accumulate through atomics) on the side stream while the forward runs
            self._fork()
            self._ks("zero_grad", self.grad.zero_)
            self._zero_ev = self._side_event()
        self._preprocess(source, target)
        # (the weight re-pack was tried on the side stream under the pre-processing kernel and conv1: no gain -- its 9.6 k small blocks fill
        # the SMs' thread slots first and the persistent GEMM CTAs wait for them anyway)
        if self.lean:
            sc = range(1, self.N_SCALES + 1)
            self._k("pack_weights", ops.head_wz_pack, [P[f"pr{s}/weights"] for s in sc], [self.head_wz[s] for s in sc])
        if mth != MATH_FP32:        # every layer's tensor-core weight copies in one launch (no-op while they are current)
            self._k("pack_weights", ops.pack_weights_batch, self._pack_jobs(), mth == MATH_BF16)
        for L in self.tower:
            self._fwd_layer(L)
        for R in self.refine:
            s = R["s"]
            x, _ = self.feat[s]
            # lean engine: the flow-head chain of the level (Z GEMM -> tap sum -> up_pr) runs on the side stream next to the transposed
            # convolution; both read feat_s and write disjoint channel slices of feat_{s-1}
            self._fork()
            self._head_fwd(s, x, side=True)
            self._ks("uppr_fwd:" + R["uppr"], ops.uppr_fwd, self.pr[s], P[R["uppr"] + "/weights"], P[R["uppr"] + "/biases"], R["pr_y"])
            self._k("deconv_fwd:" + R["up"], ops.conv_dgrad, R["g"], x, P[R["up"] + "/weights"], P[R["up"] + "/biases"],
                    R["up_y"], ACT_ELU, False, mth)
            self._join()
        self._head_fwd(1, self.feat[1][0])
        lw = [float(v) for v in loss_weight]
        self.loss_weight = lw
        hp = self.hyper
        scales = []
        for s in range(1, self.N_SCALES + 1):
            wgt = lw[s - 1]
            scales.append(dict(flow=self.pr[s], src=self.pyr_src[s], tgt=self.pyr_tgt[s],
                               recon=self.recon1 if s == 1 else None, dflow=self.dpr[s] if with_grad else None,
                               loss4=self.loss4[s - 1], flow_scale=FLOW_SCALES[s], epsilon=hp["epsilon"],
                               alpha_c=hp["alpha_c"], alpha_s=hp["alpha_s"], lambda_smooth=hp["lambda_smooth"],
                               g_charb=wgt, g_u=wgt * hp["lambda_smooth"], g_v=wgt * hp["lambda_smooth"],
                               variant=self.variant))
        self._k("warp_loss", self.warp_loss, scales)

    def outputs(self):
        """(losses, flows_all, prev1) exactly as flowNet returns them (flyingChairsWrapFlow.py:126-129)."""
        keys = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
        losses = [{k: self.loss4[s, i] for i, k in enumerate(keys)} for s in range(self.N_SCALES)]
        flows_all = [self.pr[s] * FLOW_SCALES[s] for s in range(1, self.N_SCALES + 1)]
        return losses, flows_all, self.recon1

    def total_loss(self) -> torch.Tensor:
        """sum_s loss_weight[s] * total_s (flyingChairsWrapFlow.py:122-123) as a device scalar (no host sync)."""
        key = tuple(self.loss_weight)
        if getattr(self, "_lw_key", None) != key:      # the weights live on the device; re-uploaded only when they change
            self._lw_dev = torch.tensor(self.loss_weight, dtype=torch.float32, device=self.device)
            self._lw_key = key
        return torch.dot(self.loss4[:, 0], self._lw_dev)

    # ------------------------------------------------------------------ backward
    def _grad_ready(self, reducer, *names, side=False):
        """Tell the gradient reducer that every parameter at or above the lowest offset of ``names`` has its final gradient
        (``side``: the last writers were enqueued on the side stream -- the all-reduce must order itself behind that stream)."""
        if reducer is not None:
            if side and self._side_on():
                with torch.cuda.stream(self._side):
                    reducer.ready(min(self.arena.offsets[n + "/weights"] for n in names))
            else:
                reducer.ready(min(self.arena.offsets[n + "/weights"] for n in names))

    def _head_wgrad_tc(self, s, side=False):
        """dW_pr_s on the tensor pipe: D9 = bf16 im2col of dpr_s (+ bias gradient), dW = feat_s^T . D9 (1x1 weight-gradient GEMM, feat_s read
        once) accumulated straight into the canonical [3,3,C,2] gradient."""
        G = self.grads
        x, _ = self.feat[s]
        d9 = self.head_d9[s]
        k = self._ks if side else self._k
        k(f"head_dpr9:pr{s}", ops.head_dpr9, self.dpr[s], d9, G[f"pr{s}/biases"])
        ev = self._side_event() if side else None
        k(f"head_wgrad:pr{s}", ops.head_wgrad_tc, x, d9, G[f"pr{s}/weights"])
        return ev

    def _backward_lean(self, reducer=None):
        """Backward of the lean bf16 engine: same order as backward(), but no kernel writes a flow head's input gradient: the pass that
        finishes each channel slab of feat_s (ELU' + bias gradient + bf16 shadow) adds it on the fly (dofb_head_dgrad_elu_bf16)."""
        P, G, mth = self.params, self.grads, self.math
        if self._zero_ev is not None:       # the gradient arena was cleared on the side stream under the forward pass
            self._join_at(self._zero_ev)
            self._zero_ev = None
        else:
            self._k("zero_grad", self.grad.zero_)
        if reducer is not None:
            reducer.begin()
        if self._side_extra:        # D9 of pr1 on the main stream (the first ELU' pass reads it), its weight-gradient GEMM behind it on the side stream
            self._k("head_dpr9:pr1", ops.head_dpr9, self.dpr[1], self.head_d9[1], G["pr1/biases"])
            self._fork()
            self._ks("head_wgrad:pr1", ops.head_wgrad_tc, self.feat[1][0], self.head_d9[1], G["pr1/weights"])
            self._grad_ready(reducer, "pr1", side=True)
        else:
            self._head_wgrad_tc(1)
            self._grad_ready(reducer, "pr1")
        for R in reversed(self.refine):                      # s = 2,3,4,5,6
            s, fs = R["s"], R["s"] - 1
            x, dx = self.feat[s]
            fy, fd = self.feat[fs]
            skipc, upc = R["skipc"], R["upc"]
            # gradient of [upconv | up_pr] outputs inside feat_{s-1}: (deconv_dgrad of the previous iteration, none at scale 1) + head pr_{s-1}
            slab_d = fd.sub(skipc, upc + 2)
            self._k("elu_bwd:" + R["up"], ops.head_dgrad_elu, self.head_d9[fs], self.head_wz[fs], skipc, None if fs == 1 else slab_d,
                    fy.sub(skipc, upc + 2), slab_d, upc, G[R["up"] + "/biases"], self.dpr_up[fs])
            # side stream: up_pr backward -> D9 of pr_s -> head weight gradient (reads feat_s, dpr; writes dpr_s, D9_s, their gradients);
            # main stream: the transposed convolution's weight and input gradients (read the finished slab, write d feat_s)
            self._fork()
            self._ks("uppr_bwd:" + R["uppr"], ops.uppr_bwd, self.pr[s], full(self.dpr_up[fs]), P[R["uppr"] + "/weights"], self.dpr[s],
                     G[R["uppr"] + "/weights"], G[R["uppr"] + "/biases"])
            ev_d9 = self._head_wgrad_tc(s, side=True)        # (event: D9_s written -- all the next ELU' pass needs from the side stream)
            wk = self._ks if self._side_wgrad else self._k
            wk("deconv_wgrad:" + R["up"], ops.conv_wgrad, R["g"], R["up_dy"], x, G[R["up"] + "/weights"], None, mth)
            self._k("deconv_dgrad:" + R["up"], ops.conv_fwd, R["g"], R["up_dy"], P[R["up"] + "/weights"], None, Slab(dx.t, dx.c0, dx.c),
                    ACT_NONE, mth)                           # first (plain-store) writer of d feat_s
            if self._side_wgrad:
                self._join_at(ev_d9)
                self._grad_ready(reducer, f"pr{s}", R["up"], R["uppr"], side=True)
            else:
                self._join()
                self._grad_ready(reducer, f"pr{s}", R["up"], R["uppr"])
        rev = list(reversed(self.tower))
        for i, L in enumerate(rev):
            self._bwd_layer(L)
            if L["op"] == "conv" and all(M.get("wname") != L["wname"] for M in rev[i + 1:]):
                self._grad_ready(reducer, L["wname"], side=self._side_wgrad)
        self._join()                                          # every weight gradient is in before the all-reduce join / Adam

    def backward(self, reducer=None):
        if self.lean:
            return self._backward_lean(reducer)
        P, G, mth, mthw = self.params, self.grads, self.math, self.math_wgrad
        self._k("zero_grad", self.grad.zero_)
        if reducer is not None:
            reducer.begin()
        # refinement part, finest scale first (each pr_s gradient is complete when its scale is reached)
        x1, dx1 = self.feat[1]
        self._k("head_wgrad:pr1", ops.head_wgrad, x1, self.dpr[1], G["pr1/weights"], G["pr1/biases"])
        self._k("head_dgrad:pr1", ops.head_dgrad, self.dpr[1], P["pr1/weights"], dx1, accumulate=False)
        self._grad_ready(reducer, "pr1")
        for R in reversed(self.refine):                      # s = 2,3,4,5,6
            s = R["s"]
            x, dx = self.feat[s]
            # up_pr (linear): dpr_s += ..., dW, db
            self._k("uppr_bwd:" + R["uppr"], ops.uppr_bwd, self.pr[s], R["pr_dy"], P[R["uppr"] + "/weights"], self.dpr[s],
                    G[R["uppr"] + "/weights"], G[R["uppr"] + "/biases"])
            # upconv (ELU): gradient through the activation, then weight / bias / input gradients
            self._k("elu_bwd:" + R["up"], ops.elu_bwd, R["up_dy"], R["up_y"], G[R["up"] + "/biases"],           # + bias gradient
                    mth == MATH_BF16 and mthw == MATH_BF16)
            self._k("deconv_wgrad:" + R["up"], ops.conv_wgrad, R["g"], R["up_dy"], x, G[R["up"] + "/weights"], None, mthw)
            self._k("deconv_dgrad:" + R["up"], ops.conv_fwd, R["g"], R["up_dy"], P[R["up"] + "/weights"], None, dx, ACT_NONE,
                    mth)                                                   # first writer of d feat_s
            # pr_s head (its read-modify-write of d feat_s streams through a cp.async ring; measured faster than letting the GEMM
            # epilogue of the deconv accumulate behind a pure-store head: 11.56 vs 11.76 ms per step)
            self._k(f"head_wgrad:pr{s}", ops.head_wgrad, x, self.dpr[s], G[f"pr{s}/weights"], G[f"pr{s}/biases"])
            self._k(f"head_dgrad:pr{s}", ops.head_dgrad, self.dpr[s], P[f"pr{s}/weights"], dx, accumulate=True)
            self._grad_ready(reducer, f"pr{s}", R["up"], R["uppr"])
        # contracting tower, top down (shared siamese weights are final after their LAST use, which the reverse order reaches last)
        rev = list(reversed(self.tower))
        for i, L in enumerate(rev):
            self._bwd_layer(L)
            if L["op"] == "conv" and all(M.get("wname") != L["wname"] for M in rev[i + 1:]):
                self._grad_ready(reducer, L["wname"])

    # ------------------------------------------------------------------ optimiser
    def adam_step(self, lr: float, grad_scale: float = 1.0, beta1=0.9, beta2=0.999, eps=1e-8):
        """TF-form Adam over the flat arena (flyingChairsTrain.py:124)."""
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - beta2 ** self.t) / (1.0 - beta1 ** self.t)
        self._k("adam", ops.adam, self.theta, self.grad, self.m, self.v, lr_t, beta1, beta2, eps, grad_scale)
        ops.invalidate_weight_cache()           # the tcgen05 path re-packs its K-major weight copies lazily

    def train_step(self, source, target, loss_weight=LOSS_WEIGHTS, lr: float = 1.6e-5, allreduce=None):
        """One ``train_op.run(feed_dict)`` (flyingChairsTrain.py:178): forward, backward, Adam."""
        self.forward(source, target, loss_weight, with_grad=True)
        self.backward(reducer=allreduce)
        scale = allreduce.finish() if allreduce is not None else 1.0
        self.adam_step(lr, grad_scale=scale)


# ---------------------------------------------------------------------------------------------------------------------
# FlowNetC: siamese conv1-3 + correlation cost volume (FlowNet paper; the reference has NO such model -- SURVEY.md 0.2 --
# so this architecture is specified here and its parity is pinned only against our own oracle, oracle/flownet_c.py).
#   conv1 7x7/2 3->64, conv2 5x5/2 64->128, conv3 5x5/2 128->256 on each image (shared weights)
#   corr(conv3a, conv3b): max displacement 20, stride2 2 -> 21x21 = 441 channels, /C, ELU
#   conv_redir 1x1 256->32 (ELU) on conv3a ; concat [conv_redir, corr] = 473 -> conv3_1 3x3 -> 256
#   conv4_1 ... conv6_2 and the refinement exactly as FlowNetS (skips: conv5_2, conv4_2, conv3_1, conv2a, conv1a)
# ---------------------------------------------------------------------------------------------------------------------
CORR_MAX_DISP, CORR_STRIDE2 = 20, 2


def param_shapes_c() -> "OrderedDict[str, tuple]":
    sh: "OrderedDict[str, tuple]" = OrderedDict()
    for name, k, cin, cout in [("conv1", 7, 3, 64), ("conv2", 5, 64, 128), ("conv3", 5, 128, 256), ("conv_redir", 1, 256, 32),
                               ("conv3_1", 3, 473, 256), ("conv4_1", 3, 256, 512), ("conv4_2", 3, 512, 512), ("conv5_1", 3, 512, 512),
                               ("conv5_2", 3, 512, 512), ("conv6_1", 3, 512, 1024), ("conv6_2", 3, 1024, 1024)]:
        sh[name + "/weights"] = (k, k, cin, cout)
        sh[name + "/biases"] = (cout,)
    for key, shape in param_shapes().items():
        if key.startswith(("pr", "up")):
            sh[key] = shape
    return sh


class FlowNetC(FlowNetS):
    ARCH = "C"
    NEED32 = frozenset({"c3a", "c3b", "cat3"})      # the correlation band-GEMMs read / write fp32 maps (TF32)

    @staticmethod
    def param_shapes():
        return param_shapes_c()

    def _plan_tower(self):
        a, d, S, R, full = self.act, self.dact, self._S, self._conv_rec, self._F
        fp32 = self.math == MATH_FP32
        xa = S(self.x6, 0, 3) if fp32 else None
        xb = S(self.x6b, 0, 3) if fp32 else None
        pa = None if fp32 else self.x6
        pb = None if fp32 else self.x6b
        D = 2 * (CORR_MAX_DISP // CORR_STRIDE2) + 1
        corr = dict(op="corr", name="corr", f1=full(a["c3a"]), f2=full(a["c3b"]), y=S(a["cat3"], 32, D * D), dy=S(d["cat3"], 32, D * D),
                    y4=S(a["cat3"], 32, 444), dy4=S(d["cat3"], 32, 444),       # ELU' runs on a multiple of 4 channels (pad = 0)
                    df1=full(d["c3a"]), df2=full(d["c3b"]), max_disp=CORR_MAX_DISP, stride2=CORR_STRIDE2)
        return [
            R("conv1a", 7, 2, 3, 64, xa, None, S(a["concat1"], 0, 64), S(d["concat1"], 0, 64), wname="conv1", xpad=pa),
            R("conv1b", 7, 2, 3, 64, xb, None, full(a["c1b"]), full(d["c1b"]), wname="conv1", xpad=pb),
            R("conv2a", 5, 2, 64, 128, S(a["concat1"], 0, 64), S(d["concat1"], 0, 64), S(a["concat2"], 0, 128), S(d["concat2"], 0, 128),
              acc=True, wname="conv2"),
            R("conv2b", 5, 2, 64, 128, full(a["c1b"]), full(d["c1b"]), full(a["c2b"]), full(d["c2b"]), wname="conv2"),
            R("conv3a", 5, 2, 128, 256, S(a["concat2"], 0, 128), S(d["concat2"], 0, 128), full(a["c3a"]), full(d["c3a"]), acc=True, wname="conv3"),
            R("conv3b", 5, 2, 128, 256, full(a["c2b"]), full(d["c2b"]), full(a["c3b"]), full(d["c3b"]), wname="conv3"),
            # backward runs in reverse: conv3_1 (writes d cat3) -> corr (overwrites d c3a, d c3b) -> conv_redir (accumulates d c3a)
            R("conv_redir", 1, 1, 256, 32, full(a["c3a"]), full(d["c3a"]), S(a["cat3"], 0, 32), S(d["cat3"], 0, 32), acc=True),
            corr,
            R("conv3_1", 3, 1, 473, 256, S(a["cat3"], 0, 473), S(d["cat3"], 0, 473), S(a["concat3"], 0, 256), S(d["concat3"], 0, 256)),
        ] + self._plan_tower_top()


# ---------------------------------------------------------------------------------------------------------------------
# VGG16 guided model: flyingChairsWrapFlow_vgg.VGG16 (flyingChairsWrapFlow_vgg.py:7-132) -- what deepOF_fc.py
# really launches.  13 3x3 ELU convs + 5 2x2 max-pools, the same decoder at 5 scales (pool outputs are the skips), loss variant B,
# separate photo (network input) and geo (loss images) pairs, all pre-scaled by the trainer (flyingChairsTrain_vgg.py:181-188).
# ---------------------------------------------------------------------------------------------------------------------
VGG_CONVS = [("conv1_1", 6, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128), ("conv3_1", 128, 256),
             ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
             ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)]
VGG_REFINE = [(5, 512, "upconv4", 256, "up_pr5to4", 512), (4, 770, "upconv3", 128, "up_pr4to3", 256),
              (3, 386, "upconv2", 64, "up_pr3to2", 128), (2, 194, "upconv1", 32, "up_pr2to1", 64)]
VGG_LOSS_WEIGHTS = (16.0, 8.0, 4.0, 2.0, 1.0)         # flyingChairsTrain_vgg.py:171


def param_shapes_vgg() -> "OrderedDict[str, tuple]":
    sh: "OrderedDict[str, tuple]" = OrderedDict()
    for name, cin, cout in VGG_CONVS:
        sh[name + "/weights"] = (3, 3, cin, cout)
        sh[name + "/biases"] = (cout,)
    for s, cfeat, up, upc, uppr, _skip in VGG_REFINE:
        sh[f"pr{s}/weights"] = (3, 3, cfeat, 2)
        sh[f"pr{s}/biases"] = (2,)
        sh[up + "/weights"] = (4, 4, upc, cfeat)
        sh[up + "/biases"] = (upc,)
        sh[uppr + "/weights"] = (4, 4, 2, 2)
        sh[uppr + "/biases"] = (2,)
    sh["pr1/weights"] = (3, 3, 98, 2)
    sh["pr1/biases"] = (2,)
    return sh


class VGG16Flow(FlowNetS):
    ARCH = "V"
    N_SCALES = 5
    REFINE_SPEC = VGG_REFINE

    def __init__(self, batch, height=320, width=448, device="cuda", variant="B", **kw):
        super().__init__(batch, height, width, device=device, variant=variant, **kw)

    @staticmethod
    def param_shapes():
        return param_shapes_vgg()

    def _buffer_shapes(self):
        H, W = self.H, self.W
        shp = {"concat1": (H // 2, W // 2, 128), "concat2": (H // 4, W // 4, 224), "concat3": (H // 8, W // 8, 416),
               "concat4": (H // 16, W // 16, 800), "pool5": (H // 32, W // 32, 512)}
        for lvl, (div, c, n) in enumerate([(1, 64, 2), (2, 128, 2), (4, 256, 3), (8, 512, 3), (16, 512, 3)], start=1):
            for i in range(1, n + 1):
                shp[f"c{lvl}{i}"] = (H // div, W // div, c)
        return shp

    def _plan_tower(self):
        a, d, S, R, full = self.act, self.dact, self._S, self._conv_rec, self._F
        recs = []
        skip_c = {1: 64, 2: 128, 3: 256, 4: 512}
        prev, dprev, prev_acc = S(self.x6, 0, 6), None, False
        for lvl, n in enumerate([2, 2, 3, 3, 3], start=1):
            for i in range(1, n + 1):
                name = f"conv{lvl}_{i}"
                cin, cout = next((ci, co) for nm, ci, co in VGG_CONVS if nm == name)
                y, dy = full(a[f"c{lvl}{i}"]), full(d[f"c{lvl}{i}"])
                recs.append(R(name, 3, 1, cin, cout, prev, dprev, y, dy, acc=prev_acc))
                prev, dprev, prev_acc = y, dy, False
            if lvl < 5:
                py, pdy = S(a[f"concat{lvl}"], 0, skip_c[lvl]), S(d[f"concat{lvl}"], 0, skip_c[lvl])
            else:
                py, pdy = full(a["pool5"]), full(d["pool5"])
            recs.append(dict(op="pool", name=f"pool{lvl}", x=prev, dx=dprev, y=py, dy=pdy))
            prev, dprev, prev_acc = py, pdy, True        # the next conv's input gradient ADDS to the decoder's contributions
        return recs

    def _plan_feat(self):
        a, d, S, full = self.act, self.dact, self._S, self._F
        return {5: (full(a["pool5"]), full(d["pool5"])), 4: (S(a["concat4"], 0, 770), S(d["concat4"], 0, 770)),
                3: (S(a["concat3"], 0, 386), S(d["concat3"], 0, 386)), 2: (S(a["concat2"], 0, 194), S(d["concat2"], 0, 194)),
                1: (S(a["concat1"], 0, 98), S(d["concat1"], 0, 98))}

    def forward(self, photo_source, photo_target, loss_weight=VGG_LOSS_WEIGHTS, with_grad=True, geo_source=None, geo_target=None,
                prescaled=True):
        """VGG16(photo_source, photo_target, geo_source, geo_target, loss_weight) (flyingChairsWrapFlow_vgg.py:7).
        ``prescaled`` (the reference's contract): the four images are already (x - mean)/255."""
        self._geo = (geo_source if geo_source is not None else photo_source, geo_target if geo_target is not None else photo_target)
        self._prescaled = prescaled
        return super().forward(photo_source, photo_target, loss_weight, with_grad)

    def _preprocess(self, source, target):
        mean, div = ((0.0, 0.0, 0.0), 1.0) if self._prescaled else (self.mean, 255.0)
        n = self.N_SCALES
        self._k("preprocess", ops.preprocess, source, target, mean, self.x6, [], [], (0, 0), None, div)
        if self.math == MATH_BF16:
            self._k("cast:x6", ops.cast_bf16, self._F(self.x6))
        self._k("preprocess", ops.preprocess, self._geo[0], self._geo[1], mean, None, [self.pyr_src[s] for s in range(1, n + 1)],
                [self.pyr_tgt[s] for s in range(1, n + 1)], (0, 0), None, div)
