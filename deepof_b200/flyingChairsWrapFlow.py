"""Drop-in for the reference module ``flyingChairsWrapFlow`` (hot-path subset).

Same names, argument order and return structure as the reference
(flyingChairsWrapFlow.py:5 ``flowNet``, :752 ``loss_interp``), operating on CUDA tensors instead
of TF graph nodes: calling them *runs* the sm_100a kernels.
"""
from __future__ import annotations

import torch

from . import ops
from .flownet import FlowNetS, LOSS_WEIGHTS, HYPER, FLOW_SCALES  # noqa: F401

_KEYS = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
_warp_loss_cache: dict = {}


def _wl(device):
    key = str(device)
    if key not in _warp_loss_cache:
        _warp_loss_cache[key] = ops.WarpLoss(device)
    return _warp_loss_cache[key]


class _LossInterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, variant, edge_w=None):
        flows, inputs, outputs = flows.contiguous(), inputs.contiguous(), outputs.contiguous()
        loss4 = torch.empty(4, dtype=torch.float32, device=flows.device)
        recon = torch.empty_like(inputs)
        cfg = dict(flow_scale=flow_scale, epsilon=epsilon, alpha_c=alpha_c, alpha_s=alpha_s, lambda_smooth=lambda_smooth,
                   variant=variant, edge_w=edge_w)
        _wl(flows.device)([dict(flow=flows, src=inputs, tgt=outputs, recon=recon, dflow=None, loss4=loss4, **cfg)])
        ctx.save_for_backward(flows, inputs, outputs)
        ctx.cfg = cfg
        ctx.mark_non_differentiable(recon)
        return loss4[0], loss4[1], loss4[2], loss4[3], recon

    @staticmethod
    def backward(ctx, g_total, g_charb, g_u, g_v, _g_recon):
        flows, inputs, outputs = ctx.saved_tensors
        cfg = ctx.cfg
        g = [float(t) if t is not None else 0.0 for t in (g_total, g_charb, g_u, g_v)]   # one host sync
        lam = cfg["lambda_smooth"]
        dflow = torch.empty_like(flows)
        scratch = torch.empty(4, dtype=torch.float32, device=flows.device)
        _wl(flows.device)([dict(flow=flows, src=inputs, tgt=outputs, recon=None, dflow=dflow, loss4=scratch,
                                g_charb=g[0] + g[1], g_u=g[0] * lam + g[2], g_v=g[0] * lam + g[3], **cfg)])
        return dflow, None, None, None, None, None, None, None, None, None


def _loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, variant, edge_w=None):
    t, c, u, v, recon = _LossInterpFn.apply(flows, inputs, outputs, float(epsilon), float(alpha_c), float(alpha_s),
                                            float(lambda_smooth), float(flow_scale), variant, edge_w)
    return dict(zip(_KEYS, (t, c, u, v))), recon


def loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """flyingChairsWrapFlow.loss_interp (flyingChairsWrapFlow.py:752): legacy smoothness (variant A).

    ``deltaWeights`` is accepted for signature compatibility; the FlowDeltaWeights constant of :48
    is baked into the kernel (it is a compile-time constant in the reference too)."""
    return _loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, 0)


_engines: dict = {}


def get_engine(batch, height, width, device="cuda", variant="A", math_mode="fp32", **kw) -> FlowNetS:
    """The engine behind flowNet for a given static shape (TF builds one graph per shape, too)."""
    key = (batch, height, width, str(device), variant, math_mode)
    if key not in _engines:
        _engines[key] = FlowNetS(batch, height, width, device=device, variant=variant, math_mode=math_mode, **kw)
    return _engines[key]


def flowNet(inputs, outputs, loss_weight, engine: FlowNetS | None = None):
    """flowNet(inputs, outputs, loss_weight) -> (losses, flows_all, prev1)  (flyingChairsWrapFlow.py:5-129).

    inputs/outputs: [B,H,W,3] BGR 0..255 float32 CUDA tensors; loss_weight: 6 floats."""
    B, H, W, _ = inputs.shape
    eng = engine or get_engine(B, H, W, device=inputs.device)
    lw = loss_weight.tolist() if isinstance(loss_weight, torch.Tensor) else list(loss_weight)
    eng.forward(inputs.contiguous(), outputs.contiguous(), lw, with_grad=False)
    return eng.outputs()
