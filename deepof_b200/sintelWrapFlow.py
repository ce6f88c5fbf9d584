"""Drop-in for the multi-frame loss of ``sintelWrapFlow.py``: ``loss_interp_multi`` (sintelWrapFlow.py:492-630).

Same name, argument order and return structure; operates on CUDA tensors and runs the sm_100a kernel (dofb_warp_loss_multi).
The two-frame ``loss_interp`` of that file (:632-766) is the variant-A loss of ``deepof_b200.flyingChairsWrapFlow``."""
from __future__ import annotations

import torch

from . import ops
from .flyingChairsWrapFlow import loss_interp  # noqa: F401  (sintelWrapFlow.loss_interp == flyingChairsWrapFlow.loss_interp)
from .flownet import SINTEL_MEAN  # noqa: F401

_KEYS = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
FLOW_DELTA_VALUES = (0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0)          # sintelWrapFlow.py:378


def flow_delta_weights(flow_channels: int) -> torch.Tensor:
    """tf.constant(list, shape=[3,3,Cf,Cf]) (:378): the 18 values fill the tensor in row-major order, the remainder takes the last value (0)."""
    w = torch.zeros(3 * 3 * flow_channels * flow_channels)
    n = min(len(FLOW_DELTA_VALUES), w.numel())
    w[:n] = torch.tensor(FLOW_DELTA_VALUES[:n], dtype=torch.float32)
    w[n:] = float(FLOW_DELTA_VALUES[-1])
    return w.view(3, 3, flow_channels, flow_channels)


class _LossInterpMultiFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flows, inputs, cfg, stencil):
        flows, inputs = flows.contiguous(), inputs.contiguous()
        loss4, recon, _ = ops.warp_loss_multi(flows, inputs, stencil, want_grad=False, **cfg)
        ctx.save_for_backward(flows, inputs)
        ctx.cfg, ctx.stencil = cfg, stencil
        ctx.mark_non_differentiable(recon)
        return loss4[0], loss4[1], loss4[2], loss4[3], recon

    @staticmethod
    def backward(ctx, g_total, g_charb, g_u, g_v, _g_recon):
        flows, inputs = ctx.saved_tensors
        g = [float(t) if t is not None else 0.0 for t in (g_total, g_charb, g_u, g_v)]   # one host sync
        lam = ctx.cfg["lambda_smooth"]
        _l4, _r, dflow = ops.warp_loss_multi(flows, inputs, ctx.stencil, g=(g[0] + g[1], g[0] * lam + g[2], g[0] * lam + g[3]),
                                             want_recon=False, want_grad=True, **ctx.cfg)
        return dflow, None, None, None


def loss_interp_multi(flows, inputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """loss_interp_multi(flows, inputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights) -> (lossDict, reconstructs).

    inputs [B,h,w,3T]: T frames stacked on the channel axis; flows [B,h,w,2(T-1)].  ``deltaWeights["FlowDeltaWeights"]`` is the dense
    [3,3,2(T-1),2(T-1)] smoothness constant (default: the reference's, :378)."""
    cf = flows.shape[3]
    dw = (deltaWeights or {}).get("FlowDeltaWeights") if isinstance(deltaWeights, dict) else deltaWeights
    stencil = ops.flow_stencil(dw if dw is not None else flow_delta_weights(cf))
    cfg = dict(flow_scale=float(flow_scale), epsilon=float(epsilon), alpha_c=float(alpha_c), alpha_s=float(alpha_s),
               lambda_smooth=float(lambda_smooth))
    t, c, u, v, recon = _LossInterpMultiFn.apply(flows, inputs, cfg, stencil)
    return dict(zip(_KEYS, (t, c, u, v))), recon
