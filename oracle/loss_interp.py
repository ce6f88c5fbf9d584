"""Warp + Charbonnier photometric + smoothness loss -- CPU oracle (torch, autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference
has no golden vectors for this op; the formulas below follow the cited lines.

Two non-equivalent variants exist in the reference:

* variant "A" -- ``flyingChairsWrapFlow.loss_interp`` (flyingChairsWrapFlow.py:752-876):
  smoothness = a dense 3x3 conv of the SCALED flow with the short-list
  ``FlowDeltaWeights`` constant (:48), which only ever reads U; denominators are
  the image valid-pixel count; no border mask on the smoothness term.
* variant "B" -- ``flyingChairsWrapFlow_vgg.loss_interp`` (:135-317) ==
  ``version1/model/warpflow.loss_interp`` (:4-173): depthwise forward differences
  of the UN-scaled flow, border mask applied, denominator N*2/3.

The warp itself (:785-838 == warpflow.py:36-89) is common to both.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .tf_ops import tf_constant_fill

# flyingChairsWrapFlow.py:48 -- 18 values poured into a [3,3,2,2] constant.
FLOW_DELTA_VALUES = [0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0]


def flow_delta_weights(dtype=torch.float32) -> torch.Tensor:
    return tf_constant_fill([float(v) for v in FLOW_DELTA_VALUES], (3, 3, 2, 2), dtype)


def border_width(height: int, border_ratio: float = 0.1) -> int:
    """flyingChairsWrapFlow.py:764-766: ceil(height * 0.1), same width for rows and cols."""
    return int(math.ceil(height * border_ratio))


def border_mask(height: int, width: int, dtype=torch.float32) -> torch.Tensor:
    """[h,w] mask, 1 inside a frame of ``border_width`` rows/cols (:767-768)."""
    bw = border_width(height)
    m = torch.zeros(height, width, dtype=dtype)
    if height - 2 * bw > 0 and width - 2 * bw > 0:
        m[bw:height - bw, bw:width - bw] = 1
    return m


def smoothness_mask(height: int, width: int, dtype=torch.float32) -> torch.Tensor:
    """[h,w,2]: ch0 zero in the last column, ch1 zero in the last row (:773-778)."""
    m = torch.ones(height, width, 2, dtype=dtype)
    m[:, width - 1, 0] = 0
    m[height - 1, :, 1] = 0
    return m


def warp(flows_scaled: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Backward-warp ``target`` by the (already scaled) flow; :785-838.

    flows_scaled [B,h,w,2] (ch0 = U horizontal, ch1 = V vertical),
    target [B,h,w,C] -> reconstruction [B,h,w,C].
    floor/frac split; x and y clamped INDEPENDENTLY to the image (:815-818);
    gradient flows only through the fractional weights (floor has zero grad).
    """
    B, h, w, C = target.shape
    fl = torch.floor(flows_scaled)
    frac = flows_scaled - fl
    xi = fl[..., 0].to(torch.int64)
    yi = fl[..., 1].to(torch.int64)
    xw = frac[..., 0].unsqueeze(-1)
    yw = frac[..., 1].unsqueeze(-1)
    cols = torch.arange(w).view(1, 1, w)
    rows = torch.arange(h).view(1, h, 1)
    x0 = torch.clamp(cols + xi, 0, w - 1)
    x1 = torch.clamp(cols + xi + 1, 0, w - 1)
    y0 = torch.clamp(rows + yi, 0, h - 1)
    y1 = torch.clamp(rows + yi + 1, 0, h - 1)
    flat = target.reshape(B, h * w, C)

    def gather(yy, xx):
        idx = (yy * w + xx).reshape(B, h * w, 1).expand(B, h * w, C)
        return torch.gather(flat, 1, idx).reshape(B, h, w, C)

    Ia = gather(y0, x0)
    Ib = gather(y1, x0)
    Ic = gather(y0, x1)
    Id = gather(y1, x1)
    wa = (1 - xw) * (1 - yw)
    wb = (1 - xw) * yw
    wc = xw * (1 - yw)
    wd = xw * yw
    return Ia * wa + Ib * wb + Ic * wc + Id * wd


def photometric(recon: torch.Tensor, source: torch.Tensor, epsilon: float, alpha_c: float):
    """:841-849 -- ((255*(recon-src))^2 + eps^2)^alpha_c, border-masked mean.

    Returns (Charbonnier_reconstruct, numValidPixels)."""
    B, h, w, C = source.shape
    diff = 255.0 * (recon - source)
    ele = torch.pow(diff * diff + epsilon * epsilon, alpha_c)
    mask = border_mask(h, w, source.dtype).view(1, h, w, 1)
    n_valid = float(B * C) * float(mask.sum().item())
    return (ele * mask).sum() / n_valid, n_valid


def loss_interp_A(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """flyingChairsWrapFlow.loss_interp (flyingChairsWrapFlow.py:752-876), literal.

    Returns (lossDict, reconstructs[B,h,w,3])."""
    B, h, w, C = inputs.shape
    scaled = flows * flow_scale                      # :783 (re-binds `flows`)
    recon = warp(scaled, outputs)
    charb, n_valid = photometric(recon, inputs, epsilon, alpha_c)
    # :854 dense conv of the SCALED flow with the short-list constant
    wdelta = flow_delta_weights(flows.dtype) if deltaWeights is None else deltaWeights
    fd = F.conv2d(scaled.permute(0, 3, 1, 2), wdelta.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    clean = fd * smoothness_mask(h, w, flows.dtype).unsqueeze(0)
    e2 = epsilon * epsilon
    u_loss = torch.pow(clean[..., 0] ** 2 + e2, alpha_s).sum() / n_valid
    v_loss = torch.pow(clean[..., 1] ** 2 + e2, alpha_s).sum() / n_valid
    total = charb + lambda_smooth * (u_loss + v_loss)
    return {"total": total, "Charbonnier_reconstruct": charb, "U_loss": u_loss, "V_loss": v_loss}, recon


def loss_interp_B(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """version1/model/warpflow.loss_interp (:4-173) with needMask=True,
    needImageGradients=False (version1/model/Flownet.py:71-84)."""
    B, h, w, C = inputs.shape
    scaled = flows * flow_scale                      # warpflow.py:36
    recon = warp(scaled, outputs)
    charb, n_valid = photometric(recon, inputs, epsilon, alpha_c)
    # warpflow.py:133-136 depthwise forward differences of the UN-scaled flow, zero pad
    fpad = F.pad(flows, (0, 0, 0, 1, 0, 1))          # pad w and h by one at the end
    hgrad = flows - fpad[:, :h, 1:w + 1, :]          # F[y,x] - F[y,x+1]
    vgrad = flows - fpad[:, 1:h + 1, :w, :]          # F[y,x] - F[y+1,x]
    sm = smoothness_mask(h, w, flows.dtype).unsqueeze(0)
    bm = border_mask(h, w, flows.dtype).view(1, h, w, 1)
    n_flow = n_valid / 3 * 2                         # :140
    e2 = epsilon * epsilon
    u_delta = torch.stack([hgrad[..., 0], vgrad[..., 0]], dim=3) * sm
    v_delta = torch.stack([hgrad[..., 1], vgrad[..., 1]], dim=3) * sm
    u_loss = (torch.pow(u_delta ** 2 + e2, alpha_s) * bm).sum() / n_flow
    v_loss = (torch.pow(v_delta ** 2 + e2, alpha_s) * bm).sum() / n_flow
    total = charb + lambda_smooth * (u_loss + v_loss)
    return {"total": total, "Charbonnier_reconstruct": charb, "U_loss": u_loss, "V_loss": v_loss}, recon


def loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None, variant="A"):
    fn = loss_interp_A if variant == "A" else loss_interp_B
    return fn(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights)
