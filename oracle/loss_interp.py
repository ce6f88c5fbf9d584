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


# ---------------------------------------------------------------------------------------------------------------------
# Edge-aware smoothness: version1/model/warpflow.py:91-116,148-157 (needImageGradients=True)
# ---------------------------------------------------------------------------------------------------------------------
SOBEL_X = [[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]       # version1/model/Flownet.py:87-90 (sobel_y = its transpose)
GRAY_WEIGHTS = (0.2989, 0.5870, 0.1140)                                 # tf.image.rgb_to_grayscale, applied to channels 0,1,2 as stored


def edge_weights(inputs: torch.Tensor) -> torch.Tensor:
    """gradientsMaskFlow [B,h,w,2] = (1 - |sobel_x|/max|sobel_x|, 1 - |sobel_y|/max|sobel_y|) of the re-quantised grayscale image.

    warpflow.py:92-116: every image of the batch is stretched to 0..255 with its OWN min/max over all pixels and channels (:95-99),
    truncated to int32 and clipped (:100), converted to grayscale (:105; TF converts int32 images through float and truncates back),
    filtered with the 3x3 Sobel pair (depthwise, SAME zero padding, :107-108) and each response is divided by its maximum magnitude
    over the WHOLE batch (:109-110); eta = 1 (:113).  No gradient flows into the images (placeholders)."""
    B, h, w, C = inputs.shape
    x = inputs.detach().double()
    mn = x.amin(dim=(1, 2, 3), keepdim=True)
    mx = x.amax(dim=(1, 2, 3), keepdim=True)
    q = torch.clamp(torch.trunc(255.0 * (x - mn) / (mx - mn)), 0, 255)
    gray = torch.trunc(q[..., 0] * GRAY_WEIGHTS[0] + q[..., 1] * GRAY_WEIGHTS[1] + q[..., 2] * GRAY_WEIGHTS[2])      # [B,h,w]
    kx = torch.tensor(SOBEL_X, dtype=torch.float64).view(1, 1, 3, 3)
    gx = F.conv2d(gray.unsqueeze(1), kx, padding=1)[:, 0]
    gy = F.conv2d(gray.unsqueeze(1), kx.transpose(2, 3), padding=1)[:, 0]
    gx = gx / gx.abs().max()
    gy = gy / gy.abs().max()
    return torch.stack([1.0 - gx.abs(), 1.0 - gy.abs()], dim=3).to(inputs.dtype)


def loss_interp_B_edge(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """warpflow.loss_interp with needMask=True and needImageGradients=True (:148-157): the element-wise smoothness losses of U and V
    ([h-delta, v-delta] channels) are multiplied by gradientsMaskFlow before the border mask."""
    B, h, w, C = inputs.shape
    scaled = flows * flow_scale
    recon = warp(scaled, outputs)
    charb, n_valid = photometric(recon, inputs, epsilon, alpha_c)
    fpad = F.pad(flows, (0, 0, 0, 1, 0, 1))
    hgrad = flows - fpad[:, :h, 1:w + 1, :]
    vgrad = flows - fpad[:, 1:h + 1, :w, :]
    sm = smoothness_mask(h, w, flows.dtype).unsqueeze(0)
    bm = border_mask(h, w, flows.dtype).view(1, h, w, 1)
    ew = edge_weights(inputs)
    n_flow = n_valid / 3 * 2
    e2 = epsilon * epsilon
    u_delta = torch.stack([hgrad[..., 0], vgrad[..., 0]], dim=3) * sm
    v_delta = torch.stack([hgrad[..., 1], vgrad[..., 1]], dim=3) * sm
    u_loss = (torch.pow(u_delta ** 2 + e2, alpha_s) * ew * bm).sum() / n_flow
    v_loss = (torch.pow(v_delta ** 2 + e2, alpha_s) * ew * bm).sum() / n_flow
    total = charb + lambda_smooth * (u_loss + v_loss)
    return {"total": total, "Charbonnier_reconstruct": charb, "U_loss": u_loss, "V_loss": v_loss}, recon


# ---------------------------------------------------------------------------------------------------------------------
# Multi-frame warp / loss: sintelWrapFlow.loss_interp_multi (sintelWrapFlow.py:492-630)
# ---------------------------------------------------------------------------------------------------------------------
def flow_delta_weights_multi(flow_channels: int, dtype=torch.float32) -> torch.Tensor:
    """sintelWrapFlow.py:378: the same 18-value list poured into a [3,3,Cf,Cf] constant (row-major, remainder = last value 0)."""
    return tf_constant_fill([float(v) for v in FLOW_DELTA_VALUES], (3, 3, flow_channels, flow_channels), dtype)


def loss_interp_multi(flows, inputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """sintelWrapFlow.loss_interp_multi (:492-630), literal.

    inputs [B,h,w,3T] (T frames stacked on channels), flows [B,h,w,2(T-1)].  Channel c < 3(T-1) of the reconstruction is channel c+3
    (the NEXT frame) warped by flow pair 2*(c//3), 2*(c//3)+1 (:544-571) and is compared with channel c (:581).  Smoothness: dense 3x3
    conv of the scaled flows with deltaWeights["FlowDeltaWeights"] (:606), smoothness mask then border mask BEFORE the pow (:612-613),
    even channels -> U_loss, odd -> V_loss, both divided by the image N (:614-617).
    Returns (lossDict, reconstructs [B,h,w,3(T-1)])."""
    B, h, w, C = inputs.shape
    P = C // 3 - 1
    assert flows.shape[3] == 2 * P
    scaled = flows * flow_scale                                       # :526
    recs = []
    for k in range(P):
        recs.append(warp(scaled[..., 2 * k:2 * k + 2], inputs[..., 3 * (k + 1):3 * (k + 2)]))
    recon = torch.cat(recs, dim=3)
    diff = 255.0 * (recon - inputs[..., :C - 3])
    ele = torch.pow(diff * diff + epsilon * epsilon, alpha_c)
    bmask = border_mask(h, w, inputs.dtype).view(1, h, w, 1)
    n_valid = float(B * (C - 3)) * float(bmask.sum().item())
    charb = (ele * bmask).sum() / n_valid
    wdelta = flow_delta_weights_multi(2 * P, flows.dtype) if deltaWeights is None else deltaWeights
    fd = F.conv2d(scaled.permute(0, 3, 1, 2), wdelta.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    sm = smoothness_mask(h, w, flows.dtype).unsqueeze(0).repeat(1, 1, 1, P)      # [.., (masky, maskx) * P]  (:517-518)
    clean = fd * sm * bmask
    e2 = epsilon * epsilon
    u_loss = torch.pow(clean[..., 0::2] ** 2 + e2, alpha_s).sum() / n_valid
    v_loss = torch.pow(clean[..., 1::2] ** 2 + e2, alpha_s).sum() / n_valid
    total = charb + lambda_smooth * (u_loss + v_loss)
    return {"total": total, "Charbonnier_reconstruct": charb, "U_loss": u_loss, "V_loss": v_loss}, recon
