"""End-point-error metric and the evaluation recipe -- CPU oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).
``flow_ee`` follows utils.py:64-68; ``eval_flow`` follows
flyingChairsTrain.py:264-266 (x2, clip, cv2 bilinear resize to the ground-truth size).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def flow_ee(flow: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """AEE = mean over all pixels and samples of sqrt(du^2 + dv^2)  (utils.py:64-68)."""
    d = flow - gt
    return torch.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2).mean()


def eval_flow(flow_scaled_s1: torch.Tensor, out_h: int, out_w: int, clip=(-300.0, 250.0)) -> torch.Tensor:
    """flows_all[0] * 2 -> clip -> bilinear resize to (out_h,out_w).

    cv2.resize(INTER_LINEAR) uses half-pixel centres with edge clamp, which is
    torch's bilinear/align_corners=False for up-sampling."""
    f = torch.clamp(flow_scaled_s1 * 2.0, clip[0], clip[1])
    f = F.interpolate(f.permute(0, 3, 1, 2), size=(out_h, out_w), mode="bilinear", align_corners=False)
    return f.permute(0, 2, 3, 1).contiguous()
