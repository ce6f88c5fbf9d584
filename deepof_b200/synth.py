"""Synthetic FlyingChairs-shaped inputs.  Shared by tests, smoke and bench.

Not reference code: the reference trains on the real dataset, which is not
available offline.  Deterministic in (seed, B, H, W)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _box(x_nchw: torch.Tensor, k: int) -> torch.Tensor:
    C = x_nchw.shape[1]
    w = torch.ones(C, 1, k, k, dtype=x_nchw.dtype) / (k * k)
    return F.conv2d(F.pad(x_nchw, (k // 2,) * 4, mode="replicate"), w, groups=C)


def make_pairs(B: int, H: int = 384, W: int = 512, seed: int = 0, max_flow: float = 8.0):
    """-> (source, target, gt_flow): [B,H,W,3] BGR 0..255 fp32 x2, [B,H,W,2] px.

    source = box-smoothed uniform noise; gt = smooth random field; target =
    source sampled at x - gt (so that warping target by gt reconstructs source)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
    src = _box(_box(src, 5), 5)
    src = (src - src.amin()) / (src.amax() - src.amin()) * 255.0
    lo = torch.randn(B, 2, H // 32, W // 32, generator=g)
    gt = F.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False) * max_flow
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    gx = (xs.unsqueeze(0) - gt[:, 0]) / (W - 1) * 2 - 1
    gy = (ys.unsqueeze(0) - gt[:, 1]) / (H - 1) * 2 - 1
    tgt = F.grid_sample(src, torch.stack([gx, gy], dim=-1), mode="bilinear", padding_mode="border", align_corners=True)
    return (src.permute(0, 2, 3, 1).contiguous(), tgt.permute(0, 2, 3, 1).contiguous(),
            gt.permute(0, 2, 3, 1).contiguous())
