"""Drop-in for ``flyingChairsWrapFlow_vgg`` (what ``deepOF_fc.py`` -> ``flyingChairsTrain_vgg.py`` builds)."""
from __future__ import annotations

import torch

from .flownet import VGG16Flow, VGG_LOSS_WEIGHTS  # noqa: F401
from .warpflow import loss_interp  # noqa: F401  (flyingChairsWrapFlow_vgg.py:135 is the variant-B loss)

_engines: dict = {}


def get_engine(batch, height, width, device="cuda", math_mode="fp32", **kw) -> VGG16Flow:
    key = (batch, height, width, str(device), math_mode)
    if key not in _engines:
        _engines[key] = VGG16Flow(batch, height, width, device=device, math_mode=math_mode, **kw)
    return _engines[key]


def VGG16(photo_source, photo_target, geo_source, geo_target, loss_weight, engine: VGG16Flow | None = None):
    """VGG16(photo_source, photo_target, geo_source, geo_target, loss_weight) -> (losses[5], flows_all[5], prev1)
    (flyingChairsWrapFlow_vgg.py:7-132).  Inputs: [B,H,W,3] float32 CUDA tensors already scaled to (x - mean)/255."""
    B, H, W, _ = photo_source.shape
    eng = engine or get_engine(B, H, W, device=photo_source.device)
    lw = loss_weight.tolist() if isinstance(loss_weight, torch.Tensor) else list(loss_weight)
    eng.forward(photo_source.contiguous(), photo_target.contiguous(), lw, with_grad=False,
                geo_source=geo_source.contiguous(), geo_target=geo_target.contiguous())
    return eng.outputs()
