"""Drop-in for ``version1/model/Flownet.py``: ``model(source_imgs, target_imgs, sample_mean, loss_weight, hyper_params, is_training)``
(version1/model/Flownet.py:22-166) -- FlowNetS with the clean loss (variant B), the dataset mean as an argument and the hyper-parameters
as a length-4 sequence in the order ``[lambda_smooth, epsilon, alpha_c, alpha_s]`` (:65-68).

The v1 graph augments inside the model when ``is_training`` (:40-41); augmentation is outside the hot path here, so the caller passes
already-augmented pairs: ``geo=(geo_source, geo_target)`` supplies the loss images when they differ from the network input
(:43-44 LRN-normalises the GEOMETRICALLY augmented copy)."""
from __future__ import annotations

import torch

from .flownet import FlowNetS
from ._lib import DeepOFError

LOSS_WEIGHT_V1 = [9, 7, 5, 3, 3, 1]          # version1/trainOF.py:82
_engines: dict = {}


def hyper_from_list(hyper_params) -> dict:
    """[lambda_smooth, epsilon, alpha_c, alpha_s] (Flownet.py:65-68) -> engine keyword form.  (sintelTrain.py:181 feeds its own order
    [epsilon, alpha_c, alpha_s, lambda_smooth]: use ``hyper_from_sintel_list`` for that one.)"""
    lam, eps, ac, as_ = [float(v) for v in hyper_params]
    return dict(lambda_smooth=lam, epsilon=eps, alpha_c=ac, alpha_s=as_)


def hyper_from_sintel_list(hyper_param_list) -> dict:
    eps, ac, as_, lam = [float(v) for v in hyper_param_list]
    return dict(lambda_smooth=lam, epsilon=eps, alpha_c=ac, alpha_s=as_)


def get_engine(batch, height, width, sample_mean, hyper_params, device="cuda", math_mode="fp32", **kw) -> FlowNetS:
    hp = hyper_from_list(hyper_params)
    key = (batch, height, width, str(device), math_mode, tuple(float(m) for m in sample_mean), tuple(sorted(hp.items())))
    if key not in _engines:
        _engines[key] = FlowNetS(batch, height, width, device=device, variant="B", math_mode=math_mode, mean=sample_mean, hyper=hp, **kw)
    return _engines[key]


def model(source_imgs, target_imgs, sample_mean, loss_weight, hyper_params, is_training=False, engine: FlowNetS | None = None):
    """-> (losses[6] of dict{total, Charbonnier_reconstruct, U_loss, V_loss}, flows_all[6], prev1)  (Flownet.py:160-166).

    source_imgs / target_imgs: [B,H,W,3] BGR 0..255 float32 CUDA tensors; sample_mean: 3 floats; loss_weight: 6 floats."""
    if is_training:
        raise DeepOFError("Flownet.model(is_training=True) augments inside the TF graph (Flownet.py:40-41); feed augmented pairs with is_training=False")
    B, H, W, _ = source_imgs.shape
    mean = sample_mean.tolist() if isinstance(sample_mean, torch.Tensor) else [float(m) for m in sample_mean]
    hp = hyper_params.tolist() if isinstance(hyper_params, torch.Tensor) else list(hyper_params)
    eng = engine or get_engine(B, H, W, mean, hp, device=source_imgs.device)
    lw = loss_weight.tolist() if isinstance(loss_weight, torch.Tensor) else list(loss_weight)
    eng.forward(source_imgs.contiguous(), target_imgs.contiguous(), lw, with_grad=False)
    return eng.outputs()
