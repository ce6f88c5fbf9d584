"""Drop-in for ``version1/model/warpflow.py`` / ``flyingChairsWrapFlow_vgg.loss_interp`` (variant B)."""
from __future__ import annotations

from .flyingChairsWrapFlow import _loss_interp
from ._lib import DeepOFError


def loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """version1/model/warpflow.loss_interp (warpflow.py:4-173) with needMask=True and
    needImageGradients=False, the values version1/model/Flownet.py:71-84 passes."""
    dw = deltaWeights or {}
    if not dw.get("needMask", True):
        raise DeepOFError("loss_interp: needMask=False is not implemented on the CUDA path")
    if dw.get("needImageGradients", False):
        raise DeepOFError("loss_interp: needImageGradients=True (edge-aware smoothness) is not implemented yet (SURVEY.md 8f.4)")
    return _loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, 1)
