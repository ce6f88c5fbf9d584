"""Drop-in for ``version1/model/warpflow.py`` / ``flyingChairsWrapFlow_vgg.loss_interp`` (variant B)."""
from __future__ import annotations

from .flyingChairsWrapFlow import _loss_interp
from ._lib import DeepOFError
from . import ops


def loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, deltaWeights=None):
    """version1/model/warpflow.loss_interp (warpflow.py:4-173) with needMask=True (the value version1/model/Flownet.py:71-73 passes).

    ``deltaWeights["needImageGradients"]`` (default False, Flownet.py:84) switches the edge-aware smoothness on (:91-116,148-157): the
    element-wise smoothness losses are weighted by 1 - |Sobel(gray(inputs))| / max; the Sobel filters themselves are fixed constants
    of the reference (Flownet.py:87-92) and are baked into the kernel."""
    dw = deltaWeights or {}
    if not dw.get("needMask", True):
        raise DeepOFError("loss_interp: needMask=False is not implemented on the CUDA path")
    edge_w = ops.edge_weights(inputs.contiguous()) if dw.get("needImageGradients", False) else None
    return _loss_interp(flows, inputs, outputs, epsilon, alpha_c, alpha_s, lambda_smooth, flow_scale, 1, edge_w=edge_w)
