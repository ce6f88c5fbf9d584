"""VGG16 guided model -- CPU oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no reference fixtures).

Follows ``flyingChairsWrapFlow_vgg.VGG16(photo_source, photo_target, geo_source, geo_target, loss_weight)``
(flyingChairsWrapFlow_vgg.py:7-132): LRN of the geo pair for the loss (:10-11), 13 3x3 ELU convs + 5 2x2 max-pools on
concat(photo pair) (:20-41), decoder at 5 scales with the pool outputs as skips (:72-122), loss variant B (:135-317),
all four images pre-scaled by the trainer (flyingChairsTrain_vgg.py:181-188)."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import tf_ops
from .flownet_s import FLOW_SCALES, HYPER
from .loss_interp import loss_interp

CONVS = [("conv1_1", 6, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128), ("conv3_1", 128, 256),
         ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
         ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)]
REFINE = [(5, 512, "upconv4", 256, "up_pr5to4"), (4, 770, "upconv3", 128, "up_pr4to3"), (3, 386, "upconv2", 64, "up_pr3to2"),
          (2, 194, "upconv1", 32, "up_pr2to1")]
LOSS_WEIGHTS = (16.0, 8.0, 4.0, 2.0, 1.0)      # flyingChairsTrain_vgg.py:171


def param_shapes():
    sh = OrderedDict()
    for name, cin, cout in CONVS:
        sh[name + "/weights"] = (3, 3, cin, cout)
        sh[name + "/biases"] = (cout,)
    for s, cfeat, up, upc, uppr in REFINE:
        sh[f"pr{s}/weights"] = (3, 3, cfeat, 2)
        sh[f"pr{s}/biases"] = (2,)
        sh[up + "/weights"] = (4, 4, upc, cfeat)
        sh[up + "/biases"] = (upc,)
        sh[uppr + "/weights"] = (4, 4, 2, 2)
        sh[uppr + "/biases"] = (2,)
    sh["pr1/weights"] = (3, 3, 98, 2)
    sh["pr1/biases"] = (2,)
    return sh


def init_params(seed=1, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    params = OrderedDict()
    for name, shape in param_shapes().items():
        if name.endswith("biases"):
            params[name] = torch.zeros(shape, dtype=dtype)
        else:
            w = tf_ops.xavier_uniform_(shape, gen, dtype)
            params[name] = tf_ops.bilinear_deconv_weights(shape, dtype) if name.startswith("up") else w   # flyingChairsTrain_vgg.py:141-159
    return params


def max_pool2(x):
    """slim.max_pool2d(x, [2, 2]): stride 2, VALID."""
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def forward(params, photo_source, photo_target, geo_source, geo_target, loss_weight=LOSS_WEIGHTS, *, variant="B", hyper=None):
    hp = dict(HYPER)
    if hyper:
        hp.update(hyper)
    in_norm = tf_ops.lrn(geo_source, depth_radius=4, beta=0.7)       # :10
    out_norm = tf_ops.lrn(geo_target, depth_radius=4, beta=0.7)      # :11
    x = torch.cat([photo_source, photo_target], dim=3)               # :20
    pools = {}
    it = iter(CONVS)
    for lvl, n in enumerate([2, 2, 3, 3, 3], start=1):
        for _ in range(n):
            name, _ci, _co = next(it)
            x = tf_ops.elu(tf_ops.conv2d_same(x, params[name + "/weights"], params[name + "/biases"], 1))
        x = max_pool2(x)
        pools[lvl] = x
    losses, prs = {}, {}
    feat = pools[5]
    for s, _cfeat, upname, _upc, uppr in REFINE:
        pr = tf_ops.conv2d_same(feat, params[f"pr{s}/weights"], params[f"pr{s}/biases"], 1)
        prs[s] = pr
        hs, ws = pr.shape[1], pr.shape[2]
        losses[s], _ = loss_interp(pr, tf_ops.resize_bilinear_legacy(in_norm, hs, ws), tf_ops.resize_bilinear_legacy(out_norm, hs, ws),
                                   hp["epsilon"], hp["alpha_c"], hp["alpha_s"], hp["lambda_smooth"], FLOW_SCALES[s], variant=variant)
        up = tf_ops.elu(tf_ops.conv2d_transpose_same(feat, params[upname + "/weights"], params[upname + "/biases"]))
        up_pr = tf_ops.conv2d_transpose_same(pr, params[uppr + "/weights"], params[uppr + "/biases"])
        feat = torch.cat([pools[s - 1], up, up_pr], dim=3)
    pr1 = tf_ops.conv2d_same(feat, params["pr1/weights"], params["pr1/biases"], 1)
    prs[1] = pr1
    h1, w1 = pr1.shape[1], pr1.shape[2]
    losses[1], recon1 = loss_interp(pr1, tf_ops.resize_bilinear_legacy(in_norm, h1, w1), tf_ops.resize_bilinear_legacy(out_norm, h1, w1),
                                    hp["epsilon"], hp["alpha_c"], hp["alpha_s"], hp["lambda_smooth"], FLOW_SCALES[1], variant=variant)
    lw = [float(v) for v in loss_weight]
    total = sum(lw[i] * losses[i + 1]["total"] for i in range(5))    # :124-125
    return [losses[s] for s in range(1, 6)], [prs[s] * FLOW_SCALES[s] for s in range(1, 6)], recon1, total


def loss_and_grads(params, *imgs, **kw):
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    losses, flows_all, prev1, total = forward(leaf, *imgs, **kw)
    total.backward()
    return total.detach(), OrderedDict((k, v.grad.detach()) for k, v in leaf.items()), losses, flows_all, prev1
