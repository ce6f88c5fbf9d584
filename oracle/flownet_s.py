"""FlowNetS guided model -- CPU oracle (torch, NHWC, TF weight layouts).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no reference
fixtures exist).  Follows ``flyingChairsWrapFlow.flowNet``
(flyingChairsWrapFlow.py:5-129; clean twin version1/model/Flownet.py:22-166).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import tf_ops
from .loss_interp import loss_interp

# flyingChairsWrapFlow.py:16 (BGR)
FLYINGCHAIRS_MEAN = (97.533268117955444, 99.238235788550085, 97.055973199626948)
# sintelWrapFlow.py:773
SINTEL_MEAN = (70.1433, 83.1915, 92.8827)

# (name, k, stride, cin, cout) -- flyingChairsWrapFlow.py:31-40
TOWER = [
    ("conv1", 7, 2, 6, 64),
    ("conv2", 5, 2, 64, 128),
    ("conv3_1", 5, 2, 128, 256),
    ("conv3_2", 3, 1, 256, 256),
    ("conv4_1", 3, 2, 256, 512),
    ("conv4_2", 3, 1, 512, 512),
    ("conv5_1", 3, 2, 512, 512),
    ("conv5_2", 3, 1, 512, 512),
    ("conv6_1", 3, 2, 512, 1024),
    ("conv6_2", 3, 1, 1024, 1024),
]
# per refinement stage s=6..2: skip layer, channels of feat_s, upconv out channels (:58-111)
REFINE = [
    # (s, feat_channels_at_s, upconv_name, upconv_cout, up_pr_name, skip_name)
    (6, 1024, "upconv5", 512, "up_pr6to5", "conv5_2"),
    (5, 1026, "upconv4", 256, "up_pr5to4", "conv4_2"),
    (4, 770, "upconv3", 128, "up_pr4to3", "conv3_2"),
    (3, 386, "upconv2", 64, "up_pr3to2", "conv2"),
    (2, 194, "upconv1", 32, "up_pr2to1", "conv1"),
]
FEAT_CHANNELS = {6: 1024, 5: 1026, 4: 770, 3: 386, 2: 194, 1: 98}
FLOW_SCALES = {1: 10.0, 2: 5.0, 3: 2.5, 4: 1.25, 5: 0.625, 6: 0.3125}   # :118,107,96,85,74,63
HYPER = dict(epsilon=1e-4, alpha_c=0.25, alpha_s=0.37, lambda_smooth=1.0)  # :43-46
LOSS_WEIGHTS = (16.0, 8.0, 4.0, 2.0, 1.0, 1.0)                             # flyingChairsTrain.py:165


def param_shapes() -> "OrderedDict[str, tuple]":
    """All 52 trainable tensors in creation order, TF layouts.

    conv: weights [k,k,cin,cout]; transposed conv: [k,k,cout,cin]."""
    shapes: "OrderedDict[str, tuple]" = OrderedDict()
    for name, k, _s, cin, cout in TOWER:
        shapes[name + "/weights"] = (k, k, cin, cout)
        shapes[name + "/biases"] = (cout,)
    for s, cfeat, upname, upc, uppr, _skip in REFINE:
        shapes[f"pr{s}/weights"] = (3, 3, cfeat, 2)
        shapes[f"pr{s}/biases"] = (2,)
        shapes[upname + "/weights"] = (4, 4, upc, cfeat)
        shapes[upname + "/biases"] = (upc,)
        shapes[uppr + "/weights"] = (4, 4, 2, 2)
        shapes[uppr + "/biases"] = (2,)
    shapes["pr1/weights"] = (3, 3, 98, 2)
    shapes["pr1/biases"] = (2,)
    return shapes


def init_params(seed: int = 1, dtype=torch.float32, bilinear_deconv: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """slim defaults (xavier-uniform weights, zero biases) then the trainer's
    bilinear overwrite of every variable whose name starts with 'up'
    (flyingChairsTrain.py:135,150-154 -> upconv* AND up_pr*)."""
    gen = torch.Generator().manual_seed(seed)
    params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_shapes().items():
        if name.endswith("biases"):
            params[name] = torch.zeros(shape, dtype=dtype)
        else:
            params[name] = tf_ops.xavier_uniform_(shape, gen, dtype)
            if bilinear_deconv and name.startswith("up"):
                params[name] = tf_ops.bilinear_deconv_weights(shape, dtype)
    return params


def num_params() -> int:
    n = 0
    for shape in param_shapes().values():
        c = 1
        for d in shape:
            c *= d
        n += c
    return n


def preprocess(img: torch.Tensor, mean=FLYINGCHAIRS_MEAN):
    """:16-26 -- (img - mean)/255 and its LRN."""
    m = torch.tensor(mean, dtype=img.dtype).view(1, 1, 1, 3)
    x = (img - m) / 255.0
    return x, tf_ops.lrn(x, depth_radius=4, beta=0.7)


def forward(params, inputs: torch.Tensor, outputs: torch.Tensor, loss_weight=LOSS_WEIGHTS, *, variant: str = "A",
            mean=FLYINGCHAIRS_MEAN, hyper=None, return_feats: bool = False):
    """flowNet(inputs, outputs, loss_weight) -> (losses, flows_all, prev1[, total, feats]).

    inputs/outputs: [B,H,W,3] BGR 0..255 float.  H, W multiples of 64."""
    hp = dict(HYPER)
    if hyper:
        hp.update(hyper)
    x_in, in_norm = preprocess(inputs, mean)
    x_out, out_norm = preprocess(outputs, mean)
    feats: dict[str, torch.Tensor] = {}
    x = torch.cat([x_in, x_out], dim=3)                                   # :31
    for name, _k, stride, _cin, _cout in TOWER:
        x = tf_ops.elu(tf_ops.conv2d_same(x, params[name + "/weights"], params[name + "/biases"], stride))
        feats[name] = x

    losses: dict[int, dict] = {}
    prs: dict[int, torch.Tensor] = {}
    recon1 = None
    feat = feats["conv6_2"]
    for s, _cfeat, upname, _upc, uppr, skip in REFINE:
        pr = tf_ops.conv2d_same(feat, params[f"pr{s}/weights"], params[f"pr{s}/biases"], 1)   # :58
        prs[s] = pr
        hs, ws = pr.shape[1], pr.shape[2]
        src = tf_ops.resize_bilinear_legacy(in_norm, hs, ws)              # :61
        tgt = tf_ops.resize_bilinear_legacy(out_norm, hs, ws)             # :62
        losses[s], _ = loss_interp(pr, src, tgt, hp["epsilon"], hp["alpha_c"], hp["alpha_s"], hp["lambda_smooth"],
                                   FLOW_SCALES[s], variant=variant)      # :64
        up = tf_ops.elu(tf_ops.conv2d_transpose_same(feat, params[upname + "/weights"], params[upname + "/biases"]))
        up_pr = tf_ops.conv2d_transpose_same(pr, params[uppr + "/weights"], params[uppr + "/biases"])
        feat = torch.cat([feats[skip], up, up_pr], dim=3)                 # :67
        feats[f"concat{s - 1}"] = feat
    pr1 = tf_ops.conv2d_same(feat, params["pr1/weights"], params["pr1/biases"], 1)            # :113
    prs[1] = pr1
    h1, w1 = pr1.shape[1], pr1.shape[2]
    src = tf_ops.resize_bilinear_legacy(in_norm, h1, w1)
    tgt = tf_ops.resize_bilinear_legacy(out_norm, h1, w1)
    losses[1], recon1 = loss_interp(pr1, src, tgt, hp["epsilon"], hp["alpha_c"], hp["alpha_s"], hp["lambda_smooth"],
                                    FLOW_SCALES[1], variant=variant)     # :119

    lw = [float(v) for v in loss_weight]
    total = sum(lw[i] * losses[i + 1]["total"] for i in range(6))         # :122-123
    losses_list = [losses[s] for s in range(1, 7)]                        # :126
    flows_all = [prs[s] * FLOW_SCALES[s] for s in range(1, 7)]            # :127
    if return_feats:
        feats.update({f"pr{s}": prs[s] for s in prs})
        return losses_list, flows_all, recon1, total, feats
    return losses_list, flows_all, recon1, total


def loss_and_grads(params, inputs, outputs, loss_weight=LOSS_WEIGHTS, **kw):
    """total loss + d(total)/d(param) by autograd (== TF's minimize() backward)."""
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    losses, flows_all, prev1, total = forward(leaf, inputs, outputs, loss_weight, **kw)
    total.backward()
    grads = OrderedDict((k, v.grad.detach()) for k, v in leaf.items())
    return total.detach(), grads, losses, flows_all, prev1
