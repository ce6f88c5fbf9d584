"""FlowNetC guided model -- CPU oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED, DOUBLY SO: the reference contains NO FlowNetC / correlation layer at all (`grep -i corr`
finds nothing), although BASELINE.json configs 3-4 name one.  The architecture below is therefore a SPECIFICATION, taken from
the FlowNet paper (Dosovitskiy et al., ICCV'15) and completed deepOF-style (ELU activations, the FlowNetS refinement of
flyingChairsWrapFlow.py:58-113 with its warp/loss at 6 scales):

  conv1 7x7/2 3->64, conv2 5x5/2 64->128, conv3 5x5/2 128->256   on source and target with SHARED weights
  corr[b,y,x,(i,j)] = (1/256) * sum_c conv3a[b,y,x,c] * conv3b[b,y+dy_i,x+dx_j,c],  dy,dx in {-20,-18,...,20}, zero outside; ELU
  conv_redir 1x1 256->32 (ELU) on conv3a;  concat [conv_redir(32), corr(441)] -> conv3_1 3x3 473->256
  conv4_1/2, conv5_1/2, conv6_1/2 and the refinement as in FlowNetS (skips conv5_2, conv4_2, conv3_1, conv2a, conv1a)
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import tf_ops
from .flownet_s import REFINE, FLOW_SCALES, HYPER, LOSS_WEIGHTS, FLYINGCHAIRS_MEAN, preprocess, param_shapes as _s_shapes
from .loss_interp import loss_interp

MAX_DISP, STRIDE2 = 20, 2
FRONT = [("conv1", 7, 2, 3, 64), ("conv2", 5, 2, 64, 128), ("conv3", 5, 2, 128, 256)]
TOP = [("conv4_1", 3, 2, 256, 512), ("conv4_2", 3, 1, 512, 512), ("conv5_1", 3, 2, 512, 512), ("conv5_2", 3, 1, 512, 512),
       ("conv6_1", 3, 2, 512, 1024), ("conv6_2", 3, 1, 1024, 1024)]


def correlation(f1: torch.Tensor, f2: torch.Tensor, max_disp: int = MAX_DISP, stride2: int = STRIDE2) -> torch.Tensor:
    """[B,h,w,C] x2 -> [B,h,w,D*D], channel = dy_index*D + dx_index, normalised by C."""
    B, h, w, c = f1.shape
    D = 2 * (max_disp // stride2) + 1
    f2p = F.pad(f2, (0, 0, max_disp, max_disp, max_disp, max_disp))
    outs = []
    for i in range(D):
        for j in range(D):
            dy, dx = -max_disp + i * stride2, -max_disp + j * stride2
            sh = f2p[:, max_disp + dy:max_disp + dy + h, max_disp + dx:max_disp + dx + w]
            outs.append((f1 * sh).sum(-1) / c)
    return torch.stack(outs, dim=-1)


def param_shapes() -> "OrderedDict[str, tuple]":
    sh: "OrderedDict[str, tuple]" = OrderedDict()
    for name, k, cin, cout in [(n, k, ci, co) for n, k, _s, ci, co in FRONT] + [("conv_redir", 1, 256, 32), ("conv3_1", 3, 473, 256)] + \
            [(n, k, ci, co) for n, k, _s, ci, co in TOP]:
        sh[name + "/weights"] = (k, k, cin, cout)
        sh[name + "/biases"] = (cout,)
    for key, shape in _s_shapes().items():
        if key.startswith(("pr", "up")):
            sh[key] = shape
    return sh


def init_params(seed: int = 1, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    params = OrderedDict()
    for name, shape in param_shapes().items():
        if name.endswith("biases"):
            params[name] = torch.zeros(shape, dtype=dtype)
        elif name.startswith("up"):
            tf_ops.xavier_uniform_(shape, gen, dtype)            # keep the RNG stream aligned with the product initialiser
            params[name] = tf_ops.bilinear_deconv_weights(shape, dtype)
        else:
            params[name] = tf_ops.xavier_uniform_(shape, gen, dtype)
    return params


def forward(params, inputs, outputs, loss_weight=LOSS_WEIGHTS, *, variant="A", mean=FLYINGCHAIRS_MEAN, hyper=None):
    hp = dict(HYPER)
    if hyper:
        hp.update(hyper)
    x_in, in_norm = preprocess(inputs, mean)
    x_out, out_norm = preprocess(outputs, mean)

    def conv(name, x, stride):
        return tf_ops.elu(tf_ops.conv2d_same(x, params[name + "/weights"], params[name + "/biases"], stride))

    fa, fb = x_in, x_out
    skips = {}
    for name, _k, stride, _ci, _co in FRONT:
        fa, fb = conv(name, fa, stride), conv(name, fb, stride)
        skips[name] = fa
    corr = tf_ops.elu(correlation(fa, fb))
    redir = conv("conv_redir", fa, 1)
    x = conv("conv3_1", torch.cat([redir, corr], dim=3), 1)
    skips["conv3_1"] = x
    for name, _k, stride, _ci, _co in TOP:
        x = conv(name, x, stride)
        skips[name] = x
    skip_of = {6: "conv5_2", 5: "conv4_2", 4: "conv3_1", 3: "conv2", 2: "conv1"}
    losses, prs = {}, {}
    feat = skips["conv6_2"]
    for s, _cfeat, upname, _upc, uppr, _skip in REFINE:
        pr = tf_ops.conv2d_same(feat, params[f"pr{s}/weights"], params[f"pr{s}/biases"], 1)
        prs[s] = pr
        hs, ws = pr.shape[1], pr.shape[2]
        losses[s], _ = loss_interp(pr, tf_ops.resize_bilinear_legacy(in_norm, hs, ws), tf_ops.resize_bilinear_legacy(out_norm, hs, ws),
                                   hp["epsilon"], hp["alpha_c"], hp["alpha_s"], hp["lambda_smooth"], FLOW_SCALES[s], variant=variant)
        up = tf_ops.elu(tf_ops.conv2d_transpose_same(feat, params[upname + "/weights"], params[upname + "/biases"]))
        up_pr = tf_ops.conv2d_transpose_same(pr, params[uppr + "/weights"], params[uppr + "/biases"])
        feat = torch.cat([skips[skip_of[s]], up, up_pr], dim=3)
    pr1 = tf_ops.conv2d_same(feat, params["pr1/weights"], params["pr1/biases"], 1)
    prs[1] = pr1
    h1, w1 = pr1.shape[1], pr1.shape[2]
    losses[1], recon1 = loss_interp(pr1, tf_ops.resize_bilinear_legacy(in_norm, h1, w1), tf_ops.resize_bilinear_legacy(out_norm, h1, w1),
                                    hp["epsilon"], hp["alpha_c"], hp["alpha_s"], hp["lambda_smooth"], FLOW_SCALES[1], variant=variant)
    lw = [float(v) for v in loss_weight]
    total = sum(lw[i] * losses[i + 1]["total"] for i in range(6))
    return [losses[s] for s in range(1, 7)], [prs[s] * FLOW_SCALES[s] for s in range(1, 7)], recon1, total


def loss_and_grads(params, inputs, outputs, loss_weight=LOSS_WEIGHTS, **kw):
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    losses, flows_all, prev1, total = forward(leaf, inputs, outputs, loss_weight, **kw)
    total.backward()
    return total.detach(), OrderedDict((k, v.grad.detach()) for k, v in leaf.items()), losses, flows_all, prev1
