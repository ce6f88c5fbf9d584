"""TensorFlow-0.1x op semantics restated on torch CPU tensors (NHWC in/out).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each helper names the TF op it
restates and the reference call site that relies on it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def same_pad(in_size: int, k: int, stride: int) -> tuple[int, int, int]:
    """TF 'SAME' padding: returns (out_size, pad_before, pad_after).

    out = ceil(in / stride); total = max((out-1)*stride + k - in, 0);
    before = total // 2, after = total - before (the extra pixel goes AFTER).
    Relied on by every slim.conv2d in flyingChairsWrapFlow.py:31-40.
    """
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def conv2d_same(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, stride: int = 1) -> torch.Tensor:
    """slim.conv2d(..., padding='SAME') without activation.

    x: [B,H,W,Cin] NHWC.  w: TF layout [kh,kw,Cin,Cout].  Cross-correlation.
    (flyingChairsWrapFlow.py:31-40,58,69,...)
    """
    kh, kw = w.shape[0], w.shape[1]
    _, pt, pb = same_pad(x.shape[1], kh, stride)
    _, pl, pr = same_pad(x.shape[2], kw, stride)
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_transpose_same(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, stride: int = 2) -> torch.Tensor:
    """slim.conv2d_transpose(k=2*stride, stride) 'SAME': out = stride * in.

    w: TF layout [kh,kw,Cout,Cin].  It is the input-gradient of the SAME conv
    (k, stride) over a stride*in sized map, whose padding is (k-stride)/2 on
    both sides for k = 2*stride (flyingChairsWrapFlow.py:65-66).
    """
    kh = w.shape[0]
    assert kh == 2 * stride and w.shape[1] == kh
    pad = (kh - stride) // 2
    xn = x.permute(0, 3, 1, 2)
    y = F.conv_transpose2d(xn, w.permute(3, 2, 0, 1), b, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def elu(x: torch.Tensor) -> torch.Tensor:
    """tf.nn.elu: x > 0 ? x : exp(x) - 1."""
    return torch.where(x > 0, x, torch.expm1(x))


def lrn(x: torch.Tensor, depth_radius: int = 4, bias: float = 1.0, alpha: float = 1.0, beta: float = 0.7) -> torch.Tensor:
    """tf.nn.local_response_normalization (flyingChairsWrapFlow.py:25-26).

    out[c] = x[c] / (bias + alpha * sum_{|j-c|<=r} x[j]^2) ** beta, TF defaults
    bias=1, alpha=1.  With 3 channels and r=4 the window is all channels.
    """
    C = x.shape[-1]
    sq = x * x
    if depth_radius >= C - 1:
        s = sq.sum(-1, keepdim=True).expand_as(x)
    else:
        pad = F.pad(sq, (depth_radius, depth_radius))
        s = sum(pad[..., j:j + C] for j in range(2 * depth_radius + 1))
    return x / (bias + alpha * s) ** beta


def resize_bilinear_legacy(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """tf.image.resize_bilinear(align_corners=False), TF<=1.x 'legacy' mapping.

    src = dst * (in / out)  (no half-pixel offset); lerp between floor(src) and
    min(floor(src)+1, in-1).  At integer ratios this is exact decimation
    x[:, ::r, ::r] (flyingChairsWrapFlow.py:61-62).
    """
    B, H, W, C = x.shape
    sy = torch.arange(out_h, dtype=torch.float64) * (H / out_h)
    sx = torch.arange(out_w, dtype=torch.float64) * (W / out_w)
    y0 = sy.floor().long()
    x0 = sx.floor().long()
    y1 = torch.clamp(y0 + 1, max=H - 1)
    x1 = torch.clamp(x0 + 1, max=W - 1)
    fy = (sy - y0).to(x.dtype).view(1, out_h, 1, 1)
    fx = (sx - x0).to(x.dtype).view(1, 1, out_w, 1)
    top = x[:, y0][:, :, x0] * (1 - fx) + x[:, y0][:, :, x1] * fx
    bot = x[:, y1][:, :, x0] * (1 - fx) + x[:, y1][:, :, x1] * fx
    return top * (1 - fy) + bot * fy


def tf_constant_fill(values: list[float], shape: tuple[int, ...], dtype=torch.float32) -> torch.Tensor:
    """tf.constant(list, shape=...) with a short list: row-major fill, the rest
    of the tensor takes the LAST list element (flyingChairsWrapFlow.py:48)."""
    n = math.prod(shape)
    flat = list(values) + [values[-1]] * (n - len(values))
    return torch.tensor(flat[:n], dtype=dtype).reshape(shape)


def xavier_uniform_(shape: tuple[int, ...], gen: torch.Generator, dtype=torch.float32) -> torch.Tensor:
    """slim's default weights_initializer (xavier_initializer, uniform=True).

    For a [kh,kw,cin,cout] filter fan_in = kh*kw*cin, fan_out = kh*kw*cout,
    limit = sqrt(6 / (fan_in + fan_out)).  (Explicit in version1/model/Flownet.py:46-51.)
    For a transposed-conv filter [kh,kw,cout,cin] TF computes the fans from the
    same positional rule (dim -2 is 'in', dim -1 is 'out').
    """
    rf = 1
    for d in shape[:-2]:
        rf *= d
    fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)


def bilinear_deconv_weights(shape: tuple[int, int, int, int], dtype=torch.float32) -> torch.Tensor:
    """train.load_deconv_weights (flyingChairsTrain.py:78-92).

    f = ceil(k/2) (a float in py2), c = (2f - 1 - f%2) / (2f); value(x,y) =
    (1-|x/f - c|)(1-|y/f - c|); placed on the channel diagonal weights[:,:,i,i]
    for i < shape[2].  k=4 -> outer([.25,.75,.75,.25]).
    """
    k = shape[0]
    f = float(math.ceil(k / 2.0))
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    ax = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float64)
    bil = torch.outer(ax, ax)
    w = torch.zeros(shape, dtype=torch.float64)
    for i in range(shape[2]):
        if i < shape[3]:
            w[:, :, i, i] = bil
        else:
            # the reference would raise IndexError here; never happens for its own
            # layers where Cout(shape[2]) <= Cin(shape[3]).
            raise IndexError("bilinear diagonal init needs shape[2] <= shape[3]")
    return w.to(dtype)
