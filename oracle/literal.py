"""Second, independent restatements (literal NumPy loops) of the dense ops the oracle otherwise takes from torch:

* TF 'SAME' convolution (slim.conv2d, flyingChairsWrapFlow.py:31-40) -- cross-correlation, asymmetric padding;
* conv2d_transpose 'SAME' (slim.conv2d_transpose, :65-66) -- written as the scatter it is;
* the FlowNetC correlation (no reference symbol; FlowNet paper definition).

TEST INFRASTRUCTURE (see oracle/__init__.py): used only to cross-check oracle/tf_ops.py and oracle/flownet_c.py on small cases."""
from __future__ import annotations

import numpy as np


def conv2d_same_literal(x: np.ndarray, w: np.ndarray, b, stride: int) -> np.ndarray:
    """x [B,H,W,Ci], w [kh,kw,Ci,Co] -> [B,ceil(H/s),ceil(W/s),Co]; pad_before = floor(total/2) (TF)."""
    B, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    oh, ow = -(-H // stride), -(-W // stride)
    pt = max((oh - 1) * stride + kh - H, 0) // 2
    pl = max((ow - 1) * stride + kw - W, 0) // 2
    y = np.zeros((B, oh, ow, Co), dtype=np.float64)
    for oy in range(oh):
        for ox in range(ow):
            for i in range(kh):
                iy = oy * stride + i - pt
                if iy < 0 or iy >= H:
                    continue
                for j in range(kw):
                    ix = ox * stride + j - pl
                    if ix < 0 or ix >= W:
                        continue
                    y[:, oy, ox, :] += x[:, iy, ix, :].astype(np.float64) @ w[i, j].astype(np.float64)
    if b is not None:
        y += np.asarray(b, dtype=np.float64)
    return y


def conv2d_transpose_same_literal(x: np.ndarray, w: np.ndarray, b, stride: int) -> np.ndarray:
    """x [B,h,w,Ci], w [kh,kw,Co,Ci] (TF transposed-conv layout) -> [B,s*h,s*w,Co]: every input pixel SCATTERS its kh x kw patch to
    out[s*y + i - pt, s*x + j - pl] where (pt, pl) is the SAME padding of the forward conv this op is the gradient of."""
    B, h, wd, Ci = x.shape
    kh, kw, Co, _ = w.shape
    H, W = h * stride, wd * stride
    pt = max((h - 1) * stride + kh - H, 0) // 2
    pl = max((wd - 1) * stride + kw - W, 0) // 2
    y = np.zeros((B, H, W, Co), dtype=np.float64)
    for yy in range(h):
        for xx in range(wd):
            for i in range(kh):
                oy = yy * stride + i - pt
                if oy < 0 or oy >= H:
                    continue
                for j in range(kw):
                    ox = xx * stride + j - pl
                    if ox < 0 or ox >= W:
                        continue
                    y[:, oy, ox, :] += x[:, yy, xx, :].astype(np.float64) @ w[i, j].astype(np.float64).T
    if b is not None:
        y += np.asarray(b, dtype=np.float64)
    return y


def correlation_literal(f1: np.ndarray, f2: np.ndarray, max_disp: int, stride2: int) -> np.ndarray:
    """out[b,y,x,(dy_i, dx_i)] = (1/C) <f1[b,y,x,:], f2[b,y+dy,x+dx,:]>, dy, dx in {-md, -md+s2, ..., md}, zero outside the map."""
    B, h, w, C = f1.shape
    D = 2 * (max_disp // stride2) + 1
    out = np.zeros((B, h, w, D * D), dtype=np.float64)
    for y in range(h):
        for x in range(w):
            for a in range(D):
                yy = y - max_disp + a * stride2
                if yy < 0 or yy >= h:
                    continue
                for c in range(D):
                    xx = x - max_disp + c * stride2
                    if xx < 0 or xx >= w:
                        continue
                    out[:, y, x, a * D + c] = (f1[:, y, x, :].astype(np.float64) * f2[:, yy, xx, :].astype(np.float64)).sum(axis=1) / C
    return out
