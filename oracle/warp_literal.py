"""Second, independent restatement of the warp: a literal per-pixel NumPy loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the Python-unrolled graph
construction of flyingChairsWrapFlow.py:801-838 index for index (``pos_x`` is the
ROW grid, ``pos_y`` the COLUMN grid, :790-795) in float32 scalars.  Used only to
cross-check the vectorised ``oracle.loss_interp.warp`` on small cases.
(The author's own NumPy check, check_loss.py:61-138, clamps the FLAT index
instead (:106-109); the TF graph -- which is what trains -- clamps x and y
separately, and that is what is restated here.)
"""
from __future__ import annotations

import numpy as np


def warp_literal(flows_scaled: np.ndarray, target: np.ndarray) -> np.ndarray:
    flows_scaled = np.asarray(flows_scaled, dtype=np.float32)
    target = np.asarray(target, dtype=np.float32)
    num_batch, height, width, channels = target.shape
    out = np.zeros_like(target)
    one = np.float32(1.0)
    for b in range(num_batch):
        flat = target[b].reshape(height * width, channels)
        for r in range(height):          # pos_x
            for q in range(width):       # pos_y
                u = flows_scaled[b, r, q, 0]
                v = flows_scaled[b, r, q, 1]
                fx = np.floor(u)
                fy = np.floor(v)
                x = int(fx)
                y = int(fy)
                xw = np.float32(u - fx)
                yw = np.float32(v - fy)
                x0 = q + x
                x1 = x0 + 1
                y0 = r + y
                y1 = y0 + 1
                x0 = min(max(x0, 0), width - 1)
                x1 = min(max(x1, 0), width - 1)
                y0 = min(max(y0, 0), height - 1)
                y1 = min(max(y1, 0), height - 1)
                wa = (one - xw) * (one - yw)
                wb = (one - xw) * yw
                wc = xw * (one - yw)
                wd = xw * yw
                for c in range(channels):
                    Ia = flat[y0 * width + x0, c]
                    Ib = flat[y1 * width + x0, c]
                    Ic = flat[y0 * width + x1, c]
                    Id = flat[y1 * width + x1, c]
                    out[b, r, q, c] = Ia * wa + Ib * wb + Ic * wc + Id * wd
    return out
