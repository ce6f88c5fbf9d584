"""Drop-in for the hot-path helpers of ``utils.py``: Middlebury .flo I/O (utils.py:4-21,23-53) and the end-point error (:64-68)."""
from __future__ import annotations

import numpy as np

TAG_FLOAT = 202021.25        # utils.py:12
TAG_CHAR = b"PIEH"


def readFlow(fn):
    """Read a .flo file in Middlebury format (utils.py:4-21) -> float32 [h,w,2], or None when the magic number is wrong."""
    with open(fn, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or TAG_FLOAT != magic[0]:
            print("Magic number incorrect. Invalid .flo file")
            return None
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
        return np.resize(data, (h, w, 2))


def writeFlow(filename, uv, v=None):
    """Write optical flow in Middlebury format (utils.py:23-53)."""
    if v is None:
        assert uv.ndim == 3 and uv.shape[2] == 2
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u = uv
    assert u.shape == v.shape
    height, width = u.shape
    with open(filename, "wb") as f:
        f.write(TAG_CHAR)
        np.array(width).astype(np.int32).tofile(f)
        np.array(height).astype(np.int32).tofile(f)
        tmp = np.zeros((height, width * 2))
        tmp[:, np.arange(width) * 2] = u
        tmp[:, np.arange(width) * 2 + 1] = v
        tmp.astype(np.float32).tofile(f)


def flow_ee(f1, f2, mask=None):
    """Average end-point error (utils.py:64-68).  CUDA tensors run the device reduction (dofb_epe_sum); numpy arrays the numpy formula."""
    import torch
    if isinstance(f1, torch.Tensor) and f1.is_cuda:
        from . import ops
        out = torch.zeros(1, dtype=torch.float64, device=f1.device)
        ops.epe_sum(f1.contiguous(), torch.as_tensor(f2, dtype=torch.float32, device=f1.device).contiguous(), out)
        return float(out.item()) / (f1.numel() // 2)
    f1, f2 = np.asarray(f1), np.asarray(f2)
    ee_tot = np.sqrt((f1[:, :, :, 0] - f2[:, :, :, 0]) ** 2 + (f1[:, :, :, 1] - f2[:, :, :, 1]) ** 2)
    return np.mean(ee_tot, axis=None)


def pad_to_multiple(images, mult: int = 64, mode: str = "replicate"):
    """[B,H,W,C] -> [B,ceil(H/mult)*mult,ceil(W/mult)*mult,C] by padding the bottom / right edge (Sintel's 436 x 1024 -> 448 x 1024: the
    engines need H and W to be multiples of 64).  Returns (padded, (H, W)) so that outputs can be cropped back with ``crop_to``."""
    import torch
    import torch.nn.functional as F
    t = torch.as_tensor(images)
    B, H, W, C = t.shape
    ph, pw = (-H) % mult, (-W) % mult
    if ph == 0 and pw == 0:
        return t, (H, W)
    x = t.permute(0, 3, 1, 2).float()
    x = F.pad(x, (0, pw, 0, ph), mode=mode if mode != "zeros" else "constant")
    return x.permute(0, 2, 3, 1).contiguous(), (H, W)


def crop_to(t, hw, scale: int = 1):
    """Crop a [B,h,w,C] map produced at 1/scale resolution of a padded input back to the un-padded size hw."""
    h, w = -(-hw[0] // scale), -(-hw[1] // scale)
    return t[:, :h, :w, :]
