"""Scratch check of the cta_group::2 weight-gradient path against the single-CTA path (run under `timeout`)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepof_b200 import ops, _lib
lib = _lib.load()
def run(case, mth):
    B, H, W, ci, co, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    xl, yl = (ci + 63) // 64 * 64, (co + 63) // 64 * 64
    x = torch.zeros(B, H, W, xl, device="cuda"); x[..., :ci] = torch.randn(B, H, W, ci, generator=g).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    dy = torch.zeros(B, geom.oh, geom.ow, yl, device="cuda"); dy[..., :co] = torch.randn(B, geom.oh, geom.ow, co, generator=g).cuda()
    outs = []
    for pairs in (0, 1):
        lib.dofb_enable_cta_pairs(pairs)
        dw = torch.zeros(k, k, ci, co, device="cuda")
        ops.conv_wgrad(geom, ops.Slab(x, 0, ci, x.to(torch.bfloat16)), ops.Slab(dy, 0, co, dy.to(torch.bfloat16)), dw, None, mth)
        torch.cuda.synchronize()
        outs.append(dw)
    lib.dofb_enable_cta_pairs(0)
    d = (outs[0] - outs[1]).abs().max().item() / outs[0].abs().max().item()
    print(case, "math", mth, "rel max|single - pairs| =", d, flush=True)
    return d
bad = 0
for case in [(8, 48, 64, 256, 256, 3, 1), (8, 24, 32, 512, 512, 3, 1), (8, 48, 64, 128, 256, 5, 2), (8, 96, 128, 64, 128, 5, 2), (4, 12, 16, 512, 512, 3, 2),
             (8, 96, 128, 32, 194, 4, 2), (3, 6, 8, 1024, 1024, 3, 1), (8, 48, 64, 256, 130, 3, 1)]:
    for mth in (ops.MATH_BF16, ops.MATH_TF32):
        bad += run(case, mth) > 1e-4          # (fp32 atomics: the order of the split-K partial sums differs)
print("FAIL" if bad else "OK")
