"""Scratch check of the cta_group::2 gather-GEMM path against the single-CTA path (run under `timeout`)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepof_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
def run(case, mth):
    B, H, W, ci, co, k, s = case
    g = torch.Generator().manual_seed(sum(case))
    ld = (ci + 63) // 64 * 64
    x = torch.zeros(B, H, W, ld, device="cuda"); x[..., :ci] = torch.randn(B, H, W, ci, generator=g).cuda()
    x16 = x.to(torch.bfloat16)
    w = (torch.randn(k, k, ci, co, generator=g) / math.sqrt(k * k * ci)).cuda()
    b = (torch.randn(co, generator=g) * 0.1).cuda()
    geom = ops.conv_geom(B, H, W, ci, co, k, s)
    yl = (co + 63) // 64 * 64
    outs = []
    for pairs in (0, 1):
        lib.dofb_enable_cta_pairs(pairs)
        y = torch.zeros(B, geom.oh, geom.ow, yl, device="cuda")
        ops.conv_fwd(geom, ops.Slab(x, 0, ci, x16), w, b, ops.Slab(y, 0, co), ops.ACT_ELU, mth)
        torch.cuda.synchronize()
        outs.append(y)
    lib.dofb_enable_cta_pairs(0)
    d = (outs[0] - outs[1]).abs().max().item()
    print(case, "math", mth, "max|single - pairs| =", d, "max|y| =", outs[0].abs().max().item(), flush=True)
    return d
bad = 0
for case in [(32, 48, 64, 256, 256, 3, 1), (32, 24, 32, 512, 512, 3, 1), (32, 48, 64, 256, 512, 3, 2), (31, 12, 16, 512, 512, 3, 1), (32, 6, 8, 1024, 1024, 3, 1)]:
    for mth in (ops.MATH_BF16, ops.MATH_TF32):
        bad += run(case, mth) > 1e-5
print("FAIL" if bad else "OK")
