"""One bf16 FlowNetS (or, with MODEL=c, FlowNetC) step (B=32, 384x512) for ncu captures of individual kernels (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepof_b200.flownet import FlowNetS, FlowNetC
from deepof_b200.synth import make_pairs
B, H, W = 32, 384, 512
s, t, _ = make_pairs(B, H, W, seed=1)
e = (FlowNetC if os.environ.get("MODEL", "s") == "c" else FlowNetS)(B, H, W, math_mode=sys.argv[1] if len(sys.argv) > 1 else "bf16", tc_wgrad=True)
s, t = s.cuda(), t.cuda()
for i in range(2):
    e.train_step(s, t, lr=1.6e-5)
torch.cuda.synchronize()
print("loss", float(e.total_loss()))
