"""Stream-ordering check of the overlapped gradient all-reduce (run under torchrun, world >= 2): every rank feeds the SAME batch, so the
all-reduced gradient must equal world x the local gradient of a non-distributed backward on the same engine (forward is bit-stable with
DOFB_SPLITK=0; the weight gradients differ only in fp32 atomic order).  A missing stream dependency shows up as O(1) per-tensor errors."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DOFB_SPLITK"] = "0"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from deepof_b200.flownet import FlowNetS
from deepof_b200.ddp import GradReducer
from deepof_b200.synth import make_pairs
B, H, W = 8, 384, 512
src, tgt, _ = make_pairs(B, H, W, seed=77)
src, tgt = src.cuda(), tgt.cuda()
eng = FlowNetS(B, H, W, device=f"cuda:{local}", math_mode="bf16", seed=1, tc_wgrad=True)
red = GradReducer(eng, bucket_mb=float(os.environ.get("BUCKET_MB", "8")), tail_mb=1.0)
red.broadcast_params()
worst = 0.0
for it in range(3):
    eng.forward(src, tgt, with_grad=True)
    eng.backward(reducer=None)
    torch.cuda.synchronize()
    g_local = {k: v.clone() for k, v in eng.grads.items()}
    eng.forward(src, tgt, with_grad=True)
    eng.backward(reducer=red)
    scale = red.finish()
    torch.cuda.synchronize()
    for k, v in eng.grads.items():
        ref = g_local[k].double() * world
        err = float((v.double() - ref).abs().max() / (ref.abs().max() + 1e-30))
        worst = max(worst, err)
        if err > 1e-3:
            print(f"rank {rank} iter {it}: {k} rel err {err:.3e}", flush=True)
t = torch.tensor([worst], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"ddp_check: world={world} side_stream={eng._side is not None} side_wgrad={eng._side_wgrad} buckets={len(red.bounds)} worst rel err {float(t):.3e}", "OK" if float(t) < 1e-3 else "FAIL")
dist.destroy_process_group()
