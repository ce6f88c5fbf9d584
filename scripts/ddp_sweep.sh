#!/bin/bash
# Bucket-plan / NCCL-CTA sweep of the data-parallel step on N GPUs of one box (run under gpurun --gpus N).
N=${1:-8}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-accuracy --no-extras > gpurun_out/ddp_$tag.json 2> gpurun_out/ddp_$tag.err
  python -c "
import json; raw=open('gpurun_out/ddp_$tag.json').read(); d=json.loads(raw[raw.index('{'):]); print('$tag', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" || tail -3 gpurun_out/ddp_$tag.err; }
run default X=1
run single DOFB_DDP_TAIL_MB=100000
run b64 DOFB_DDP_BUCKET_MB=64
run b16 DOFB_DDP_BUCKET_MB=16 DOFB_DDP_TAIL_MB=2
run cta8 NCCL_MAX_CTAS=8
run cta4_single NCCL_MAX_CTAS=4 DOFB_DDP_TAIL_MB=100000
