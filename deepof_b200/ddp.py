"""Batch-sharded data parallelism: one process per GPU, ONE exchange step per iteration.

The reference is single-GPU; image pairs are independent, so the path shards by
batch and the only exchange is the gradient all-reduce (sum) over NVLink/NVSwitch via NCCL,
followed by a 1/world scale folded into the Adam kernel.  Because the loss normaliser contains the
local batch (flyingChairsWrapFlow.py:848) and every rank has the same local batch, mean-of-local
gradients averaged over ranks equals the global-batch gradient.

The flat gradient arena is reduced in a few large buckets issued in backward order; NVSwitch makes
cost insensitive to bucket count, so buckets are sized for overlap, not link count.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, engine, bucket_mb: float = 64.0, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before building a GradReducer")
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        n = engine.grad.numel()
        per = max(int(bucket_mb * (1 << 20) // 4), 1)
        self.bounds = [(i, min(i + per, n)) for i in range(0, n, per)]

    def broadcast_params(self, src: int = 0):
        dist.broadcast(self.engine.theta, src=src, group=self.group)

    def __call__(self, flat_grad: torch.Tensor) -> float:
        """Sum-all-reduce the arena; returns the scale (1/world) the optimiser must apply."""
        if self.world == 1:
            return 1.0
        works = [dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                 for a, b in reversed(self.bounds)]
        for w in works:
            w.wait()
        return 1.0 / self.world


def shard_batch(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Even split of the global batch; (start, local_batch)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} must divide evenly over {world} ranks "
                         "(equal local batches keep the loss normaliser exact)")
    lb = global_batch // world
    return rank * lb, lb
