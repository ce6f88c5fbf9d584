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
    """Bucketed sum-all-reduce of the flat gradient arena, overlapped with the backward pass.

    The backward completes gradients in exactly the reverse of the arena order (decoder from the finest scale up, then the tower
    top-down), so the arena is cut into buckets from its END; the engine reports the lowest finished offset after every layer
    (``ready(offset)``) and each bucket's all-reduce is issued the moment it is complete.  NCCL runs on its own stream: issuing it
    makes that stream wait for the gradients already enqueued, and the remaining backward kernels overlap the transfer over
    NVLink 5 / NVSwitch.  ``finish()`` joins before Adam and returns the 1/world scale the optimiser folds in.

    Measured on 8 B200s (FlowNetS, 32 pairs per GPU, 6.2 ms single-GPU step; scripts/ddp_sweep.sh): the 155 MB all-reduce alone costs
    0.6 ms when issued after the backward (6.81 ms); buckets of 16 / 32 / 64 MB give 6.86 / 6.82 / 6.67 ms -- the persistent GEMMs own
    every SM, so an NCCL kernel only gets in between launches and fewer, larger buckets disturb them least; capping NCCL's CTAs
    (NCCL_MAX_CTAS = 8 / 4) is slower (6.97 / 8.40 ms).  Default: 64 MB buckets, 4 MB tail."""

    def __init__(self, engine, bucket_mb: float | None = None, tail_mb: float | None = None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before building a GradReducer")
        import os
        bucket_mb = float(os.environ.get("DOFB_DDP_BUCKET_MB", "64")) if bucket_mb is None else bucket_mb      # (tuning knobs of the bench scripts)
        tail_mb = float(os.environ.get("DOFB_DDP_TAIL_MB", "4")) if tail_mb is None else tail_mb
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.bounds = self.plan_buckets(list(engine.arena.offsets.values()), engine.grad.numel(), bucket_mb, tail_mb)
        self._next = len(self.bounds) - 1
        self._works = []

    @staticmethod
    def plan_buckets(tensor_offsets, numel: int, bucket_mb: float = 64.0, tail_mb: float = 4.0):
        """Ascending (start, end) element ranges, cut at TENSOR boundaries so that a bucket can be issued the moment the layer that completes
        it has run.  The backward finishes the arena back to front, so the LAST bucket to become ready is the one at offset 0 (conv1 ...):
        nothing can overlap its all-reduce, hence it is kept small (<= tail_mb); the others are >= bucket_mb so that NVLink sees few, large
        messages."""
        starts = sorted(set(int(o) for o in tensor_offsets))
        assert starts and starts[0] == 0
        per, tail = max(int(bucket_mb * (1 << 20) // 4), 1), max(int(tail_mb * (1 << 20) // 4), 1)
        # first bucket (offset 0): as many leading tensors as fit into tail_mb (at least one)
        cuts = [0]
        first_end = numel
        for o in starts[1:]:
            if o > tail:
                break
            first_end = o
        if first_end >= numel:
            return [(0, numel)]
        cuts.append(first_end)
        # the rest, walking forward: close a bucket at the first tensor boundary past bucket_mb
        for o in starts:
            if o > cuts[-1] and o - cuts[-1] >= per:
                cuts.append(o)
        if cuts[-1] != numel:
            if numel - cuts[-1] < per // 4 and len(cuts) > 2:      # a small remainder joins its neighbour
                cuts[-1] = numel
            else:
                cuts.append(numel)
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:])]

    def broadcast_params(self, src: int = 0):
        dist.broadcast(self.engine.theta, src=src, group=self.group)
        from . import ops
        ops.invalidate_weight_cache()       # the tensor-core kernels keep packed copies of the weights: theta just changed under them

    def begin(self):
        self._next = len(self.bounds) - 1
        self._works = []

    def ready(self, floor_offset: int):
        """All gradients at arena offsets >= floor_offset are final: launch every bucket that lies entirely above it."""
        if self.world == 1:
            return
        g = self.engine.grad
        while self._next >= 0 and self.bounds[self._next][0] >= floor_offset:
            a, b = self.bounds[self._next]
            self._works.append(dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._next -= 1

    def finish(self) -> float:
        self.ready(0)
        for w in self._works:
            w.wait()
        self._works = []
        return 1.0 / self.world

    def __call__(self, flat_grad: torch.Tensor) -> float:
        """Non-overlapped form (whole arena at once); returns the scale (1/world) the optimiser must apply."""
        self.begin()
        return self.finish()


def shard_batch(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Even split of the global batch; (start, local_batch)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} must divide evenly over {world} ranks "
                         "(equal local batches keep the loss normaliser exact)")
    lb = global_batch // world
    return rank * lb, lb
