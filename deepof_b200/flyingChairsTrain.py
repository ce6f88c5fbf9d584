"""Drop-in for the hot part of ``flyingChairsTrain.py``: the training step.

The reference builds placeholders + model + Adam once (flyingChairsTrain.py:94-124) and then
calls ``train_op.run(feed_dict={source_img, target_img, loss_weight, learning_rate})`` per
iteration (:178) and ``sess.run([loss, midFlows, total_loss], feed_dict)`` for logging (:181).
``TrainStep`` keeps that contract: host numpy (or pinned torch) batches in, numpy scalars out.
"""
from __future__ import annotations

import numpy as np
import torch

from .flownet import FlowNetS, FlowNetC, LOSS_WEIGHTS
from . import ddp as _ddp

LEARNING_RATE = 0.000016            # flyingChairsTrain.py:27
LR_DECAY, EPOCHS_PER_DECAY = 0.5, 18   # :29-33
WEIGHT_L = [16, 8, 4, 2, 1, 1]      # :165


class TrainStep:
    """train_op + the fetches of flyingChairsTrain.trainNet.

    Host<->device traffic is pipelined like a training loop would do it by hand: two device input buffers, a copy stream,
    and a lagged loss read-back -- the H2D copy of step i+1 and the D2H of step i's loss overlap the kernels of step i."""

    def __init__(self, batch_size: int, image_size=(384, 512), device="cuda", variant="A", math_mode="fp32",
                 seed: int | None = 1, distributed: bool = False, model: str = "flownets", **kw):
        cls = {"flownets": FlowNetS, "flownetc": FlowNetC}[model.lower()]
        self.engine = cls(batch_size, image_size[0], image_size[1], device=device, variant=variant, math_mode=math_mode,
                          seed=seed, **kw)
        self.device = self.engine.device
        B, H, W = batch_size, image_size[0], image_size[1]
        shape = (B, H, W, 3)
        # pinned staging for pageable (numpy) feeds: the feed_dict H2D copy of flyingChairsTrain.py:178
        self._pin = [[torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
        self._dev = [[torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(2)] for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._copied = [torch.cuda.Event() for _ in range(2)]       # H2D of slot i finished
        self._consumed = [torch.cuda.Event() for _ in range(2)]     # the step reading slot i has consumed its inputs
        self._loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._loss_ready = [torch.cuda.Event() for _ in range(2)]
        self._n = 0
        self.reducer = _ddp.GradReducer(self.engine) if distributed else None
        if self.reducer is not None:
            self.reducer.broadcast_params()

    def _feed(self, source, target):
        """Stage one batch; returns device tensors valid for the current stream."""
        if isinstance(source, torch.Tensor) and source.is_cuda:
            return source, target
        slot = self._n & 1
        host = []
        for j, arr in enumerate((source, target)):
            t = torch.as_tensor(arr, dtype=torch.float32)
            if not t.is_pinned():                     # pageable feed (numpy): one host memcpy into pinned staging
                self._copied[slot].synchronize()      # the previous H2D out of this staging buffer is done
                self._pin[slot][j].copy_(t)
                t = self._pin[slot][j]
            host.append(t)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed[slot])      # step n-2 no longer reads this device buffer
            self._dev[slot][0].copy_(host[0], non_blocking=True)
            self._dev[slot][1].copy_(host[1], non_blocking=True)
            self._copied[slot].record(self._copy_stream)
        cur.wait_event(self._copied[slot])
        return self._dev[slot][0], self._dev[slot][1]

    def run(self, feed_dict: dict):
        """train_op.run(feed_dict={'source_img','target_img','loss_weight','learning_rate'}) -- asynchronous."""
        slot = self._n & 1
        src, tgt = self._feed(feed_dict["source_img"], feed_dict["target_img"])
        lw = feed_dict.get("loss_weight", WEIGHT_L)
        lr = float(feed_dict.get("learning_rate", LEARNING_RATE))
        self.engine.forward(src, tgt, lw, with_grad=True)
        self._consumed[slot].record()                 # inputs are only read by the pre-processing kernel
        self.engine.backward(reducer=self.reducer)                       # bucketed all-reduce overlapped with the backward
        scale = self.reducer.finish() if self.reducer is not None else 1.0
        self.engine.adam_step(lr, grad_scale=scale)
        # lagged loss read-back: D2H into pinned memory, consumed one step later (or by last_loss(sync=True))
        self._loss_pin[slot:slot + 1].copy_(self.engine.total_loss().reshape(1), non_blocking=True)
        self._loss_ready[slot].record()
        self._n += 1

    def fetch(self, feed_dict: dict):
        """sess.run([loss, midFlows, total_loss], feed_dict) -> numpy (flyingChairsTrain.py:181)."""
        src, tgt = self._feed(feed_dict["source_img"], feed_dict["target_img"])
        lw = feed_dict.get("loss_weight", WEIGHT_L)
        self.engine.forward(src, tgt, lw, with_grad=False)
        self._consumed[self._n & 1].record()
        self._n += 1
        losses, flows_all, _prev = self.engine.outputs()
        l4 = self.engine.loss4.cpu().numpy()
        keys = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
        losses_np = [{k: l4[s, i] for i, k in enumerate(keys)} for s in range(6)]
        loss_sum = float((l4[:, 0] * np.asarray(lw, dtype=np.float32)).sum())
        return losses_np, [f.cpu().numpy() for f in flows_all], loss_sum

    def last_loss(self, lag: int = 0) -> float:
        """Weighted total loss of the most recent run() (lag=0, waits for it) or of the one before (lag=1: its D2H has
        already landed, so the host never stalls the device -- what a pipelined training loop logs)."""
        k = self._n - 1 - lag
        if k < 0:
            return float("nan")
        self._loss_ready[k & 1].synchronize()
        return float(self._loss_pin[k & 1])


class train:
    """Shape of the reference's ``train`` class (flyingChairsTrain.py:40-213) around TrainStep.

    The dataset loader / augmentation / checkpoint / evaluation code of the reference is out of
    scope; ``sample_fn(batch_size, iteration) -> (source, target)`` supplies batches."""

    def __init__(self, sample_fn, image_size=(384, 512), batch_size=4, max_iters=10, lr=LEARNING_RATE, **kw):
        self.image_size = image_size
        self.numLosses = 6
        self.batch_size = batch_size
        self.step = TrainStep(batch_size, image_size, **kw)
        self.sample_fn = sample_fn
        self.lr = lr
        self.max_iters = max_iters

    def trainNet(self, display: int = 0):
        for iteration in range(1, self.max_iters + 1):
            source, target = self.sample_fn(self.batch_size, iteration)
            self.step.run({"source_img": source, "target_img": target, "loss_weight": WEIGHT_L, "learning_rate": self.lr})
            if display and iteration % display == 0:
                _l, _f, loss_sum = self.step.fetch({"source_img": source, "target_img": target, "loss_weight": WEIGHT_L})
                assert not np.isnan(loss_sum), "Model diverged with loss = NaN"      # flyingChairsTrain.py:203
                print("---Train Batch(%d): Iter %04d: Loss_sum %4.4f" % (self.batch_size, iteration, loss_sum))
