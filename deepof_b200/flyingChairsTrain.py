"""Drop-in for the hot part of ``flyingChairsTrain.py``: the training step.

The reference builds placeholders + model + Adam once (flyingChairsTrain.py:94-124) and then
calls ``train_op.run(feed_dict={source_img, target_img, loss_weight, learning_rate})`` per
iteration (:178) and ``sess.run([loss, midFlows, total_loss], feed_dict)`` for logging (:181).
``TrainStep`` keeps that contract: host numpy (or pinned torch) batches in, numpy scalars out.
"""
from __future__ import annotations

import numpy as np
import torch

from .flownet import FlowNetS, LOSS_WEIGHTS
from . import ddp as _ddp

LEARNING_RATE = 0.000016            # flyingChairsTrain.py:27
LR_DECAY, EPOCHS_PER_DECAY = 0.5, 18   # :29-33
WEIGHT_L = [16, 8, 4, 2, 1, 1]      # :165


class TrainStep:
    """train_op + the fetches of flyingChairsTrain.trainNet."""

    def __init__(self, batch_size: int, image_size=(384, 512), device="cuda", variant="A", math_mode="fp32",
                 seed: int | None = 1, distributed: bool = False, **kw):
        self.engine = FlowNetS(batch_size, image_size[0], image_size[1], device=device, variant=variant, math_mode=math_mode,
                               seed=seed, **kw)
        self.device = self.engine.device
        B, H, W = batch_size, image_size[0], image_size[1]
        # pinned staging: the feed_dict H2D copy of flyingChairsTrain.py:178
        self._pin_src = torch.empty(B, H, W, 3, dtype=torch.float32).pin_memory()
        self._pin_tgt = torch.empty(B, H, W, 3, dtype=torch.float32).pin_memory()
        self._dev_src = torch.empty(B, H, W, 3, dtype=torch.float32, device=self.device)
        self._dev_tgt = torch.empty(B, H, W, 3, dtype=torch.float32, device=self.device)
        self.reducer = _ddp.GradReducer(self.engine) if distributed else None
        if self.reducer is not None:
            self.reducer.broadcast_params()

    def _feed(self, source, target):
        if isinstance(source, torch.Tensor) and source.is_cuda:
            return source, target
        self._pin_src.copy_(torch.as_tensor(source, dtype=torch.float32))
        self._pin_tgt.copy_(torch.as_tensor(target, dtype=torch.float32))
        self._dev_src.copy_(self._pin_src, non_blocking=True)
        self._dev_tgt.copy_(self._pin_tgt, non_blocking=True)
        return self._dev_src, self._dev_tgt

    def run(self, feed_dict: dict):
        """train_op.run(feed_dict={'source_img','target_img','loss_weight','learning_rate'})."""
        src, tgt = self._feed(feed_dict["source_img"], feed_dict["target_img"])
        lw = feed_dict.get("loss_weight", WEIGHT_L)
        lr = float(feed_dict.get("learning_rate", LEARNING_RATE))
        self.engine.train_step(src, tgt, lw, lr, allreduce=self.reducer)

    def fetch(self, feed_dict: dict):
        """sess.run([loss, midFlows, total_loss], feed_dict) -> numpy (flyingChairsTrain.py:181)."""
        src, tgt = self._feed(feed_dict["source_img"], feed_dict["target_img"])
        lw = feed_dict.get("loss_weight", WEIGHT_L)
        self.engine.forward(src, tgt, lw, with_grad=False)
        losses, flows_all, _prev = self.engine.outputs()
        l4 = self.engine.loss4.cpu().numpy()
        keys = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
        losses_np = [{k: l4[s, i] for i, k in enumerate(keys)} for s in range(6)]
        loss_sum = float((l4[:, 0] * np.asarray(lw, dtype=np.float32)).sum())
        return losses_np, [f.cpu().numpy() for f in flows_all], loss_sum

    def last_loss(self) -> float:
        """D2H read of the weighted total of the step that just ran."""
        return float(self.engine.total_loss().item())


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
