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


def learning_rate_at(epoch: int, base_lr: float = LEARNING_RATE, decay: float = LR_DECAY, epochs_per_decay: int = EPOCHS_PER_DECAY) -> float:
    """The reference's schedule (flyingChairsTrain.py:27-33,208-209): ``lr *= 0.5`` after every 18th epoch; ``epoch`` counts from 1."""
    return base_lr * decay ** ((epoch - 1) // epochs_per_decay)


def load_deconv_weights(var: torch.Tensor, sess=None) -> None:
    """train.load_deconv_weights(var, sess) (flyingChairsTrain.py:78-92): overwrite a transposed-conv weight [k,k,co,ci] with the
    diagonal bilinear-upsampling kernel (k = 4 -> outer([.25,.75,.75,.25])), in place.  ``sess`` is accepted for signature
    compatibility; after changing weights of a live engine call ``engine.load_params`` / ``ops.invalidate_weight_cache``."""
    from .flownet import bilinear_deconv
    from . import ops
    var.copy_(bilinear_deconv(tuple(var.shape)).to(var.device))
    ops.invalidate_weight_cache()


class _StepBase:
    """train_op + fetches around one engine: host<->device traffic pipelined like a training loop would do it by hand -- two sets of device
    input buffers, a copy stream, and a lagged loss read-back: the H2D copy of step i+1 and the D2H of step i's loss overlap the kernels
    of step i.  Subclasses name the image feeds (``FEEDS``) and say how the engine consumes them."""
    FEEDS: tuple = ()
    DEFAULT_WEIGHTS: list = []
    U8_FEEDS = True             # raw 0..255 images: uint8 feeds are shipped as bytes (the pre-scaled feeds of the VGG trainer are float)

    def _init_io(self, engine, distributed):
        self.engine = engine
        self.device = engine.device
        shape = (engine.B, engine.H, engine.W, 3)
        n = len(self.FEEDS)
        # pinned staging for pageable (numpy) feeds: the feed_dict H2D copy of flyingChairsTrain.py:178
        self._pin = [[torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(n)] for _ in range(2)]
        self._dev = [[torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(n)] for _ in range(2)]
        # uint8 feeds (what the reference's loader returns: cv2.imread / cv2.resize arrays, flyingChairsLoader.py:64-80; TF casts them to the
        # float32 placeholders on the host): shipped as bytes -- a quarter of the H2D traffic -- and cast by the pre-processing kernel
        self._pin8 = self._dev8 = None
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._copied = [torch.cuda.Event() for _ in range(2)]       # H2D of slot i finished
        self._consumed = [torch.cuda.Event() for _ in range(2)]     # the step reading slot i has consumed its inputs
        self._loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._loss_ready = [torch.cuda.Event() for _ in range(2)]
        self._n_feed = 0            # staged batches (run + fetch): picks the input slot
        self._n_run = 0             # training steps: picks the loss slot
        self.reducer = _ddp.GradReducer(self.engine) if distributed else None
        if self.reducer is not None:
            self.reducer.broadcast_params()

    def _feed(self, feed_dict):
        """Stage one batch; returns (slot, device tensors valid for the current stream)."""
        arrays = [feed_dict[k] for k in self.FEEDS]
        slot = self._n_feed & 1
        self._n_feed += 1
        if all(isinstance(a, torch.Tensor) and a.is_cuda for a in arrays):
            return None, arrays
        as_u8 = self.U8_FEEDS and all((isinstance(a, np.ndarray) and a.dtype == np.uint8) or (isinstance(a, torch.Tensor) and a.dtype == torch.uint8) for a in arrays)
        if as_u8 and self._pin8 is None:
            shape = tuple(self._pin[0][0].shape)
            self._pin8 = [[torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in arrays] for _ in range(2)]
            self._dev8 = [[torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in arrays] for _ in range(2)]
        pin, devb = (self._pin8, self._dev8) if as_u8 else (self._pin, self._dev)
        host = []
        for j, arr in enumerate(arrays):
            t = torch.as_tensor(arr) if as_u8 else torch.as_tensor(arr, dtype=torch.float32)
            if tuple(t.shape) != tuple(self._pin[slot][j].shape):
                raise ValueError(f"feed '{self.FEEDS[j]}': expected shape {tuple(self._pin[slot][j].shape)}, got {tuple(t.shape)} "
                                 "(static shapes, like the reference's placeholders)")
            if not t.is_pinned():                     # pageable feed (numpy): one host memcpy into pinned staging
                self._copied[slot].synchronize()      # the previous H2D out of this staging buffer is done
                pin[slot][j].copy_(t)
                t = pin[slot][j]
            host.append(t)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed[slot])      # the step that last read this device buffer is past its inputs
            for j, t in enumerate(host):
                devb[slot][j].copy_(t, non_blocking=True)
            self._copied[slot].record(self._copy_stream)
        cur.wait_event(self._copied[slot])
        return slot, devb[slot]

    # ---- engine adapters (overridden for the 4-feed guided model) ----
    def _forward(self, dev, lw, with_grad):
        self.engine.forward(dev[0], dev[1], lw, with_grad=with_grad)

    def run(self, feed_dict: dict):
        """train_op.run(feed_dict={<image feeds>, 'loss_weight', 'learning_rate'}) -- asynchronous (flyingChairsTrain.py:178)."""
        slot, dev = self._feed(feed_dict)
        lw = feed_dict.get("loss_weight", self.DEFAULT_WEIGHTS)
        lr = float(feed_dict.get("learning_rate", LEARNING_RATE))
        self._forward(dev, lw, True)
        if slot is not None:
            self._consumed[slot].record()             # inputs are only read by the pre-processing kernel
        self.engine.backward(reducer=self.reducer)                       # bucketed all-reduce overlapped with the backward
        scale = self.reducer.finish() if self.reducer is not None else 1.0
        self.engine.adam_step(lr, grad_scale=scale)
        # lagged loss read-back: D2H into pinned memory, consumed one step later (or by last_loss(lag=0))
        ls = self._n_run & 1
        self._loss_pin[ls:ls + 1].copy_(self.engine.total_loss().reshape(1), non_blocking=True)
        self._loss_ready[ls].record()
        self._n_run += 1

    def fetch(self, feed_dict: dict):
        """sess.run([loss, midFlows, total_loss], feed_dict) -> numpy (flyingChairsTrain.py:181)."""
        slot, dev = self._feed(feed_dict)
        lw = feed_dict.get("loss_weight", self.DEFAULT_WEIGHTS)
        self._forward(dev, lw, False)
        if slot is not None:
            self._consumed[slot].record()
        n = self.engine.N_SCALES
        _losses, flows_all, _prev = self.engine.outputs()
        l4 = self.engine.loss4.cpu().numpy()
        keys = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")
        losses_np = [{k: l4[s, i] for i, k in enumerate(keys)} for s in range(n)]
        loss_sum = float((l4[:, 0] * np.asarray(lw, dtype=np.float32)[:n]).sum())
        return losses_np, [f.cpu().numpy() for f in flows_all], loss_sum

    def last_loss(self, lag: int = 0) -> float:
        """Weighted total loss of the most recent run() (lag=0, waits for it) or of the one before (lag=1: its D2H has
        already landed, so the host never stalls the device -- what a pipelined training loop logs)."""
        k = self._n_run - 1 - lag
        if k < 0:
            return float("nan")
        self._loss_ready[k & 1].synchronize()
        return float(self._loss_pin[k & 1])


class TrainStep(_StepBase):
    """train_op + the fetches of flyingChairsTrain.trainNet: feeds ``source_img``, ``target_img`` (raw BGR 0..255), ``loss_weight`` [6],
    ``learning_rate`` (flyingChairsTrain.py:98-124,178-181)."""
    FEEDS = ("source_img", "target_img")
    DEFAULT_WEIGHTS = WEIGHT_L

    def __init__(self, batch_size: int, image_size=(384, 512), device="cuda", variant="A", math_mode="fp32",
                 seed: int | None = 1, distributed: bool = False, model: str = "flownets", **kw):
        cls = {"flownets": FlowNetS, "flownetc": FlowNetC}[model.lower()]
        engine = cls(batch_size, image_size[0], image_size[1], device=device, variant=variant, math_mode=math_mode, seed=seed, **kw)
        self._init_io(engine, distributed)


class train:
    """Shape of the reference's ``train`` class (flyingChairsTrain.py:40-213) around TrainStep.

    Batches come from ``sample_fn(batch_size, iteration) -> (source, target)`` or, with ``data_path``, from the FlyingChairs reader
    (deepof_b200.flyingChairsLoader).  The learning rate halves every 18 epochs (:27-33,208-209)."""
    STEP = TrainStep
    WEIGHTS = WEIGHT_L

    def __init__(self, sample_fn=None, image_size=(384, 512), batch_size=4, max_iters=10, lr=LEARNING_RATE, data_path=None,
                 iters_per_epoch=None, max_epochs=1, **kw):
        self.image_size = image_size
        self.numLosses = len(self.WEIGHTS)
        self.batch_size = batch_size
        if sample_fn is None:
            if data_path is None:
                raise ValueError("train: give sample_fn(batch_size, iteration) or data_path (a FlyingChairs_release directory)")
            from .flyingChairsLoader import flyingChairsLoader
            self.flyingChairs = flyingChairsLoader(data_path, image_size)
            sample_fn = lambda bs, it: self.flyingChairs.sampleTrain(bs, it)[:2]      # noqa: E731
            iters_per_epoch = iters_per_epoch or max(1, len(self.flyingChairs.trainList) // batch_size)
        self.step = self.STEP(batch_size, image_size, **kw)
        self.sample_fn = sample_fn
        self.lr = lr
        self.max_iters = max_iters
        self.iters_per_epoch = iters_per_epoch or max_iters
        self.max_epochs = max_epochs

    def load_deconv_weights(self, var, sess=None):
        """train.load_deconv_weights(var, sess) (flyingChairsTrain.py:78-92)."""
        load_deconv_weights(var, sess)

    def _feed(self, batch, lr=None):
        d = {"source_img": batch[0], "target_img": batch[1], "loss_weight": self.WEIGHTS}
        if lr is not None:
            d["learning_rate"] = lr
        return d

    def trainNet(self, display: int = 0):
        done = 0
        for epoch in range(1, self.max_epochs + 1):
            lr = learning_rate_at(epoch, self.lr)
            for iteration in range(1, self.iters_per_epoch + 1):
                if done >= self.max_iters:
                    return
                batch = self.sample_fn(self.batch_size, iteration)
                self.step.run(self._feed(batch, lr))
                done += 1
                if display and iteration % display == 0:
                    _l, _f, loss_sum = self.step.fetch(self._feed(batch))
                    assert not np.isnan(loss_sum), "Model diverged with loss = NaN"      # flyingChairsTrain.py:203
                    print("---Train Batch(%d): Epoch %03d Iter %04d: Loss_sum %4.4f" % (self.batch_size, epoch, iteration, loss_sum))
