"""Drop-in for the hot part of ``flyingChairsTrain_vgg.py`` -- what ``deepOF_fc.py`` launches.

The reference builds four placeholders (photo / geo source and target, already scaled to (x - mean)/255 by the trainer,
flyingChairsTrain_vgg.py:95-112,181-188), ``flyingChairsWrapFlow_vgg.VGG16(...)``, Adam, and then runs
``train_op.run(feed_dict={photo_source, photo_target, geo_source, geo_target, loss_weight, learning_rate})`` (:190-195) and
``sess.run([loss, midFlows, total_loss], ...)`` (:198-203).  ``TrainStep`` keeps that contract on the CUDA engine.
"""
from __future__ import annotations

import numpy as np
import torch

from .flownet import VGG16Flow, VGG_LOSS_WEIGHTS, FLYINGCHAIRS_MEAN
from .flyingChairsTrain import _StepBase, train as _train_base, LEARNING_RATE, learning_rate_at, load_deconv_weights  # noqa: F401

WEIGHT_L = [16, 8, 4, 2, 1]          # flyingChairsTrain_vgg.py:171
BATCH_SIZE = 8                       # :16


class TrainStep(_StepBase):
    """Feeds ``photo_source``, ``photo_target`` (network input) and ``geo_source``, ``geo_target`` (loss images), all pre-scaled;
    ``loss_weight`` [5]; ``learning_rate``."""
    FEEDS = ("photo_source", "photo_target", "geo_source", "geo_target")
    DEFAULT_WEIGHTS = WEIGHT_L
    U8_FEEDS = False            # the trainer pre-scales its feeds on the host (flyingChairsTrain_vgg.py:181-188): float arrays

    def __init__(self, batch_size: int = BATCH_SIZE, image_size=(320, 448), device="cuda", math_mode="fp32", seed: int | None = 1,
                 distributed: bool = False, **kw):
        engine = VGG16Flow(batch_size, image_size[0], image_size[1], device=device, math_mode=math_mode, seed=seed, **kw)
        self._init_io(engine, distributed)

    def _forward(self, dev, lw, with_grad):
        self.engine.forward(dev[0], dev[1], lw, with_grad=with_grad, geo_source=dev[2], geo_target=dev[3], prescaled=True)


class train(_train_base):
    """flyingChairsTrain_vgg.train (:40-233): pre-scaling by the dataset mean (:181-183), then the four feeds.  The reference's host
    augmentations (utils.geoAugmentation / photoAugmentation, :186-187) are supplied by ``augment_fn(source, target) ->
    (photo_source, photo_target, geo_source, geo_target)``; without one the un-augmented pair feeds both roles (the reference's own
    evaluation path, :247-249)."""
    STEP = TrainStep
    WEIGHTS = WEIGHT_L

    def __init__(self, sample_fn=None, image_size=(320, 448), batch_size=BATCH_SIZE, augment_fn=None, **kw):
        self.mean = np.array(FLYINGCHAIRS_MEAN, dtype=np.float32)       # :47
        self.augment_fn = augment_fn
        super().__init__(sample_fn, image_size=image_size, batch_size=batch_size, **kw)

    def _feed(self, batch, lr=None):
        if isinstance(batch[0], torch.Tensor):               # device-decoded batches (flyingChairsLoader): pre-scale where the data is
            mean = torch.as_tensor(self.mean, device=batch[0].device)
            source, target = (batch[0].float() - mean) / 255.0, (batch[1].float() - mean) / 255.0
        else:
            source = (np.asarray(batch[0], dtype=np.float32) - self.mean) / 255.0        # :182-183
            target = (np.asarray(batch[1], dtype=np.float32) - self.mean) / 255.0
        if self.augment_fn is not None:
            ps, pt, gs, gt = self.augment_fn(source, target)
        else:
            ps, pt, gs, gt = source, target, source, target
        d = {"photo_source": ps, "photo_target": pt, "geo_source": gs, "geo_target": gt, "loss_weight": self.WEIGHTS}
        if lr is not None:
            d["learning_rate"] = lr
        return d
