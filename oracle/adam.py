"""TF-form Adam -- CPU oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

``tf.train.AdamOptimizer(learning_rate)`` as used at flyingChairsTrain.py:124
(beta1=0.9, beta2=0.999, epsilon=1e-8, TF "epsilon-hat" formulation):

    t      <- t + 1
    lr_t   <- lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    m      <- beta1*m + (1-beta1)*g
    v      <- beta2*v + (1-beta2)*g*g
    theta  <- theta - lr_t * m / (sqrt(v) + epsilon)

No weight decay, no clipping (add_regularization_losses=False, :121).
"""
from __future__ import annotations

import math

import torch


class TFAdam:
    def __init__(self, params, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.params = params
        self.beta1, self.beta2, self.eps = beta1, beta2, epsilon
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def lr_t(self, lr: float) -> float:
        return lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)

    @torch.no_grad()
    def step(self, grads, lr: float):
        self.t += 1
        lr_t = self.lr_t(lr)
        for k, p in self.params.items():
            g = grads[k]
            self.m[k].mul_(self.beta1).add_(g, alpha=1 - self.beta1)
            self.v[k].mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
            p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))
