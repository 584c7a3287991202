"""CPU oracle of the training step's parameter update (TEST INFRASTRUCTURE ONLY -- nothing on the product path imports this).

The reference's update is ``torch.optim.Adam`` (``utils/optimize.py:6-8``: betas (0.9, 0.999), L2 weight decay, amsgrad off; a
third-party dependency, PyTorch, whose published algorithm is restated here op by op from ``torch/optim/adam.py``
``_single_tensor_adam``) followed by ``EMAHelper.update`` (``models/ddm_wavelet.py:48-53``). Pinned by
``tests/test_oracle_pinning.py`` against ``torch.optim.Adam`` itself and against the reference-style EMA loop on CPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class AdamEmaOracle:
    def __init__(self, params: List[torch.Tensor], lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 mu: Optional[float] = None):
        self.p = [p.detach().clone().float() for p in params]
        self.m = [torch.zeros_like(p) for p in self.p]
        self.v = [torch.zeros_like(p) for p in self.p]
        self.shadow = [p.clone() for p in self.p] if mu is not None else None   # EMAHelper.register: ddm_wavelet.py:40-46
        self.lr, self.betas, self.eps, self.wd, self.mu = lr, betas, eps, weight_decay, mu
        self.t = 0

    @torch.no_grad()
    def step(self, grads: List[Optional[torch.Tensor]]):
        """One optimizer.step() (+ ema_helper.update() when mu was given); a None gradient skips Adam for that tensor."""
        b1, b2 = self.betas
        self.t += 1
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = bc2 ** 0.5
        for i, g in enumerate(grads):
            if g is None:
                continue
            g = g.float()
            if self.wd != 0:
                g = g.add(self.p[i], alpha=self.wd)
            self.m[i].lerp_(g, 1 - b1)
            self.v[i].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[i].sqrt() / bc2_sqrt).add_(self.eps)
            self.p[i].addcdiv_(self.m[i], denom, value=-step_size)
        if self.shadow is not None:
            for i in range(len(self.p)):
                self.shadow[i] = (1. - self.mu) * self.p[i] + self.mu * self.shadow[i]
