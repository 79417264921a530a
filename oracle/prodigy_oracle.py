"""TEST INFRASTRUCTURE - CPU restatement of the reference's Prodigy optimizer step (ldm/prodigy.py:97-256), fp32 torch.

Only tests/ may import this file, and only as the checker.  Pinned against the UNMODIFIED reference class
(`ldm.prodigy.Prodigy`, importable stand-alone) run in the build container: oracle/make_golden_prodigy.py ->
tests/golden/prodigy.pt.
"""
from __future__ import annotations

import math
from typing import List

import torch


class ProdigyOracle:
    """Single parameter group.  State layout and update order follow ldm/prodigy.py:108-256."""

    def __init__(self, params: List[torch.Tensor], lr=1.0, betas=(0.9, 0.999), beta3=None, eps=1e-8, weight_decay=0.0,
                 decouple=True, use_bias_correction=False, safeguard_warmup=False, d0=1e-6, d_coef=1.0,
                 growth_rate=float("inf")):
        self.params = params
        self.lr, self.betas, self.beta3, self.eps = lr, betas, beta3, eps
        self.weight_decay, self.decouple = weight_decay, decouple
        self.use_bias_correction, self.safeguard_warmup = use_bias_correction, safeguard_warmup
        self.d = self.d0 = self.d_max = d0
        self.d_coef, self.growth_rate = d_coef, growth_rate
        self.d_numerator, self.k = 0.0, 0
        self.state = None

    def step(self, grads: List[torch.Tensor]):
        beta1, beta2 = self.betas
        beta3 = math.sqrt(beta2) if self.beta3 is None else self.beta3                       # :113-115
        d, lr, k = self.d, self.lr, self.k
        bc = ((1 - beta2 ** (k + 1)) ** 0.5) / (1 - beta1 ** (k + 1)) if self.use_bias_correction else 1   # :123-126
        dlr = d * lr * bc                                                                     # :128
        d_numerator = self.d_numerator * beta3                                                # :134-135
        d_denom = 0.0
        if self.state is None:                                                                # :163-170
            self.state = [{"s": torch.zeros_like(p), "p0": p.clone(), "exp_avg": torch.zeros_like(p),
                           "exp_avg_sq": torch.zeros_like(p)} for p in self.params]
        for p, g, st in zip(self.params, grads, self.state):
            if self.weight_decay != 0 and not self.decouple:                                  # :157-158
                g = g + self.weight_decay * p
            d_numerator += (d / self.d0) * dlr * torch.dot(g.flatten(), (st["p0"] - p).flatten()).item()   # :179
            st["exp_avg"].mul_(beta1).add_(g, alpha=d * (1 - beta1))                          # :182
            st["exp_avg_sq"].mul_(beta2).addcmul_(g, g, value=d * d * (1 - beta2))            # :183
            a = (d / self.d0) * d if self.safeguard_warmup else (d / self.d0) * dlr           # :185-188
            st["s"].mul_(beta3).add_(g, alpha=a)
            d_denom += st["s"].abs().sum().item()                                             # :189
        if d_denom == 0:                                                                      # :197-198
            return
        d_hat = self.d_coef * d_numerator / d_denom                                           # :212
        if d == self.d0:                                                                      # :213-214
            d = max(d, d_hat)
        self.d_max = max(self.d_max, d_hat)                                                   # :215
        d = min(self.d_max, d * self.growth_rate)                                             # :216
        self.d_numerator, self.d = d_numerator, d
        for p, st in zip(self.params, self.state):
            denom = st["exp_avg_sq"].sqrt().add_(d * self.eps)                                # :240
            if self.weight_decay != 0 and self.decouple:                                      # :243-244
                p.add_(p, alpha=-self.weight_decay * dlr)
            p.addcdiv_(st["exp_avg"], denom, value=-dlr)                                      # :248
        self.k = k + 1                                                                        # :250
