"""Host mirror of ldm/prodigy.py:Prodigy (the reference trainer's optimizer) over ONE flat fp32 bucket.

Same constructor arguments and update rule (ldm/prodigy.py:54-256, single parameter group).  The parameters are
re-bound as views of a flat bucket, the state (s, p0, exp_avg, exp_avg_sq) is four flat buffers, and a step is two
fused kernels (af_prodigy_moments / af_prodigy_apply) around the scalar d estimate - the gradient bucket is the one
the all-reduce produced (train_cond.allreduce_gradients), so no per-tensor launches remain.  Parameters whose .grad is
None at the first step are left out, like the reference's `continue` (:149-150); the set must not change afterwards.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional

import torch

from . import _lib


BUCKET_ALIGN = 64   # elements: every parameter starts on a 256-byte boundary of the flat buckets (kernels read the
#                     parameter views with 16-byte vector loads / TMA)


def flat_layout(params):
    """-> (offsets, total) of the flat fp32 bucket layout shared by the parameter, state and gradient buckets."""
    offs, o = [], 0
    for p in params:
        offs.append(o)
        o += (p.numel() + BUCKET_ALIGN - 1) // BUCKET_ALIGN * BUCKET_ALIGN
    return offs, o


class Prodigy:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1.0, betas=(0.9, 0.999), beta3=None, eps=1e-8,
                 weight_decay=0, decouple=True, use_bias_correction=False, safeguard_warmup=False, d0=1e-6,
                 d_coef=1.0, growth_rate=float("inf"), fsdp_in_use=False):
        if not 0.0 < d0:
            raise ValueError(f"Invalid d0 value: {d0}")
        if not 0.0 < lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 < eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if fsdp_in_use:
            raise NotImplementedError("FSDP sharding of the optimizer state")
        self.all_params: List[torch.nn.Parameter] = list(params)
        self.param_groups = [dict(lr=lr, betas=betas, beta3=beta3, eps=eps, weight_decay=weight_decay, d=d0, d0=d0,
                                  d_max=d0, d_numerator=0.0, d_denom=0.0, d_hat=d0, d_coef=d_coef, k=0,
                                  growth_rate=growth_rate, use_bias_correction=use_bias_correction, decouple=decouple,
                                  safeguard_warmup=safeguard_warmup, params=self.all_params)]
        self.d0 = d0
        self.params: Optional[List[torch.nn.Parameter]] = None
        self.flat_p = self.p0 = self.s = self.exp_avg = self.exp_avg_sq = None
        self._sums = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.all_params:
            p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)

    def _init_state(self):
        self.params = [p for p in self.all_params if p.grad is not None]
        if not self.params:
            return False
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("Prodigy: parameters must be fp32 CUDA tensors (no CPU fallback)")
        offs, total = flat_layout(self.params)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
        for p, o in zip(self.params, offs):         # parameters become views of the bucket (aligned starts, zero gaps)
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view_as(p)
        self.p0 = self.flat_p.clone()
        self.s = torch.zeros_like(self.flat_p)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self._sums = torch.zeros(2, dtype=torch.float64, device=self.flat_p.device)
        return True

    @torch.no_grad()
    def step(self, flat_grad: Optional[torch.Tensor] = None):
        """flat_grad: the fp32 gradient bucket in the flat_layout() of the parameters (train_cond.GradBucket.flat), or
        None to gather it from p.grad."""
        if self.params is None and not self._init_state():
            return None
        if any(p.grad is None for p in self.params):
            raise RuntimeError("Prodigy: the set of parameters with gradients changed after the first step")
        if flat_grad is None:
            offs, total = flat_layout(self.params)
            flat_grad = torch.zeros(total, dtype=torch.float32, device=self.flat_p.device)
            for p, o in zip(self.params, offs):
                flat_grad[o:o + p.numel()].copy_(p.grad.reshape(-1))
        n = self.flat_p.numel()
        if flat_grad.numel() != n or flat_grad.dtype != torch.float32 or not flat_grad.is_contiguous():
            raise ValueError("Prodigy.step: gradient bucket does not match the parameter bucket")
        lib = _lib.load()
        g = self.param_groups[0]
        beta1, beta2 = g["betas"]
        beta3 = math.sqrt(beta2) if g["beta3"] is None else g["beta3"]                       # :113-115
        k, d, lr, d0 = g["k"], g["d"], g["lr"], g["d0"]
        bc = ((1 - beta2 ** (k + 1)) ** 0.5) / (1 - beta1 ** (k + 1)) if g["use_bias_correction"] else 1
        dlr = d * lr * bc                                                                     # :128
        decay, decouple = g["weight_decay"], g["decouple"]
        s_alpha = (d / d0) * d if g["safeguard_warmup"] else (d / d0) * dlr                   # :185-188
        stream = torch.cuda.current_stream().cuda_stream
        self._sums.zero_()
        rc = lib.af_prodigy_moments(self.flat_p.data_ptr(), flat_grad.data_ptr(), self.p0.data_ptr(), self.s.data_ptr(),
                                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n, beta1, beta2, beta3, d,
                                    s_alpha, decay if (decay != 0 and not decouple) else 0.0, self._sums.data_ptr(), stream)
        _lib.check(rc, "af_prodigy_moments")
        dot, d_denom = (float(v) for v in self._sums.tolist())                                # .item() syncs, like :179,:189
        d_numerator = g["d_numerator"] * beta3 + (d / d0) * dlr * dot                         # :135,:179
        if d_denom == 0:                                                                      # :197-198
            return None
        d_hat = g["d_coef"] * d_numerator / d_denom                                           # :212
        if d == d0:
            d = max(d, d_hat)
        d_max = max(g["d_max"], d_hat)
        d = min(d_max, d * g["growth_rate"])                                                  # :216
        g.update(d_numerator=d_numerator, d_denom=d_denom, d=d, d_max=d_max, d_hat=d_hat)
        rc = lib.af_prodigy_apply(self.flat_p.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n, dlr,
                                  d * g["eps"], decay if (decay != 0 and decouple) else 0.0, stream)
        _lib.check(rc, "af_prodigy_apply")
        g["k"] = k + 1                                                                        # :250
        return None
