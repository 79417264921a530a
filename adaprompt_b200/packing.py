"""One-time repacking of reference-layout (state_dict) weights into the kernel layouts.

All functions are pure tensor reshuffles (torch as plumbing); they run once at load time.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

GEGLU_TILE = 128  # output columns per 256-wide packed GEMM tile


def pack_conv3x3(w_oihw: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d weight [Cout, Cin, 3, 3] -> bf16 [Cout, 3, 3, Cin] (K order = tap-major, channel-minor)."""
    return w_oihw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def pack_conv1x1(w_oi11: torch.Tensor) -> torch.Tensor:
    return w_oi11.reshape(w_oi11.shape[0], w_oi11.shape[1]).contiguous().to(torch.bfloat16)


def pack_geglu(w: torch.Tensor, b: Optional[torch.Tensor]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """GEGLU.proj (ldm/modules/attention.py:35-39): rows [value (inner) | gate (inner)] -> per 256-row tile
    [128 value rows | 128 gate rows] so that a GEMM tile holds both halves of the same output columns."""
    two_inner = w.shape[0]
    inner = two_inner // 2
    assert inner % GEGLU_TILE == 0, inner
    t = inner // GEGLU_TILE
    wv = w[:inner].reshape(t, GEGLU_TILE, -1)
    wg = w[inner:].reshape(t, GEGLU_TILE, -1)
    wp = torch.stack([wv, wg], dim=1).reshape(two_inner, -1).contiguous()
    bp = None
    if b is not None:
        bv = b[:inner].reshape(t, GEGLU_TILE)
        bg = b[inner:].reshape(t, GEGLU_TILE)
        bp = torch.stack([bv, bg], dim=1).reshape(two_inner).contiguous()
    return wp, bp


def head_pad(d: int) -> int:
    """Column stride of one head in the Q / K buffers (d = 40 is padded to 48 with zero weight rows)."""
    return 48 if d == 40 else d


def pack_qk(wq: torch.Tensor, wk: Optional[torch.Tensor], heads: int, fold_scale: bool = True) -> torch.Tensor:
    """to_q (and optionally to_k) -> one bf16 weight [heads*dp (+ heads*dp), C] whose GEMM output is the
    head-padded Q | K buffer the attention kernel reads.  The softmax scale d^-1/2 (attention.py:153,199)
    and log2(e) (the kernel uses exp2) are folded into the q rows in fp32 before the bf16 rounding."""
    inner, C = wq.shape
    d = inner // heads
    dp = head_pad(d)

    def padded(w: torch.Tensor, s: float) -> torch.Tensor:
        o = torch.zeros(heads, dp, w.shape[1], dtype=torch.float32, device=w.device)
        o[:, :d] = w.float().reshape(heads, d, -1) * s
        return o.reshape(heads * dp, -1)

    s = (d ** -0.5) * math.log2(math.e) if fold_scale else 1.0
    parts = [padded(wq, s)]
    if wk is not None:
        parts.append(padded(wk, 1.0))
    return torch.cat(parts, 0).contiguous().to(torch.bfloat16)
