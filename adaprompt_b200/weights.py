"""Deterministic synthetic SD-1.5-architecture weights.

There is no network in the build or GPU containers, so parity and benchmark runs use
random-init weights of the real architecture.  The recipe is a pure function of
``(key, shape, seed)`` so that the reference modules (imported from /root/reference in the
build container only), the CPU oracle and the CUDA path can all be fed the *same*
``state_dict`` without shipping 3.4 GB of floats:

* every tensor has its own generator seeded with ``seed ^ crc32(key)``;
* conv / linear weights and biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default bound);
* norm scales ~ 1 + 0.05 N(0,1), norm shifts ~ 0.05 N(0,1);
* the 39 layers the reference zero-initialises (``zero_module`` at
  ldm/modules/diffusionmodules/openaimodel.py:233-235,696 and ldm/modules/attention.py:313)
  are randomised like any other layer - otherwise the UNet output is identically zero
  and parity checks nothing ("Random-init trap").
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Iterable, Tuple

import torch

Shape = Tuple[int, ...]


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 0x9E3779B1 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def synth_tensor(key: str, shape: Shape, seed: int, fan_in: int | None = None) -> torch.Tensor:
    """One tensor of the recipe.  ``fan_in`` is needed for biases (taken from the sibling weight)."""
    g = _gen(key, seed)
    if len(shape) >= 2:  # conv / linear / embedding weight
        fi = 1
        for s in shape[1:]:
            fi *= s
        b = 1.0 / math.sqrt(fi)
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * b
    if fan_in is not None:  # bias of a conv / linear
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * b
    # 1-D parameter of a normalisation layer
    n = torch.randn(shape, generator=g, dtype=torch.float32) * 0.05
    return 1.0 + n if key.endswith("weight") else n


def synth_state_dict(spec: "OrderedDict[str, Shape] | Iterable[Tuple[str, Shape]]", seed: int = 1234
                     ) -> "OrderedDict[str, torch.Tensor]":
    """Build a full state_dict for ``spec`` (ordered ``key -> shape``)."""
    spec = OrderedDict(spec)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in spec.items():
        fan_in = None
        if len(shape) == 1 and key.endswith(".bias"):
            wkey = key[: -len("bias")] + "weight"
            if wkey in spec and len(spec[wkey]) >= 2:
                fan_in = 1
                for s in spec[wkey][1:]:
                    fan_in *= s
        out[key] = synth_tensor(key, tuple(shape), seed, fan_in)
    return out


def spec_of(module: torch.nn.Module) -> "OrderedDict[str, Shape]":
    return OrderedDict((k, tuple(v.shape)) for k, v in module.state_dict().items())
