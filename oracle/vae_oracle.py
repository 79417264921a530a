"""TEST INFRASTRUCTURE - CPU oracle of the first-stage decoder (latents -> RGB), fp32, plain torch.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file, and only as the checker;
the product path (adaprompt_b200/vae.py) never does.

A functional, state_dict-driven restatement of ldm/modules/diffusionmodules/model.py (Decoder :502-609, ResnetBlock
:83-142, AttnBlock :151-242, Upsample :43-58, Normalize :39-40, nonlinearity :34-36), AutoencoderKL.decode
(ldm/models/autoencoder.py:330-333) and LatentDiffusion.decode_first_stage (ldm/models/diffusion/ddpm.py:1260-1318,
non-split branch).  Pinning: the reference ships no golden vectors; oracle/make_golden_vae.py runs the UNMODIFIED
reference Decoder (importable in the build container) on the synthetic-weight recipe and stores its outputs in
tests/golden/vae.pt; tests/test_vae_oracle_golden.py re-checks the pin on every CPU run.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


class VAESpec:
    """configs/stable-diffusion/v1-inference-ada.yaml:56-72."""

    def __init__(self, ch=128, out_ch=3, ch_mult: Sequence[int] = (1, 2, 4, 4), num_res_blocks=2, z_channels=4,
                 embed_dim=4):
        self.ch, self.out_ch, self.ch_mult = ch, out_ch, tuple(ch_mult)
        self.num_res_blocks, self.z_channels, self.embed_dim = num_res_blocks, z_channels, embed_dim

    def state_spec(self) -> "OrderedDict[str, tuple]":
        """Key -> shape of the decode-side parameters, in the reference's registration order (model.py:527-573)."""
        sp: "OrderedDict[str, tuple]" = OrderedDict()

        def conv(k, co, ci, ks):
            sp[k + ".weight"] = (co, ci, ks, ks)
            sp[k + ".bias"] = (co,)

        def norm(k, c):
            sp[k + ".weight"] = (c,)
            sp[k + ".bias"] = (c,)

        def res(k, ci, co):
            norm(k + ".norm1", ci)
            conv(k + ".conv1", co, ci, 3)
            norm(k + ".norm2", co)
            conv(k + ".conv2", co, co, 3)
            if ci != co:
                conv(k + ".nin_shortcut", co, ci, 1)

        nres = len(self.ch_mult)
        block_in = self.ch * self.ch_mult[nres - 1]
        conv("decoder.conv_in", block_in, self.z_channels, 3)
        res("decoder.mid.block_1", block_in, block_in)
        norm("decoder.mid.attn_1.norm", block_in)
        for n in ("q", "k", "v", "proj_out"):
            conv("decoder.mid.attn_1." + n, block_in, block_in, 1)
        res("decoder.mid.block_2", block_in, block_in)
        ups = []
        for i_level in reversed(range(nres)):
            block_out = self.ch * self.ch_mult[i_level]
            keys = []
            for i_block in range(self.num_res_blocks + 1):
                keys.append((f"decoder.up.{i_level}.block.{i_block}", block_in, block_out))
                block_in = block_out
            ups.append((i_level, keys, block_in))
        for i_level, keys, c in sorted(ups):          # ModuleList order: up.0 first (model.py:563 insert(0, ...))
            for k, ci, co in keys:
                res(k, ci, co)
            if i_level != 0:
                conv(f"decoder.up.{i_level}.upsample.conv", c, c, 3)
        norm("decoder.norm_out", block_in)
        conv("decoder.conv_out", self.out_ch, block_in, 3)
        conv("post_quant_conv", self.z_channels, self.embed_dim, 1)
        return sp


def _gn(sd: SD, k: str, x: torch.Tensor) -> torch.Tensor:
    return F.group_norm(x, 32, sd[k + ".weight"], sd[k + ".bias"], eps=1e-6)          # model.py:39-40


def _swish(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)                                                       # model.py:34-36


def _conv(sd: SD, k: str, x: torch.Tensor, pad: int) -> torch.Tensor:
    return F.conv2d(x, sd[k + ".weight"], sd[k + ".bias"], padding=pad)


def resnet_block(sd: SD, k: str, x: torch.Tensor) -> torch.Tensor:
    """model.py:122-142 with temb = None."""
    h = _conv(sd, k + ".conv1", _swish(_gn(sd, k + ".norm1", x)), 1)
    h = _conv(sd, k + ".conv2", _swish(_gn(sd, k + ".norm2", h)), 1)
    if (k + ".nin_shortcut.weight") in sd:
        x = _conv(sd, k + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd: SD, k: str, x: torch.Tensor) -> torch.Tensor:
    """model.py:179-242 with mask = None."""
    h_ = _gn(sd, k + ".norm", x)
    q, kk, v = (_conv(sd, f"{k}.{n}", h_, 0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    kk = kk.reshape(b, c, h * w)
    w_ = torch.bmm(q, kk) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, k + ".proj_out", h_, 0)


def decoder_forward(sd: SD, spec: VAESpec, z: torch.Tensor) -> torch.Tensor:
    """model.py:575-609 (give_pre_end False, tanh_out False)."""
    h = _conv(sd, "decoder.conv_in", z, 1)
    h = resnet_block(sd, "decoder.mid.block_1", h)
    h = attn_block(sd, "decoder.mid.attn_1", h)
    h = resnet_block(sd, "decoder.mid.block_2", h)
    for i_level in reversed(range(len(spec.ch_mult))):
        for i_block in range(spec.num_res_blocks + 1):
            h = resnet_block(sd, f"decoder.up.{i_level}.block.{i_block}", h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")                    # model.py:55
            h = _conv(sd, f"decoder.up.{i_level}.upsample.conv", h, 1)
    h = _swish(_gn(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h, 1)


def decode(sd: SD, spec: VAESpec, z: torch.Tensor) -> torch.Tensor:
    """autoencoder.py:330-333."""
    return decoder_forward(sd, spec, _conv(sd, "post_quant_conv", z, 0))


def decode_first_stage(sd: SD, spec: VAESpec, z: torch.Tensor, scale_factor: float = 0.18215) -> torch.Tensor:
    """ddpm.py:1267 then :1318 (first_stage_model.decode)."""
    return decode(sd, spec, 1. / scale_factor * z)


def vae_latents(name: str) -> torch.Tensor:
    """Seeded test latents (sampler-output scale: ~N(0,1))."""
    shapes = {"b2_8": ((2, 4, 8, 8), 31), "b1_16": ((1, 4, 16, 16), 32), "b2_32": ((2, 4, 32, 32), 33),
              "b1_64": ((1, 4, 64, 64), 34)}
    shape, seed = shapes[name]
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))
