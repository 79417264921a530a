"""Teacher of the Stage-1 distillation step: Arc2FaceWrapper
(ldm/models/diffusion/ddpm.py:5402-5478).

The reference teacher is a SECOND SD-1.5 UNet - `diffusers.UNet2DConditionModel` loaded from 'models/arc2face' in fp16 -
run without gradients for 1..10 denoising steps on the Arc2Face ID prompt embeddings.  The architecture is the UNet this
package already implements; what differs is (i) the parameter NAMES (diffusers layout), handled by
`convert_diffusers_unet_state_dict`, the inverse of the well-known ldm -> diffusers renaming, and (ii) the conditioning:
one plain [B, L, 768] context (L = 21, ddpm.py:5427) shared by all 16 cross-attention layers instead of AdaFace's
layerwise [16 B, 77, 768] - expressed here by repeating the context once per layer, which is what the layerwise split at
openaimodel.py:866 undoes.

diffusers and the Arc2Face weights are absent offline, so the end-to-end numerics of this wrapper are "parity unpinned";
the renaming is checked as a bijection onto the exact key / shape set of UNetModel and the step logic against a
hand-rolled restatement with a stand-in UNet (tests/test_host_logic.py), the UNet itself by the U1-U8 parity tests.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

N_CA_LAYERS = 16


def _unet_key_map(channel_mult_len: int = 4, num_res_blocks: int = 2) -> List[Tuple[str, str]]:
    """(ldm prefix, diffusers prefix) pairs for the SD-1.5 UNet (model_channels 320, channel_mult [1,2,4,4], 2 res
    blocks, attention at the first three levels: v1-inference-ada.yaml:35-51)."""
    m: List[Tuple[str, str]] = [
        ("time_embed.0.", "time_embedding.linear_1."), ("time_embed.2.", "time_embedding.linear_2."),
        ("input_blocks.0.0.", "conv_in."), ("out.0.", "conv_norm_out."), ("out.2.", "conv_out."),
    ]
    n = channel_mult_len
    for i in range(n):
        for j in range(num_res_blocks):                                   # down blocks
            sd = f"input_blocks.{(num_res_blocks + 1) * i + j + 1}."
            m.append((sd + "0.", f"down_blocks.{i}.resnets.{j}."))
            if i < n - 1:
                m.append((sd + "1.", f"down_blocks.{i}.attentions.{j}."))
        for j in range(num_res_blocks + 1):                               # up blocks
            sd = f"output_blocks.{(num_res_blocks + 1) * i + j}."
            m.append((sd + "0.", f"up_blocks.{i}.resnets.{j}."))
            if i > 0:
                m.append((sd + "1.", f"up_blocks.{i}.attentions.{j}."))
        if i < n - 1:
            m.append((f"input_blocks.{(num_res_blocks + 1) * (i + 1)}.0.op.", f"down_blocks.{i}.downsamplers.0.conv."))
            m.append((f"output_blocks.{(num_res_blocks + 1) * i + num_res_blocks}.{1 if i == 0 else 2}.",
                      f"up_blocks.{i}.upsamplers.0."))
    m.append(("middle_block.1.", "mid_block.attentions.0."))
    for j in range(2):
        m.append((f"middle_block.{2 * j}.", f"mid_block.resnets.{j}."))
    return m


_RESNET_MAP = [("in_layers.0.", "norm1."), ("in_layers.2.", "conv1."), ("out_layers.0.", "norm2."),
               ("out_layers.3.", "conv2."), ("emb_layers.1.", "time_emb_proj."), ("skip_connection.", "conv_shortcut.")]


def convert_diffusers_unet_state_dict(sd_hf: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """diffusers UNet2DConditionModel state_dict (SD-1.5) -> the reference / adaprompt_b200.unet.UNetModel key names.
    Transformer internals (norm, proj_in, transformer_blocks.N.attn1.to_q ..., proj_out) carry the same names in both
    layouts; diffusers >= 0.20 stores `to_out.0`, `ff.net.0.proj`, `ff.net.2` exactly like the reference.  1x1
    projections stored as Linear [C, C] are reshaped to the reference's Conv2d [C, C, 1, 1]."""
    blocks = sorted(_unet_key_map(), key=lambda p: -len(p[1]))            # longest diffusers prefix first
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd_hf.items():
        nk = None
        for sd_p, hf_p in blocks:
            if k.startswith(hf_p):
                rest = k[len(hf_p):]
                if ".resnets." in hf_p:
                    for sd_r, hf_r in _RESNET_MAP:
                        if rest.startswith(hf_r):
                            rest = sd_r + rest[len(hf_r):]
                            break
                nk = sd_p + rest
                break
        if nk is None:
            raise KeyError(f"convert_diffusers_unet_state_dict: no rule for {k!r}")
        if (nk.endswith("proj_in.weight") or nk.endswith("proj_out.weight")) and v.dim() == 2:
            v = v[:, :, None, None]
        out[nk] = v
    return out


def convert_ldm_unet_state_dict_to_diffusers(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse renaming (used by the round-trip test and to export weights to a diffusers pipeline)."""
    blocks = sorted(_unet_key_map(), key=lambda p: -len(p[0]))
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        nk = None
        for sd_p, hf_p in blocks:
            if k.startswith(sd_p):
                rest = k[len(sd_p):]
                if ".resnets." in hf_p:
                    for sd_r, hf_r in _RESNET_MAP:
                        if rest.startswith(sd_r):
                            rest = hf_r + rest[len(sd_r):]
                            break
                nk = hf_p + rest
                break
        if nk is None:
            raise KeyError(f"convert_ldm_unet_state_dict_to_diffusers: no rule for {k!r}")
        out[nk] = v
    return out


class Arc2FaceTeacher(torch.nn.Module):
    """ddpm.py:5402-5478 on adaprompt_b200.unet.UNetModel.  `unet` holds the Arc2Face weights (load them with
    convert_diffusers_unet_state_dict); `forward` has the reference's signature and return value."""

    def __init__(self, unet: torch.nn.Module):
        super().__init__()
        self.unet = unet
        for p in self.unet.parameters():
            p.requires_grad = False

    @staticmethod
    def layerwise(context: torch.Tensor) -> torch.Tensor:
        """[B, L, 768] -> [16 B, L, 768] in the '(b l)' order the UNet splits (openaimodel.py:866)."""
        return context.float().repeat_interleave(N_CA_LAYERS, dim=0).contiguous()

    @staticmethod
    def predict_start_from_noise(ddpm_model, x_t, t, noise):
        """ddpm.py:331-335: sqrt(1 / a_t) x_t - sqrt(1 / a_t - 1) noise."""
        a = ddpm_model.alphas_cumprod.to(x_t.device)[t].view(-1, 1, 1, 1)
        return torch.sqrt(1.0 / a) * x_t - torch.sqrt(1.0 / a - 1) * noise

    @torch.no_grad()
    def forward(self, ddpm_model, x_start, noise, t, context, num_denoising_steps=1):
        assert num_denoising_steps <= 10                                                  # :5434
        x_starts, noises, ts, noise_preds = [x_start], [noise], [t], []
        ctx = self.layerwise(context)
        extra = {"use_layerwise_context": True, "use_conv_attn_kernel_size": -1, "placeholder2indices": None,
                 "is_training": False}
        for i in range(num_denoising_steps):
            x_start, t, noise = x_starts[i], ts[i], noises[i]
            x_noisy = ddpm_model.q_sample(x_start, t, noise)                              # :5448
            noise_pred = self.unet(x_noisy, t, context=ctx, extra_info=dict(extra))       # :5451-5452
            noise_preds.append(noise_pred)
            x_starts.append(self.predict_start_from_noise(ddpm_model, x_noisy, t, noise_pred))   # :5456-5457
            if i < num_denoising_steps - 1:
                relative_ts = torch.rand_like(t.float())                                  # :5461 (uniform, as the reference notes)
                t_lb = t * np.power(0.5, np.power(num_denoising_steps - 1, -0.3))         # :5467
                t_ub = t * np.power(0.7, np.power(num_denoising_steps - 1, -0.3))         # :5468
                ts.append(((t_ub - t_lb) * relative_ts + t_lb).long())                    # :5469-5473
                noises.append(torch.randn_like(x_starts[-1]))                             # :5475-5476
        return noise_preds, x_starts[1:], noises, ts                                      # :5478-5480
