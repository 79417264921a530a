"""TEST INFRASTRUCTURE - CPU oracle of the AdaFace / SD-1.5 denoising hot path (fp32, plain torch).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file, and only as the checker / baseline.  The product path (adaprompt_b200/*) never does.

This is a *restatement*, not a copy: a functional, state_dict-driven fp32 evaluation of the same
arithmetic the reference modules perform, each function citing the reference lines it follows.
Pinning: the reference ships no golden vectors, so the oracle is pinned against
outputs of the unmodified reference modules executed in the build container on the synthetic-weight
recipe (adaprompt_b200/weights.py); the generating script is oracle/make_golden.py and the vectors
live in tests/golden/.  tests/test_oracle_golden.py re-checks the pin on every CPU test run.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# UNet layer index -> cross-attention layer index (openaimodel.py:876-877)
LAYER2CA = {1: 0, 2: 1, 4: 2, 5: 3, 7: 4, 8: 5, 12: 6, 16: 7, 17: 8, 18: 9, 19: 10, 20: 11, 21: 12, 22: 13,
            23: 14, 24: 15}


# ---------------------------------------------------------------------------------------------------
# architecture spec (openaimodel.py:447-703 with v1-inference-ada.yaml:35-51)
# ---------------------------------------------------------------------------------------------------
class UNetSpec:
    def __init__(self, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768):
        self.in_channels, self.out_channels, self.model_channels = in_channels, out_channels, model_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, tuple(attention_resolutions)
        self.channel_mult, self.num_heads, self.context_dim = tuple(channel_mult), num_heads, context_dim

    def blocks(self):
        """Returns (input_blocks, middle, output_blocks): each block a list of ('conv_in'|'res'|'attn'|'down'|'up', ...)."""
        mc = self.model_channels
        inp: List[List[tuple]] = [[("conv_in", self.in_channels, mc)]]
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(self.num_res_blocks):
                layers = [("res", ch, mult * mc)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(("attn", ch, ch // self.num_heads))
                inp.append(layers)
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                inp.append([("down", ch)])
                chans.append(ch)
                ds *= 2
        mid = [("res", ch, ch), ("attn", ch, ch // self.num_heads), ("res", ch, ch)]
        out: List[List[tuple]] = []
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(self.num_res_blocks + 1):
                ich = chans.pop()
                layers = [("res", ch + ich, mc * mult)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(("attn", ch, ch // self.num_heads))
                if level and i == self.num_res_blocks:
                    layers.append(("up", ch))
                    ds //= 2
                out.append(layers)
        return inp, mid, out

    def state_spec(self) -> "OrderedDict[str, Tuple[int, ...]]":
        """key -> shape, same names as the reference state_dict."""
        spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        mc, ted = self.model_channels, self.model_channels * 4

        def lin(p, i, o, bias=True):
            spec[p + ".weight"] = (o, i)
            if bias:
                spec[p + ".bias"] = (o,)

        def conv(p, i, o, k):
            spec[p + ".weight"] = (o, i, k, k)
            spec[p + ".bias"] = (o,)

        def norm(p, c):
            spec[p + ".weight"] = (c,)
            spec[p + ".bias"] = (c,)

        def res(p, i, o):
            norm(p + ".in_layers.0", i)
            conv(p + ".in_layers.2", i, o, 3)
            lin(p + ".emb_layers.1", ted, o)
            norm(p + ".out_layers.0", o)
            conv(p + ".out_layers.3", o, o, 3)
            if i != o:
                conv(p + ".skip_connection", i, o, 1)

        def attn(p, c):
            norm(p + ".norm", c)
            conv(p + ".proj_in", c, c, 1)
            b = p + ".transformer_blocks.0"
            for a, cd in (("attn1", c), ("attn2", self.context_dim)):
                lin(f"{b}.{a}.to_q", c, c, False)
                lin(f"{b}.{a}.to_k", cd, c, False)
                lin(f"{b}.{a}.to_v", cd, c, False)
                lin(f"{b}.{a}.to_out.0", c, c)
                if a == "attn1":
                    lin(f"{b}.ff.net.0.proj", c, 8 * c)
                    lin(f"{b}.ff.net.2", 4 * c, c)
            for n in ("norm1", "norm2", "norm3"):
                norm(f"{b}.{n}", c)
            conv(p + ".proj_out", c, c, 1)

        def emit(prefix, layers):
            for j, l in enumerate(layers):
                p = f"{prefix}.{j}"
                if l[0] == "conv_in":
                    conv(p, l[1], l[2], 3)
                elif l[0] == "res":
                    res(p, l[1], l[2])
                elif l[0] == "attn":
                    attn(p, l[1])
                elif l[0] == "down":
                    conv(p + ".op", l[1], l[1], 3)
                elif l[0] == "up":
                    conv(p + ".conv", l[1], l[1], 3)

        lin("time_embed.0", mc, ted)
        lin("time_embed.2", ted, ted)
        inp, mid, out = self.blocks()
        for i, layers in enumerate(inp):
            emit(f"input_blocks.{i}", layers)
        emit("middle_block", mid)
        for i, layers in enumerate(out):
            emit(f"output_blocks.{i}", layers)
        norm("out.0", mc)
        conv("out.2", mc, self.out_channels, 3)
        return spec


# ---------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """ldm/modules/diffusionmodules/util.py:154-174 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def group_norm32(x, w, b, eps=1e-5):
    """GroupNorm32 (util.py:217-219): fp32 GroupNorm with 32 groups."""
    return F.group_norm(x.float(), 32, w, b, eps)


def res_block(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward (openaimodel.py:259-279), use_scale_shift_norm=False, no up/down."""
    h = F.silu(group_norm32(x, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"]))
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.silu(group_norm32(h, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"]))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def conv_attn_columns(sim_point: torch.Tensor, cols: torch.Tensor, infeat_size, ks: int) -> torch.Tensor:
    """replace_rows_by_conv_attn (ldm/util.py:700-878) for ONE sample, restated on pointwise scores: the grouped conv2d
    of the zero-padded query map with the ks*ks subject keys as kernel taps (:797-809) is the sum over taps (i, j) of the
    pointwise score map of tap key ks*i+j read at pixel offset (i - pad, j - pad); token m = (dy + pad) * ks + (dx + pad)
    receives that map shifted by (dy, dx) with zeros shifted in (:811-845), all divided by ks^1.5 (:766,:803).
    sim_point [H, N, T] (scaled scores of this sample), cols [ks*ks] token positions -> [ks*ks, H, N]."""
    Hf, Wf = infeat_size
    Hh = sim_point.shape[0]
    pad = (ks - 1) // 2
    g = sim_point[:, :, cols].reshape(Hh, Hf, Wf, ks * ks)
    gp = F.pad(g.permute(0, 3, 1, 2), (ks, ks, ks, ks))                      # [H, taps, Hf + 2ks, Wf + 2ks]
    A = torch.zeros(Hh, Hf, Wf, dtype=sim_point.dtype)
    for i in range(ks):
        for j in range(ks):
            A = A + gp[:, ks * i + j, ks + i - pad: ks + i - pad + Hf, ks + j - pad: ks + j - pad + Wf]
    A = A / ks ** 1.5
    Ap = F.pad(A, (ks, ks, ks, ks))
    out = []
    for m in range(ks * ks):
        dy, dx = m // ks - pad, m % ks - pad
        out.append(Ap[:, ks - dy: ks - dy + Hf, ks - dx: ks - dx + Wf].reshape(Hh, Hf * Wf))
    return torch.stack(out, 0)


def cross_attention(sd: SD, p: str, x: torch.Tensor, context=None, mask=None, heads: int = 8, conv_attn=None,
                    capture: Optional[dict] = None) -> torch.Tensor:
    """CrossAttention.forward (attention.py:172-243); context None -> self-attention; context may be a
    (v_context, k_context) tuple (:190-193); mask is a key mask [B, ...] (:223-232).  conv_attn = (placeholder2indices,
    infeat_size, ks) replaces score columns (:208-216); capture: dict filled with q / attn / attnscore (:245-255)."""
    B, N, C = x.shape
    q = F.linear(x, sd[p + ".to_q.weight"])
    if context is None:
        v_ctx = k_ctx = x
    elif isinstance(context, (tuple, list)):
        v_ctx, k_ctx = context
    else:
        v_ctx = k_ctx = context
    k = F.linear(k_ctx, sd[p + ".to_k.weight"])
    v = F.linear(v_ctx, sd[p + ".to_v.weight"])
    d = C // heads

    def split(t):
        return t.reshape(B, -1, heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    if conv_attn is not None and context is not None:
        p2i, infeat_size, ks = conv_attn
        if p2i is not None and ks > 1:
            sim2 = sim.clone()
            for subj in p2i:
                iB, iN = p2i[subj]
                uniq = torch.unique(iB)
                M = len(iN) // len(uniq)
                for bi, b_ in enumerate(uniq.tolist()):
                    cols = iN[bi * M: bi * M + ks * ks]
                    sim2[b_][:, :, cols] = conv_attn_columns(sim[b_], cols, infeat_size, ks).permute(1, 2, 0)
            sim = sim2
    if mask is not None:
        m = mask.reshape(B, -1).bool()
        sim = sim.masked_fill(~m[:, None, None, :], -torch.finfo(sim.dtype).max)
    attn = sim.softmax(dim=-1)
    if capture is not None:
        capture.update(q=q * (d ** -0.5) ** 0.5, attn=attn, attnscore=sim)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(B, N, C)
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def feed_forward(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward with GEGLU (attention.py:32-59): first half value, second half gate, exact GELU."""
    h = F.linear(x, sd[p + ".net.0.proj.weight"], sd[p + ".net.0.proj.bias"])
    val, gate = h.chunk(2, dim=-1)
    return F.linear(val * F.gelu(gate), sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])


def basic_transformer_block(sd: SD, p: str, x, context=None, mask=None, heads: int = 8, conv_attn=None, capture=None):
    """BasicTransformerBlock._forward (attention.py:275-285): the mask goes to self-attention only."""
    def ln(t, n):
        return F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"], 1e-5)

    x1 = cross_attention(sd, p + ".attn1", ln(x, "norm1"), None, mask, heads) + x
    x2 = x1 + cross_attention(sd, p + ".attn2", ln(x1, "norm2"), context, None, heads, conv_attn, capture)
    return feed_forward(sd, p + ".ff", ln(x2, "norm3")) + x2


def spatial_transformer(sd: SD, p: str, x: torch.Tensor, context=None, mask=None, heads: int = 8, conv=None, capture=None):
    """SpatialTransformer.forward (attention.py:321-341): GroupNorm eps 1e-6, 1x1 convs, depth 1.  conv = (placeholder2indices,
    kernel size) of this layer's conv attention or None; capture: dict for the attn2 activations."""
    B, C, H, W = x.shape
    x_in = x
    h = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = F.conv2d(h, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    m2 = F.interpolate(mask, size=(H, W), mode="nearest") if mask is not None else None
    conv_attn = (conv[0], (H, W), conv[1]) if conv is not None else None           # infeat_size (:330)
    h = basic_transformer_block(sd, p + ".transformer_blocks.0", h, context, m2, heads, conv_attn, capture)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + x_in


def unet_forward(sd: SD, spec: UNetSpec, x: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor,
                 extra_info: Optional[dict] = None) -> torch.Tensor:
    """UNetModel.forward (openaimodel.py:827-1052) for the layerwise-context inference path."""
    extra_info = extra_info or {}
    B = x.shape[0]
    mask = extra_info.get("img_mask", None)
    iter_type = extra_info.get("iter_type", "normal_recon")
    emb = F.linear(timestep_embedding(timesteps, spec.model_channels), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    ctx = context.reshape(B, 16, -1, context.shape[-1]).permute(1, 0, 2, 3)  # :866

    conv_ks = extra_info.get("use_conv_attn_kernel_size", -1) or -1
    p2i = extra_info.get("placeholder2indices", None)
    conv_sizes = [conv_ks] * 16
    if conv_ks > 0:
        conv_sizes[6:11] = [1] * 5                                               # openaimodel.py:932
    distill_layers = [7, 8, 12, 16, 17, 18, 19, 20, 21, 22, 23, 24] if extra_info.get("capture_distill_attn", False) else []
    acts: Dict[int, dict] = {}
    compel_prob = extra_info.get("apply_compel_cfg_prob", 0)
    empty_ctx, compel_range = extra_info.get("empty_context", None), extra_info.get("compel_cfg_weight_level_range", None)
    is_training = extra_info.get("is_training", True)

    def compel(e, level, batch_mask, gate_prob):
        """prob_apply_compel_cfg (ldm/util.py:1823-1854) on one tensor; draws from Python's global `random`."""
        import random
        if empty_ctx is None or level is None or random.random() > gate_prob:
            return e
        e2 = (e - empty_ctx) * (1.1 ** level) + empty_ctx
        m = batch_mask.reshape(-1, 1, 1)
        return e2 * m + e * (1 - m)

    def layer_ctx(layer_idx):
        c = ctx[LAYER2CA[layer_idx]]
        if iter_type == "mix_hijk":  # :885-892, (v, k) halves along the token dim
            v, k = c.chunk(2, dim=1)
        else:
            v = k = c
        if compel_prob > 0:          # :898-916
            import random
            bm = torch.ones(B, dtype=torch.float32, device=x.device)
            if is_training:
                if random.random() < 0.5:
                    bm[:B // 2] = 0
            else:
                bm[B // 2:] = 0
            if not (empty_ctx is None or compel_range is None or random.random() > compel_prob):
                level = random.uniform(*compel_range) if isinstance(compel_range, (list, tuple)) else compel_range
                v, k = compel(v, level, bm, 1), compel(k, level, bm, 1)     # recursion over the (v, k) tuple, :1837
        return (v, k)

    def run(prefix, layers, h, layer_idx):
        for j, l in enumerate(layers):
            p = f"{prefix}.{j}"
            if l[0] == "conv_in":
                h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
            elif l[0] == "res":
                h = res_block(sd, p, h, emb)
            elif l[0] == "attn":
                cap = {} if layer_idx in distill_layers else None
                ks = conv_sizes[LAYER2CA[layer_idx]]
                h = spatial_transformer(sd, p, h, layer_ctx(layer_idx), mask, spec.num_heads,
                                        (p2i, ks) if (conv_ks > 0 and p2i is not None) else None, cap)
                if cap is not None:
                    acts[layer_idx] = cap
            elif l[0] == "down":
                h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
            elif l[0] == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        return h

    inp, mid, out = spec.blocks()
    hs = []
    h = x.float()
    layer_idx = 0
    def outfeat(h, layer_idx):                     # openaimodel.py:984-988: the block output next to the attn2 activations
        if layer_idx in acts:
            acts[layer_idx]["outfeat"] = h

    for i, layers in enumerate(inp):
        h = run(f"input_blocks.{i}", layers, h, layer_idx)
        hs.append(h)
        outfeat(h, layer_idx)
        layer_idx += 1
    h = run("middle_block", mid, h, layer_idx)
    outfeat(h, layer_idx)
    layer_idx += 1
    for i, layers in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)  # :1019
        h = run(f"output_blocks.{i}", layers, h, layer_idx)
        outfeat(h, layer_idx)
        layer_idx += 1
    if extra_info is not None and distill_layers:  # :1031-1035
        extra_info["ca_layers_activations"] = {k: {li: acts[li][k] for li in acts} for k in ("outfeat", "attn", "attnscore", "q")}
    h = F.silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"]))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ---------------------------------------------------------------------------------------------------
# DDIM sampler (ldm/models/diffusion/ddim.py) + schedule (ddpm.py:240-292, util.py:21-77)
# ---------------------------------------------------------------------------------------------------
def make_alphas_cumprod(n_timestep=1000, linear_start=0.00085, linear_end=0.012) -> np.ndarray:
    """make_beta_schedule 'linear' (util.py:21-25) + cumprod in float64 (ddpm.py:247-249)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return np.cumprod(1.0 - betas.numpy(), axis=0)


def ddim_schedule(S: int, eta: float = 0.0, n_timestep: int = 1000):
    """make_ddim_timesteps 'uniform' + make_ddim_sampling_parameters (util.py:46-77), fp32 like
    DDIMSampler.make_schedule (ddim.py:28-68): alphas_cumprod is cast to fp32 *before* indexing."""
    ac = torch.tensor(make_alphas_cumprod(n_timestep), dtype=torch.float32)
    c = n_timestep // S
    ts = np.asarray(list(range(0, n_timestep, c))) + 1
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())  # float64 numpy array of fp32 values
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    sqrt_one_minus_alphas = np.sqrt(1. - alphas)
    return ts, alphas, alphas_prev, sigmas, sqrt_one_minus_alphas


def guidance_schedule(S: int, guidance_scale) -> List[float]:
    """ddim.py:168-180,215-218: tuple (max, min), linear anneal by repeated python-float subtraction."""
    max_g, min_g = guidance_scale
    delta = (max_g - min_g) / (S - 1)
    g, out = max_g, []
    for i in range(S):
        out.append(g)
        g = g - delta if i <= S - 1 else 1
    return out


def ddim_sample(apply_model, S: int, shape, cond, uncond, guidance_scale, x_T: torch.Tensor, eta: float = 0.0,
                log_every_t: int = 100, generator: Optional[torch.Generator] = None):
    """DDIMSampler.sample / ddim_sampling / p_sample_ddim (ddim.py:71-296) for the tuple-conditioning path.
    `apply_model(x, t, cond_tuple)` plays LatentDiffusion.apply_model."""
    ts, alphas, alphas_prev, sigmas, s1m = ddim_schedule(S, eta)
    b = shape[0]
    img = x_T
    inter = {"x_inter": [img], "pred_x0": [img]}
    gs = guidance_schedule(S, guidance_scale)
    for i, step in enumerate(np.flip(ts)):
        index = S - i - 1
        t = torch.full((b,), int(step), dtype=torch.long, device=img.device)
        g = gs[i]
        if uncond is None or g == 1.0:
            e_t = apply_model(img, t, cond)
        else:
            c_c, c_in_c, extra = cond
            c_u, c_in_u, _ = uncond
            c2 = (torch.cat([c_c, c_u]), sum([c_in_c, c_in_u], []), extra)
            e_t, e_u = apply_model(torch.cat([img] * 2), torch.cat([t] * 2), c2).chunk(2)
            e_t = e_u + g * (e_t - e_u)
        full = lambda v: torch.full((b, 1, 1, 1), float(v), device=img.device)
        a_t, a_prev, sigma_t, s1m_t = full(alphas[index]), full(alphas_prev[index]), full(sigmas[index]), full(s1m[index])
        pred_x0 = (img - s1m_t * e_t) / a_t.sqrt()
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * torch.randn(img.shape, generator=generator, device=img.device)
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
        if index % log_every_t == 0 or index == S - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter
