"""Host mirror of ldm/modules/diffusionmodules/openaimodel.py (reference: askerlee/adaprompt).

UNetModel :417, ResBlock :167, TimestepEmbedSequential :76, Upsample :95, Downsample :138 with the
reference constructor / forward signatures and state_dict key names, so an SD-1.5 checkpoint (or the
synthetic recipe of adaprompt_b200/weights.py) loads unchanged.  The torch.nn layers only hold
parameters; all arithmetic runs in libadaface_b200.so (NHWC, fp32 residual stream, bf16 tensor-core
operands).  Only the configuration the reference actually instantiates is supported
(configs/stable-diffusion/v1-inference-ada.yaml:35-51: dims=2, use_spatial_transformer, no scale-shift
norm, no resblock up/down, no class conditioning); anything else raises NotImplementedError.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import ops
from .attention import ContextKV, PackedModule, SpatialTransformer, exists, invalidate_all, zero_module
from .packing import pack_conv1x1, pack_conv3x3

ALL_CA_LAYER_INDICES = [1, 2, 4, 5, 7, 8, 12, 16, 17, 18, 19, 20, 21, 22, 23, 24]
L2CA = {1: 0, 2: 1, 4: 2, 5: 3, 7: 4, 8: 5, 12: 6, 16: 7, 17: 8, 18: 9, 19: 10, 20: 11, 21: 12, 22: 13, 23: 14,
        24: 15}


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for util.py:217-219 (fp32 GroupNorm, 32 groups, eps 1e-5)."""


def normalization(channels):
    return GroupNorm32(32, channels)


def extract_layerwise_value(v, layer_idx, v_is_layerwise_array, v_is_layerwise_dict):
    """ldm/util.py:1436-1447."""
    if v_is_layerwise_array:
        return v[layer_idx]
    if v_is_layerwise_dict:
        return {k2: v2[layer_idx] for k2, v2 in v.items()}
    return v


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.float().permute(0, 2, 3, 1).contiguous()


class Act:
    """An fp32 NHWC activation of the residual stream plus the per-channel GroupNorm partial statistics its
    producer's epilogue wrote (ops.GNStats) - the GroupNorm that consumes it never re-reads it for statistics.
    Tensors no tensor-core kernel produced get a stand-alone statistics pass on first use (cached: skip tensors
    are normalised twice, openaimodel.py:982,1019)."""

    __slots__ = ("t", "st")

    def __init__(self, t: torch.Tensor, st=None):
        self.t, self.st = t, st

    def stats(self):
        if self.st is None:
            self.st = ops.groupnorm_stats(self.t)
        return self.st


def _conv_out_act(B, Ho, Wo, C, device):
    out = torch.empty(B, Ho, Wo, C, dtype=torch.float32, device=device)
    st = ops.gn_stats_for_conv(B, Ho, Wo, C, device)
    return out, st


def _nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).contiguous()


class TimestepBlock(nn.Module):
    """Any module whose forward takes timestep embeddings as a second argument (openaimodel.py:63-73)."""


class Upsample(PackedModule):
    """openaimodel.py:95-123: nearest x2 then conv3x3 (fused: upsample+cast kernel, implicit-GEMM conv)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or not use_conv or padding != 1:
            raise NotImplementedError("Upsample: only dims=2, use_conv=True, padding=1")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=padding)

    def _pack(self):
        return {"w": pack_conv3x3(self.conv.weight.detach()), "b": self.conv.bias.detach().float().contiguous()}

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        pk = self.packed()
        B, H, W, C = x.shape
        assert C == self.channels
        up = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=x.device)
        ops.upsample2x_cast(x, up)
        out, st = _conv_out_act(B, 2 * H, 2 * W, self.out_channels, x.device)
        ops.conv3x3(up, pk["w"], out, bias=pk["b"], gn_stats=st.buf if st else None)
        return Act(out, st)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return _nchw(self._run(_nhwc(x)).t)


class Downsample(PackedModule):
    """openaimodel.py:138-164: conv3x3 stride 2 pad 1."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or not use_conv or padding != 1:
            raise NotImplementedError("Downsample: only dims=2, use_conv=True, padding=1")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def _pack(self):
        return {"w": pack_conv3x3(self.op.weight.detach()), "b": self.op.bias.detach().float().contiguous()}

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        pk = self.packed()
        B, H, W, C = x.shape
        assert C == self.channels
        xb = ops.cast_bf16(x)
        out, st = _conv_out_act(B, H // 2, W // 2, self.out_channels, x.device)
        ops.conv3x3(xb, pk["w"], out, stride=2, bias=pk["b"], gn_stats=st.buf if st else None)
        return Act(out, st)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return _nchw(self._run(_nhwc(x)).t)


class ResBlock(TimestepBlock, PackedModule):
    """openaimodel.py:167-279.  GN+SiLU -> conv3x3 (+bias +time-emb) -> GN+SiLU -> conv3x3 (+bias +skip)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or dims != 2 or use_conv:
            raise NotImplementedError("ResBlock: scale-shift norm / up / down / use_conv are not used by SD-1.5")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.updown = False
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def _pack(self):
        f = lambda t: t.detach().float().contiguous()
        pk = {"gn1_w": f(self.in_layers[0].weight), "gn1_b": f(self.in_layers[0].bias),
              "w1": pack_conv3x3(self.in_layers[2].weight.detach()), "b1": f(self.in_layers[2].bias),
              "we": f(self.emb_layers[1].weight), "be": f(self.emb_layers[1].bias),
              "gn2_w": f(self.out_layers[0].weight), "gn2_b": f(self.out_layers[0].bias),
              "w2": pack_conv3x3(self.out_layers[3].weight.detach()), "b2": f(self.out_layers[3].bias),
              "eps1": float(self.in_layers[0].eps), "eps2": float(self.out_layers[0].eps)}
        if isinstance(self.skip_connection, nn.Conv2d):
            pk["ws"] = pack_conv1x1(self.skip_connection.weight.detach())
            pk["bs"] = f(self.skip_connection.bias)
        return pk

    def _run(self, parts, emb_rows: torch.Tensor) -> "Act":
        """parts: (x,) or (h, skip) Acts whose channel concat is the block input (openaimodel.py:1019 torch.cat is
        never materialised in fp32); emb_rows fp32 [B, Cout] = emb_layers(emb) (may be a column slice of the
        UNet-level batched projection)."""
        pk = self.packed()
        a0 = parts[0]
        a1 = parts[1] if len(parts) > 1 else None
        x0 = a0.t
        x1 = a1.t if a1 is not None else None
        B, H, W, C0 = x0.shape
        Cin = C0 + (x1.shape[-1] if x1 is not None else 0)
        assert Cin == self.channels, (Cin, self.channels)
        Cout = self.out_channels
        dev = x0.device
        has_skip_conv = "ws" in pk
        y = torch.empty(B, H, W, Cin, dtype=torch.bfloat16, device=dev)
        raw = torch.empty(B, H, W, Cin, dtype=torch.bfloat16, device=dev) if has_skip_conv else None
        ops.groupnorm_apply(x0, a0.stats(), pk["gn1_w"], pk["gn1_b"], pk["eps1"], True, y, x1=x1,
                            st1=a1.stats() if a1 is not None else None, raw=raw)                   # :205-207
        h1, st1 = _conv_out_act(B, H, W, Cout, dev)
        ops.conv3x3(y, pk["w1"], h1, bias=pk["b1"], rowbias=emb_rows, gn_stats=st1.buf if st1 else None)  # :208,:277
        y2 = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(h1, st1 if st1 else ops.groupnorm_stats(h1), pk["gn2_w"], pk["gn2_b"], pk["eps2"], True,
                            y2)                                                                    # :229-231
        if has_skip_conv:
            res = torch.empty(B, H, W, Cout, dtype=torch.float32, device=dev)
            ops.gemm(raw.reshape(B * H * W, Cin), pk["ws"], res, bias=pk["bs"])                    # :245
        else:
            assert x1 is None
            res = x0
        out, sto = _conv_out_act(B, H, W, Cout, dev)
        ops.conv3x3(y2, pk["w2"], out, bias=pk["b2"], residual=res, gn_stats=sto.buf if sto else None)  # :234,:279
        return Act(out, sto)

    def emb_proj(self, emb: torch.Tensor) -> torch.Tensor:
        pk = self.packed()
        out = torch.empty(emb.shape[0], self.out_channels, dtype=torch.float32, device=emb.device)
        return ops.linear_small(emb.float().contiguous(), pk["we"], pk["be"], out, silu_in=True)  # :222-228,:268

    def forward(self, x, emb):
        """Reference signature: x [B,C,H,W], emb [B, emb_channels] -> [B,Cout,H,W]."""
        return _nchw(self._run((Act(_nhwc(x)),), self.emb_proj(emb)).t)

    _forward = forward


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """openaimodel.py:76-92."""

    def forward(self, x, emb, context=None, mask=None):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context, mask=mask)
            else:
                x = layer(x)
        return x

    def _run(self, parts, emb_rows_of, context, mask):
        """NHWC fast path.  parts: tuple of Acts (channel concat = input) -> Act."""
        x = parts
        for layer in self:
            if isinstance(layer, ResBlock):
                x = (layer._run(x, emb_rows_of(layer)),)
            elif isinstance(layer, SpatialTransformer):
                x = (layer._run_act(x[0], context, mask),)
            elif isinstance(layer, (Upsample, Downsample)):
                x = (layer._run(x[0].t),)
            elif isinstance(layer, ConvIn):
                x = (Act(layer._run(x[0].t)),)
            else:
                raise NotImplementedError(type(layer))
        return x[0]


class ConvIn(nn.Conv2d):
    """input_blocks[0][0]: conv_nd(2, 4, 320, 3, padding=1) (openaimodel.py:527-533): reads the public NCHW
    latent and writes the internal NHWC fp32 stream."""

    def _run(self, x_nchw: torch.Tensor) -> torch.Tensor:
        B, C, H, W = x_nchw.shape
        out = torch.empty(B, H, W, self.out_channels, dtype=torch.float32, device=x_nchw.device)
        return ops.conv_in(x_nchw.float().contiguous(), self.weight.detach().float().contiguous(),
                           self.bias.detach().float().contiguous(), out)

    def forward(self, x):
        return _nchw(self._run(x))


class UNetModel(PackedModule):
    """openaimodel.py:417-1052."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "context_dim is required with use_spatial_transformer"
        if context_dim is not None:
            assert use_spatial_transformer, "use_spatial_transformer is required with context_dim"
            if not isinstance(context_dim, int):
                context_dim = list(context_dim)
        if not use_spatial_transformer or dims != 2 or num_classes is not None or resblock_updown \
                or use_scale_shift_norm or n_embed is not None or not conv_resample or in_channels != 4:
            raise NotImplementedError("UNetModel: only the SD-1.5 / AdaFace configuration is implemented")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        if num_head_channels == -1:
            assert num_heads != -1, "Either num_heads or num_head_channels has to be set"

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.predict_codebook_ids = False
        self.debug_attn = False
        self.backup_vars = {"use_conv_attn_kernel_size:layerwise": [-1] * 16, "save_attn_vars": False,
                            "is_training": True}

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))

        def make_st(ch, nh):
            if num_head_channels == -1:
                dim_head, heads = ch // nh, nh
            else:
                heads, dim_head = ch // num_head_channels, num_head_channels
            if legacy:
                dim_head = ch // heads
            return SpatialTransformer(ch, heads, dim_head, depth=transformer_depth, context_dim=context_dim)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(ConvIn(in_channels, model_channels, 3, padding=1))])
        input_block_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, dims=dims,
                                   use_checkpoint=use_checkpoint)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(make_st(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims,
                                                                            out_channels=ch)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint),
            make_st(ch, num_heads),
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = input_block_chans.pop()
                layers = [ResBlock(ch + ich, time_embed_dim, dropout, out_channels=model_channels * mult, dims=dims,
                                   use_checkpoint=use_checkpoint)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(make_st(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))
        self._ctx_cache = []

    # ------------------------------------------------------------------ flags (openaimodel.py:723-824)
    def _layer_modules(self):
        return list(self.input_blocks) + [self.middle_block] + list(self.output_blocks)

    def set_cross_attn_flags(self, ca_flag_dict=None, ca_layer_indices=None, trans_flag_dict=None,
                             trans_layer_indices=None):
        if ca_flag_dict is None and trans_flag_dict is None:
            return None, None
        if ca_layer_indices is None:
            ca_layer_indices = ALL_CA_LAYER_INDICES
        if trans_layer_indices is None:
            trans_layer_indices = ALL_CA_LAYER_INDICES

        def apply(flag_dict, layer_indices, on_attn2):
            if flag_dict is None or len(layer_indices) == 0:
                return None
            old = {}
            for k, v in flag_dict.items():
                old[k] = self.backup_vars[k]
                self.backup_vars[k] = v
                is_arr = is_dict = False
                if k.endswith(":layerwise"):
                    k = k[:-len(":layerwise")]
                    is_arr = v is not None
                if k.endswith(":layerwise-dict"):
                    k = k[:-len(":layerwise-dict")]
                    is_dict = v is not None
                for layer_idx, module in enumerate(self._layer_modules()):
                    if layer_idx in layer_indices:
                        v2 = extract_layerwise_value(v, L2CA[layer_idx], is_arr, is_dict)
                        tb = module[1].transformer_blocks[0]
                        (tb.attn2 if on_attn2 else tb).__dict__[k] = v2
            return old

        return apply(ca_flag_dict, ca_layer_indices, True), apply(trans_flag_dict, trans_layer_indices, False)

    # ------------------------------------------------------------------ packing
    def res_blocks(self):
        return [m for m in self.modules() if isinstance(m, ResBlock)]

    def _param_key(self):
        # only the parameters this module packs itself (the sub-modules key their own packs)
        srcs = [*self.time_embed.parameters(), *self.out.parameters()]
        for rb in self.res_blocks():
            srcs += [rb.emb_layers[1].weight, rb.emb_layers[1].bias]
        return tuple((p.data_ptr(), p._version) for p in srcs)

    def _pack(self):
        f = lambda t: t.detach().float().contiguous()
        rbs = self.res_blocks()
        offs, o = {}, 0
        for rb in rbs:
            offs[id(rb)] = (o, rb.out_channels)
            o += rb.out_channels
        return {"te0_w": f(self.time_embed[0].weight), "te0_b": f(self.time_embed[0].bias),
                "te2_w": f(self.time_embed[2].weight), "te2_b": f(self.time_embed[2].bias),
                "emb_w": torch.cat([f(rb.emb_layers[1].weight) for rb in rbs], 0).contiguous(),
                "emb_b": torch.cat([f(rb.emb_layers[1].bias) for rb in rbs], 0).contiguous(),
                "emb_offs": offs, "emb_total": o,
                "out_gn_w": f(self.out[0].weight), "out_gn_b": f(self.out[0].bias), "out_eps": float(self.out[0].eps),
                "out_w": self._pad_out_conv(self.out[2].weight.detach()), "out_b": self._pad_out_bias(self.out[2].bias.detach())}

    @staticmethod
    def _pad_out_conv(w_oihw):
        """UNetModel.out[-1] (openaimodel.py:696): Cout 4 -> 8 zero rows so the conv runs on the tensor cores."""
        co = w_oihw.shape[0]
        wp = torch.zeros((co + 7) // 8 * 8, *w_oihw.shape[1:], dtype=w_oihw.dtype, device=w_oihw.device)
        wp[:co] = w_oihw
        return pack_conv3x3(wp)

    @staticmethod
    def _pad_out_bias(b):
        co = b.shape[0]
        bp = torch.zeros((co + 7) // 8 * 8, dtype=torch.float32, device=b.device)
        bp[:co] = b.float()
        return bp

    def invalidate_packed(self):
        super().invalidate_packed()
        self.__dict__["_ctx_cache"] = []

    def prepare(self):
        """Packs every sub-module now (otherwise done lazily on first use)."""
        for m in self.modules():
            if isinstance(m, PackedModule):
                m.packed()
        return self

    # ------------------------------------------------------------------ context K/V cache
    def _ca_modules(self):
        out = {}
        for layer_idx, module in enumerate(self._layer_modules()):
            if layer_idx in L2CA:
                out[layer_idx] = module[1].transformer_blocks[0].attn2
        return out

    def context_kv(self, context: torch.Tensor, B: int, iter_type: str = "normal_recon"):
        """Projects the layerwise context [16*B, Nt, 768] (openaimodel.py:866) through to_k / to_v of the 16
        cross-attention layers ONCE; cached by tensor identity + version, so all DDIM steps and both CFG
        branches reuse it.  Returns {layer_idx: ContextKV}.  When the SAME tensor object is modified in place
        (new prompt copied into a static buffer) the projections are refreshed into the same K / V^T buffers,
        so CUDA graphs captured over them remain valid."""
        cache = self.__dict__.setdefault("_ctx_cache", [])
        stale = None
        for ent in cache:
            src, ptr, ver, it, kvs = ent
            if src is context and ptr == context.data_ptr() and it == iter_type:
                if ver == context._version:
                    return kvs
                stale = ent
        ctx = context.reshape(B, 16, -1, context.shape[-1]).permute(1, 0, 2, 3)
        if stale is not None:
            for layer_idx, attn2 in self._ca_modules().items():
                c = ctx[L2CA[layer_idx]].float().contiguous()
                v_c, k_c = (t.contiguous() for t in c.chunk(2, dim=1)) if iter_type == "mix_hijk" else (c, c)
                attn2.project_context(k_c, v_c, out=stale[4][layer_idx])
            cache[cache.index(stale)] = (context, context.data_ptr(), context._version, iter_type, stale[4])
            return stale[4]
        kvs = {}
        for layer_idx, attn2 in self._ca_modules().items():
            c = ctx[L2CA[layer_idx]].float().contiguous()
            if iter_type == "mix_hijk":                                                   # :885-892
                v_c, k_c = (t.contiguous() for t in c.chunk(2, dim=1))
            else:
                v_c = k_c = c
            kv = attn2.project_context(k_c, v_c)
            attn2.__dict__["_kv_cache"] = None  # the UNet-level cache owns it
            kvs[layer_idx] = kv
        cache.append((context, context.data_ptr(), context._version, iter_type, kvs))
        if len(cache) > 4:
            cache.pop(0)
        return kvs

    def compel_context_kv(self, context, B, iter_type, empty_context, prob, level_or_range, is_training):
        """Compel-style CFG on the context (openaimodel.py:898-916 + ldm/util.py:1823-1854): per cross-attention layer, with
        probability `prob`, ctx <- (ctx - empty) * 1.1**level + empty on the instances the batch mask selects (inference:
        the first half of the batch; training: all, or - at 50 % - the second half only).  The draws come from Python's
        global `random` in exactly the reference's order (per layer: [training: mask coin,] gate, level, and one more
        gate draw per element of the (v, k) tuple), so a seeded run reproduces the reference.  Nothing is cached: the
        K / V projections are recomputed from the re-weighted context on every forward."""
        import random
        ctx = context.reshape(B, 16, -1, context.shape[-1]).permute(1, 0, 2, 3)
        kvs = {}
        for layer_idx, attn2 in self._ca_modules().items():          # ascending layer order, as the forward visits them
            c = ctx[L2CA[layer_idx]].float().contiguous()
            v_c, k_c = (t.contiguous() for t in c.chunk(2, dim=1)) if iter_type == "mix_hijk" else (c, c)
            mask = torch.ones(B, dtype=torch.float32, device=c.device)
            if is_training:
                if random.random() < 0.5:
                    mask[:B // 2] = 0
            else:
                mask[B // 2:] = 0
            if not (empty_context is None or level_or_range is None or random.random() > prob):
                level = random.uniform(*level_or_range) if isinstance(level_or_range, (list, tuple)) else level_or_range
                w = 1.1 ** level
                e = empty_context.to(c.device).float()
                m = mask.reshape(-1, 1, 1)

                def reweight(t):
                    random.random()                                    # the recursive call's own (always passing) gate
                    t2 = (t - e) * w + e
                    return (t2 * m + t * (1 - m)).contiguous()

                v_c, k_c = reweight(v_c), reweight(k_c)                 # tuple order (v, k): ldm/util.py:1837
            kv = attn2.project_context(k_c, v_c)
            attn2.__dict__["_kv_cache"] = None
            kvs[layer_idx] = kv
        return kvs

    # ------------------------------------------------------------------ forward
    def time_embedding(self, timesteps: torch.Tensor):
        """-> (emb [B, 4*mc] fp32, emb_rows [B, sum Cout] = every ResBlock's emb_layers(emb))."""
        pk = self.packed()
        B = timesteps.shape[0]
        dev = timesteps.device
        t = timesteps.float().contiguous()
        t_emb = ops.timestep_embedding(t, self.model_channels)                             # :846
        e1 = torch.empty(B, pk["te0_w"].shape[0], dtype=torch.float32, device=dev)
        ops.linear_small(t_emb, pk["te0_w"], pk["te0_b"], e1, silu_out=True)              # :518-521
        emb = torch.empty(B, pk["te2_w"].shape[0], dtype=torch.float32, device=dev)
        # emb is only ever consumed through the SiLU that opens every ResBlock.emb_layers (:222-228), so the
        # activation is applied once here instead of once per output feature of the 22 projections
        ops.linear_small(e1, pk["te2_w"], pk["te2_b"], emb, silu_out=True)                # :847 + SiLU of :223
        rows = torch.empty(B, pk["emb_total"], dtype=torch.float32, device=dev)
        ops.linear_small(emb, pk["emb_w"], pk["emb_b"], rows)                             # :224-228 for all blocks
        return emb, rows

    def forward(self, x, timesteps=None, context=None, y=None, context_in=None, extra_info=None, **kwargs):
        """Reference signature (openaimodel.py:827).  x [B,4,H,W] fp32, timesteps [B], context [16*B,Nt,768]."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise RuntimeError("UNetModel.forward: x must be a CUDA tensor (no CPU fallback)")
        ei = extra_info if extra_info is not None else {}
        use_layerwise_context = ei.get("use_layerwise_context", False)
        iter_type = ei.get("iter_type", "normal_recon")
        is_training = ei.get("is_training", True)
        capture_distill_attn = ei.get("capture_distill_attn", False)
        use_conv_attn_kernel_size = ei.get("use_conv_attn_kernel_size", None)
        placeholder2indices = ei.get("placeholder2indices", None)
        img_mask = ei.get("img_mask", None)
        apply_compel_cfg_prob = ei.get("apply_compel_cfg_prob", 0)
        debug_attn = ei.get("debug_attn", self.debug_attn)
        if extra_info is None:
            raise TypeError("extra_info must be a dict (the reference writes 'ca_layers_activations' into it, "
                            "openaimodel.py:1035)")
        if not use_layerwise_context:
            # the reference's non-layerwise branch returns a 3-tuple that CrossAttention cannot unpack
            # (openaimodel.py:872 vs attention.py:186): only the layerwise path is live.
            raise ValueError("extra_info['use_layerwise_context'] must be True")
        if debug_attn:
            raise NotImplementedError("debug_attn (the reference drops into breakpoint(), openaimodel.py:1037-1038)")
        B = x.shape[0]
        if apply_compel_cfg_prob > 0:
            kvs = self.compel_context_kv(context, B, iter_type, ei.get("empty_context", None), apply_compel_cfg_prob,
                                         ei.get("compel_cfg_weight_level_range", None), is_training)
        else:
            kvs = self.context_kv(context, B, iter_type)

        # flag side channel (openaimodel.py:922-945, restored at :1041-1045)
        sizes = np.ones(16, dtype=int) * use_conv_attn_kernel_size
        if use_conv_attn_kernel_size > 0:
            sizes[6:11] = 1
        old_ca_flags, _ = self.set_cross_attn_flags(
            ca_flag_dict={"use_conv_attn_kernel_size:layerwise": sizes, "is_training": is_training},
            ca_layer_indices=None)
        conv_on = use_conv_attn_kernel_size > 0 and placeholder2indices is not None
        for kv in kvs.values():                         # what get_layer_context returns next to the tensors (:920)
            kv.placeholder2indices = placeholder2indices if conv_on else None
        distill_layer_indices, distill_old = [], None
        if capture_distill_attn:                                                          # :947-952
            distill_layer_indices = [7, 8, 12, 16, 17, 18, 19, 20, 21, 22, 23, 24]
            distill_old, _ = self.set_cross_attn_flags(ca_flag_dict={"save_attn_vars": True},
                                                       ca_layer_indices=distill_layer_indices)
        acts = {}
        try:
            out = self._forward_nhwc(x, timesteps, kvs, img_mask, distill_layer_indices, acts, ei.get("emb_rows", None))
        finally:
            if distill_old is not None:
                self.set_cross_attn_flags(ca_flag_dict=distill_old, ca_layer_indices=distill_layer_indices)
            self.set_cross_attn_flags(ca_flag_dict=old_ca_flags, ca_layer_indices=None)
            for kv in kvs.values():
                kv.placeholder2indices = None
        extra_info["ca_layers_activations"] = {key: {li: acts[li][key] for li in acts}
                                               for key in ("outfeat", "attn", "attnscore", "q")}      # :1031-1035
        return out

    def _forward_nhwc(self, x, timesteps, kvs, img_mask, capture_layers=(), acts=None, emb_rows=None):
        pk = self.packed()
        if emb_rows is not None:
            # precomputed by the caller from the SAME timesteps (graph_sampler: one table per sample() call, the 113 MB of
            # fp32 emb_layers weights are read once instead of once per step); bit-identical to computing them here
            if emb_rows.shape != (x.shape[0], pk["emb_total"]) or emb_rows.dtype != torch.float32:
                raise ValueError("extra_info['emb_rows'] must be fp32 [batch, sum of ResBlock channels]")
            rows = emb_rows
        else:
            _, rows = self.time_embedding(timesteps)
        offs = pk["emb_offs"]

        def emb_rows_of(rb):
            o, n = offs[id(rb)]
            return rows[:, o:o + n]

        hs = []
        h = Act(x)
        layer_idx = 0
        def capture(module, h, layer_idx):                                                # :984-988, 996-1000, 1023-1027
            if layer_idx in capture_layers:
                attn2 = module[1].transformer_blocks[0].attn2
                acts[layer_idx] = dict(attn2.cached_activations)
                acts[layer_idx]["outfeat"] = h.t.permute(0, 3, 1, 2)       # NCHW view of the NHWC residual stream
                attn2.cached_activations = None

        for module in self.input_blocks:                                                   # :977-990
            h = module._run((h,), emb_rows_of, kvs.get(layer_idx), img_mask)
            hs.append(h)
            capture(module, h, layer_idx)
            layer_idx += 1
        h = self.middle_block._run((h,), emb_rows_of, kvs.get(layer_idx), img_mask)        # :995
        capture(self.middle_block, h, layer_idx)
        layer_idx += 1
        for module in self.output_blocks:                                                  # :1016-1029
            h = module._run((h, hs.pop()), emb_rows_of, kvs.get(layer_idx), img_mask)
            capture(module, h, layer_idx)
            layer_idx += 1
        B, H, W, C = h.t.shape
        dev = h.t.device
        y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(h.t, h.stats(), pk["out_gn_w"], pk["out_gn_b"], pk["out_eps"], True, y)   # :693-695
        o8 = torch.empty(B, H, W, pk["out_w"].shape[0], dtype=torch.float32, device=dev)
        ops.conv3x3(y, pk["out_w"], o8, bias=pk["out_b"], bn=64)                                       # :696,:1052
        out = torch.empty(B, self.out_channels, H, W, dtype=torch.float32, device=dev)
        return ops.nhwc_to_nchw(o8, out)
