"""Host mirror of ldm/modules/attention.py (reference: askerlee/adaprompt) on the B200 C ABI.

Same classes, constructor signatures, parameter names (state_dict keys) and forward signatures as the
reference: GEGLU :32, FeedForward :42, CrossAttention :147, BasicTransformerBlock :260,
SpatialTransformer :287.  The torch.nn layers below only *hold parameters* (so SD-1.5 checkpoints load
unchanged); no torch arithmetic runs in any forward - each `_run` drives libadaface_b200.so:

    LN -> [Q|K] GEMM + V^T GEMM -> flash attention (tcgen05) -> to_out GEMM (+bias +residual)
    LN -> Q GEMM -> flash attention over the cached 77-token K / V^T -> to_out GEMM (+bias +residual)
    LN -> GEGLU GEMM -> FF2 GEMM (+bias +residual)

Internal layout: tokens x channels (NHWC), fp32 residual stream, bf16 GEMM operands.
The public forwards accept / return the reference layouts ([B,N,C] tokens, NCHW images).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .packing import head_pad, pack_conv1x1, pack_geglu, pack_qk

CTX_PAD = 8  # cached context keys are padded to a multiple of 8 rows (16-byte TMA strides for V^T)


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else (d() if callable(d) else d)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class PackedModule(nn.Module):
    """Lazily repacks fp32 reference-layout parameters into kernel layouts.  The packs (forward `_pk`, the transposed
    dgrad packs `_pk_bwd` of train.py and the frozen-layer training packs `_train_pk` of train_cond.py) are keyed on the
    (storage pointer, version counter) of every source parameter: an optimizer step, an in-place copy, a re-bound
    `.data`, `.to()` or `load_state_dict` all change the key and the next use repacks."""

    PACK_EPOCH = 0   # bumped on every repack of any module: graph_sampler re-captures when it changes

    def _pack(self) -> dict:
        raise NotImplementedError

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def packed(self) -> dict:
        pk = self.__dict__.get("_pk")
        key = self._param_key()
        if pk is None or self.__dict__.get("_pk_key") != key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError(f"{type(self).__name__}: parameters are on {dev}; the B200 path has no CPU "
                                   "fallback - move the module to a CUDA device first")
            if pk is not None:
                self.invalidate_packed()
            with torch.no_grad():
                pk = self._pack()
            self.__dict__["_pk"] = pk
            self.__dict__["_pk_key"] = key
            PackedModule.PACK_EPOCH += 1
        return pk

    def invalidate_packed(self):
        self.__dict__["_pk"] = None
        self.__dict__["_pk_key"] = None
        self.__dict__.pop("_pk_bwd", None)
        self.__dict__.pop("_train_pk", None)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate_packed()
        return r

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self.invalidate_packed()


def invalidate_all(module: nn.Module):
    for m in module.modules():
        if isinstance(m, PackedModule):
            m.invalidate_packed()


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(PackedModule):
    """attention.py:42-59 (glu=True path).  `x * gelu(gate)` is fused into the first GEMM's epilogue."""

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("only the GEGLU feed-forward of the SD-1.5 config is implemented")
        inner_dim = int(dim * mult)
        dim_out = default(dim_out, dim)
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def _pack(self):
        w1, b1 = pack_geglu(self.net[0].proj.weight.detach().float(), self.net[0].proj.bias.detach().float())
        return {"w1": w1.to(torch.bfloat16).contiguous(), "b1": b1.contiguous(),
                "w2": self.net[2].weight.detach().to(torch.bfloat16).contiguous(),
                "b2": self.net[2].bias.detach().float().contiguous()}

    def _run(self, x_ln: torch.Tensor, residual: Optional[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
        """x_ln bf16 [T, C]; out [T, C] (fp32 or bf16) = FF(x_ln) + residual."""
        pk = self.packed()
        T = x_ln.shape[0]
        hid = torch.empty(T, pk["w2"].shape[1], dtype=torch.bfloat16, device=x_ln.device)
        ops.gemm(x_ln, pk["w1"], hid, bias=pk["b1"], geglu=True)
        ops.gemm(hid, pk["w2"], out, bias=pk["b2"], residual=residual)
        return out

    def forward(self, x):
        shp = x.shape
        xb = ops.cast_bf16(x.reshape(-1, shp[-1]).float().contiguous())
        out = torch.empty(xb.shape[0], self.net[2].weight.shape[0], dtype=torch.float32, device=x.device)
        return self._run(xb, None, out).reshape(*shp[:-1], -1)


class ContextKV:
    """Projected cross-attention keys / values of one context: computed once per prompt and reused by
    every DDIM step and both CFG branches (the context is step invariant, ddim.py:243-247)."""

    __slots__ = ("k", "vt", "nk", "nk_pad", "B", "src", "placeholder2indices")

    def __init__(self, k, vt, nk, nk_pad, B, src):
        self.k, self.vt, self.nk, self.nk_pad, self.B, self.src = k, vt, nk, nk_pad, B, src
        # what the reference's callable context returns next to the tensors (openaimodel.py:920): set by UNetModel.forward
        # when conv attention is on, read by CrossAttention._run
        self.placeholder2indices = None


class CrossAttention(PackedModule):
    """attention.py:147-257.  heads must be 8 with head dim 40 / 80 / 160 (SD-1.5)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.is_self = context_dim is None
        context_dim = default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        # run-time flags of the reference (attention.py:166-170); set through UNetModel.set_cross_attn_flags
        self.save_attn_vars = False
        self.cached_activations = None
        self.use_conv_attn_kernel_size = -1
        self.infeat_size = None
        self.is_training = True
        self._kv_cache = None

    # ------------------------------------------------------------------ packing
    def _pack(self):
        h = self.heads
        wq, wk, wv = (m.weight.detach().float() for m in (self.to_q, self.to_k, self.to_v))
        pk = {"wo": self.to_out[0].weight.detach().to(torch.bfloat16).contiguous(),
              "bo": self.to_out[0].bias.detach().float().contiguous(),
              "wv": wv.to(torch.bfloat16).contiguous(),
              "wq": pack_qk(wq, None, h), "wk": pack_qk(wk, None, h, fold_scale=False)}
        if wq.shape[1] == wk.shape[1]:
            pk["wqk"] = pack_qk(wq, wk, h)  # fused [Q|K] projection for self-attention
        return pk

    def invalidate_packed(self):
        super().invalidate_packed()
        self.__dict__["_kv_cache"] = None

    # ------------------------------------------------------------------ context K / V cache
    def project_context(self, k_ctx: torch.Tensor, v_ctx: Optional[torch.Tensor] = None,
                        out: Optional[ContextKV] = None) -> ContextKV:
        """K = to_k(k_ctx) [B, nk_pad, 8*dp], V^T = (to_v(v_ctx))^T [C, B*nk_pad] (attention.py:195-196).
        `out`: refresh an existing ContextKV in place (same buffers -> captured CUDA graphs stay valid)."""
        v_ctx = k_ctx if v_ctx is None else v_ctx
        c = self.__dict__.get("_kv_cache")
        if c is not None and out is None:
            (ks, kp, kver), (vs, vp, vver) = c.src
            if ks is k_ctx and vs is v_ctx and kp == k_ctx.data_ptr() and kver == k_ctx._version \
                    and vp == v_ctx.data_ptr() and vver == v_ctx._version:
                return c
        pk = self.packed()
        B, nk, cd = k_ctx.shape
        if v_ctx.shape != k_ctx.shape:
            raise ValueError("k / v contexts must have the same shape")
        nk_pad = (nk + CTX_PAD - 1) // CTX_PAD * CTX_PAD
        dev = k_ctx.device

        def padded_bf16(ctx):
            buf = torch.zeros(B, nk_pad, cd, dtype=torch.float32, device=dev)
            buf[:, :nk] = ctx
            return ops.cast_bf16(buf).reshape(B * nk_pad, cd)

        kb = padded_bf16(k_ctx.float())
        vb = kb if v_ctx is k_ctx else padded_bf16(v_ctx.float())
        if out is not None:
            if (out.B, out.nk, out.nk_pad) != (B, nk, nk_pad):
                raise ValueError("project_context(out=...): shape changed")
            k, vt = out.k, out.vt
        else:
            k = torch.empty(B * nk_pad, pk["wk"].shape[0], dtype=torch.bfloat16, device=dev)
            vt = torch.zeros(pk["wv"].shape[0], max(64, B * nk_pad), dtype=torch.bfloat16, device=dev)
        ops.gemm(kb, pk["wk"], k)
        ops.gemm(pk["wv"], vb, vt, bn=128, ldo=vt.shape[1])
        src = ((k_ctx, k_ctx.data_ptr(), k_ctx._version), (v_ctx, v_ctx.data_ptr(), v_ctx._version))
        if out is not None:
            out.src = src
            return out
        kv = ContextKV(k, vt, nk, nk_pad, B, src)
        self.__dict__["_kv_cache"] = kv
        return kv

    # ------------------------------------------------------------------ kernels
    def _run(self, x_ln: torch.Tensor, B: int, N: int, residual: Optional[torch.Tensor], out: torch.Tensor,
             kv: Optional[ContextKV] = None, key_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_ln bf16 [B*N, C] (already layer-normed).  kv None -> self-attention.  out = attn(x) + residual."""
        pk = self.packed()
        h, d = self.heads, self.dim_head
        dp = head_pad(d)
        dev = x_ln.device
        T = B * N
        C = h * d
        if kv is None:
            qk = torch.empty(T, 2 * h * dp, dtype=torch.bfloat16, device=dev)
            ops.gemm(x_ln, pk["wqk"], qk)
            ldvt = max(64, (T + 7) // 8 * 8)  # V^T row pitch: TMA wants >= one 128-byte swizzle row
            # pad columns must be finite (masked keys still enter P.V as 0 * v): zero-fill the rare padded case
            vt = (torch.zeros if ldvt != T else torch.empty)(C, ldvt, dtype=torch.bfloat16, device=dev)
            ops.gemm(pk["wv"], x_ln, vt, bn=256 if T % 256 == 0 else 128, ldo=ldvt)
            q, k, ldq, ldk = qk, qk[:, h * dp:], 2 * h * dp, 2 * h * dp
            nk, kv_stride = N, N
        else:
            if kv.B != B:
                raise ValueError(f"context batch {kv.B} != query batch {B}")
            q = torch.empty(T, h * dp, dtype=torch.bfloat16, device=dev)
            ops.gemm(x_ln, pk["wq"], q)
            k, vt, ldq, ldk = kv.k, kv.vt, h * dp, h * dp
            nk, kv_stride, ldvt = kv.nk, kv.nk_pad, kv.vt.shape[1]
        p2i = kv.placeholder2indices if kv is not None else None
        conv = kv is not None and p2i is not None and self.use_conv_attn_kernel_size is not None \
            and self.use_conv_attn_kernel_size > 1                      # kernel size 1 leaves the scores alone (ldm/util.py:704)
        if kv is not None and (self.save_attn_vars or conv):
            o = self._explicit_cross_attention(q, k, vt, B, N, nk, kv_stride, ldq, ldk, ldvt, p2i if conv else None)
        else:
            if self.save_attn_vars:
                raise NotImplementedError("save_attn_vars on a self-attention layer (the reference only sets it on attn2)")
            o = torch.empty(T, C, dtype=torch.bfloat16, device=dev)
            ops.attention(q, k, vt, o, B=B, heads=h, Nq=N, Nk=nk, d=d, ldq=ldq, ldk=ldk, ldvt=ldvt, kv_stride=kv_stride,
                          key_mask=key_mask)
        ops.gemm(o, pk["wo"], out, bias=pk["bo"], residual=residual)
        return out

    def _explicit_cross_attention(self, q, k, vt, B, N, nk, kv_stride, ldq, ldk, ldvt, placeholder2indices):
        """The score-materialising path (xattn_explicit.cu) for save_attn_vars (attention.py:245-255) and conv attention
        (:208-216 -> ldm/util.py:700-878).  Returns o bf16 [B*N, C]; fills self.cached_activations when saving."""
        h, d = self.heads, self.dim_head
        dev = q.device
        common = dict(B=B, heads=h, N=N, nk=nk, d=d, ldq=ldq, ldk=ldk, ldvt=ldvt, kv_stride=kv_stride)
        override = cols = None
        if placeholder2indices is not None:
            ks = int(self.use_conv_attn_kernel_size)
            if self.infeat_size is None or self.infeat_size[0] * self.infeat_size[1] != N:
                raise ValueError("conv attention needs infeat_size = the (H, W) of the query map (set by SpatialTransformer)")
            Hf, Wf = self.infeat_size
            point = ops.xattn_explicit(q, k, vt, want_out=False, want_scores=True, **common)["attnscore"]
            ovs, cls = [], []
            for subj_string in placeholder2indices:                        # :209-216, one replacement per subject string
                indices_B, indices_N = placeholder2indices[subj_string]
                uniq = torch.unique(indices_B)
                BS = len(uniq)
                M = len(indices_N) // BS
                if ks * ks > M:
                    raise ValueError(f"{M} embeddings are not enough to cover a {ks}x{ks} kernel.")
                c = torch.full((B, ks * ks), -1, dtype=torch.int32, device=dev)
                c[uniq.to(dev).long()] = indices_N.reshape(BS, M)[:, :ks * ks].to(device=dev, dtype=torch.int32)
                ovs.append(ops.conv_attn_scores(point, c, B=B, heads=h, Hf=Hf, Wf=Wf, nk=nk, ks=ks))
                cls.append(c)
            override, cols = torch.cat(ovs, dim=-1).contiguous(), torch.cat(cls, dim=-1).contiguous()
        save = self.save_attn_vars
        r = ops.xattn_explicit(q, k, vt, override=override, ov_cols=cols, want_out=True, want_scores=save, want_attn=save,
                               want_q=save, **common)
        if save:
            self.cached_activations = {"q": r["q"], "attn": r["attn"], "attnscore": r["attnscore"]}
        return r["out"]

    def forward(self, x, context=None, mask=None):
        """Reference signature (attention.py:172): x [B,N,C]; context None | tensor | (v_ctx, k_ctx) | callable."""
        B, N, C = x.shape
        context_provided = exists(context)
        if callable(context):
            context, placeholder2indices = context()
        else:
            placeholder2indices = None
        kv = None
        if context_provided:
            if isinstance(context, (list, tuple)):
                v_context, k_context = context
            else:
                v_context = k_context = context
            kv = self.project_context(k_context, v_context)
            kv.placeholder2indices = placeholder2indices
        km = None
        if exists(mask):
            km = mask.reshape(B, -1).bool().to(torch.uint8).contiguous()
        xb = ops.cast_bf16(x.reshape(B * N, C).float().contiguous())
        out = torch.empty(B * N, self.to_out[0].weight.shape[0], dtype=torch.float32, device=x.device)
        return self._run(xb, B, N, None, out, kv, km).reshape(B, N, -1)


class BasicTransformerBlock(PackedModule):
    """attention.py:260-285."""

    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def _pack(self):
        return {n: (getattr(self, n).weight.detach().float().contiguous(),
                    getattr(self, n).bias.detach().float().contiguous(), float(getattr(self, n).eps))
                for n in ("norm1", "norm2", "norm3")}

    def _ln(self, name, x, T, C):
        w, b, eps = self.packed()[name]
        y = torch.empty(T, C, dtype=torch.bfloat16, device=x.device)
        return ops.layernorm(x, w, b, eps, y)

    def _run(self, x: torch.Tensor, B: int, N: int, kv: Optional[ContextKV], key_mask: Optional[torch.Tensor],
             out_dtype=torch.float32) -> torch.Tensor:
        """x fp32 [B*N, C] residual stream -> [B*N, C] (out_dtype)."""
        T, C = x.shape
        dev = x.device
        x1 = torch.empty(T, C, dtype=torch.float32, device=dev)
        self.attn1._run(self._ln("norm1", x, T, C), B, N, x, x1, None, key_mask)          # :277
        if kv is None:  # attention.py:266 "is self-attn if context is none"
            x2 = torch.empty(T, C, dtype=torch.float32, device=dev)
            self.attn2._run(self._ln("norm2", x1, T, C), B, N, x1, x2, None, None)
        else:
            x2 = torch.empty(T, C, dtype=torch.float32, device=dev)
            self.attn2._run(self._ln("norm2", x1, T, C), B, N, x1, x2, kv, None)            # :280-281
        x3 = torch.empty(T, C, dtype=out_dtype, device=dev)
        self.ff._run(self._ln("norm3", x2, T, C), x2, x3)                                  # :283
        return x3

    def resolve_context(self, context) -> Optional[ContextKV]:
        if context is None:
            return None
        if isinstance(context, ContextKV):
            return context
        if callable(context):
            context, _ = context()
        if isinstance(context, (list, tuple)):
            v_context, k_context = context
        else:
            v_context = k_context = context
        return self.attn2.project_context(k_context, v_context)

    def forward(self, x, context=None, mask=None):
        B, N, C = x.shape
        km = mask.reshape(B, -1).bool().to(torch.uint8).contiguous() if exists(mask) else None
        out = self._run(x.reshape(B * N, C).float().contiguous(), B, N, self.resolve_context(context), km)
        return out.reshape(B, N, C)

    _forward = forward


class SpatialTransformer(PackedModule):
    """attention.py:287-341 with depth = 1."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None):
        super().__init__()
        if depth != 1:
            raise NotImplementedError("transformer_depth != 1")
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim)
             for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))
        self.save_feat = False

    def _pack(self):
        return {"gn_w": self.norm.weight.detach().float().contiguous(),
                "gn_b": self.norm.bias.detach().float().contiguous(), "gn_eps": float(self.norm.eps),
                "w_in": pack_conv1x1(self.proj_in.weight.detach()), "b_in": self.proj_in.bias.detach().float().contiguous(),
                "w_out": pack_conv1x1(self.proj_out.weight.detach()),
                "b_out": self.proj_out.bias.detach().float().contiguous()}

    def _run(self, x: torch.Tensor, context, mask: Optional[torch.Tensor]) -> torch.Tensor:
        return self._run_act(x, context, mask).t

    def _run_act(self, xa, context, mask: Optional[torch.Tensor]):
        """xa: unet.Act or fp32 NHWC tensor [B,H,W,C] -> unet.Act (fp32 NHWC + the GroupNorm partial statistics the
        proj_out epilogue wrote).  mask: float/bool [B,1,H0,W0] image mask or None."""
        from .unet import Act
        if not isinstance(xa, Act):
            xa = Act(xa)
        x = xa.t
        pk = self.packed()
        B, H, W, C = x.shape
        T = B * H * W
        dev = x.device
        block = self.transformer_blocks[0]
        block.attn2.infeat_size = (H, W)                                                   # :330
        km = None
        if exists(mask):                                                                   # :332
            m2 = F.interpolate(mask.float(), size=(H, W), mode="nearest")
            km = m2.reshape(B, H * W).bool().to(torch.uint8).contiguous()
        xn = torch.empty(T, C, dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(x, xa.stats(), pk["gn_w"], pk["gn_b"], pk["gn_eps"], False, xn)  # :325
        t = torch.empty(T, pk["w_in"].shape[0], dtype=torch.float32, device=dev)
        ops.gemm(xn, pk["w_in"], t, bias=pk["b_in"])                                       # :326
        tb = block._run(t, B, H * W, block.resolve_context(context), km, out_dtype=torch.bfloat16)
        out = torch.empty(B, H, W, C, dtype=torch.float32, device=dev)
        st = ops.gn_stats_for_gemm(B, H * W, C, dev)
        ops.gemm(tb, pk["w_out"], out, bias=pk["b_out"], residual=x, gn_stats=st.buf if st else None)  # :340-341
        return Act(out, st)

    def forward(self, x, context=None, mask=None):
        """x NCHW [B,C,H,W] -> NCHW (reference layout)."""
        xh = x.float().permute(0, 2, 3, 1).contiguous()
        return self._run(xh, context, mask).permute(0, 3, 1, 2).contiguous()
