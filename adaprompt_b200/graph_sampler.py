"""CUDA-graph execution of the DDIM loop (used by DDIMSampler.ddim_sampling for "plain" calls).

One graph = one denoising step:  [x ; x] -> UNet -> fused CFG + DDIM update (in place on x) -> advance the
device-side step counter / timestep buffer.  Per-step scalars (guidance scale, alpha coefficients,
timestep) live in device tables indexed by the counter, so the same graph is replayed for all S steps.

The captured graph and every buffer it touches are kept in a per-sampler cache keyed by the call
geometry, so a new sample() call with the same shapes only
  1. copies the new conditioning into the static context buffer,
  2. asks the model to refresh the per-prompt cross-attention K / V^T projections IN PLACE
     (model.refresh_conditioning; 32 small GEMMs - the "KV cache warm-up"),
  3. rewrites the coefficient / timestep tables, and replays.
No capture, no allocation and no host synchronisation happen on that path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops
from .attention import PackedModule
from .diffusion_util import noise_like


_UNET_FLAGS = ("use_layerwise_context", "iter_type", "is_training", "capture_distill_attn", "use_conv_attn_kernel_size",
               "apply_compel_cfg_prob", "debug_attn")


def _extra_key(extra_info: Optional[dict]):
    """Hashable summary of what a captured graph bakes in: exactly the fields UNetModel.forward reads
    (openaimodel.py:849-859).  placeholder2indices / prompt_emb_mask (always present in the output of
    get_learned_conditioning) are only read by the conv-attention path, so they do not key the graph.
    None -> not cacheable (a per-call tensor would be baked into the graph by address)."""
    if extra_info is None:
        return ()
    if extra_info.get("img_mask", None) is not None:
        return None
    conv = extra_info.get("use_conv_attn_kernel_size", None)
    if conv is not None and conv > 0 and extra_info.get("placeholder2indices", None) is not None:
        return None
    items = []
    for k in _UNET_FLAGS:
        v = extra_info.get(k, None)
        if torch.is_tensor(v) or isinstance(v, (dict, list)):
            return None
        items.append((k, v))
    return tuple(items)


class _StepGraph:
    """Static buffers + captured graph for one (batch, latent shape, cfg on/off, flags) geometry."""

    def __init__(self, sampler, b, lat_shape, cfg_on, cond, uncond, max_steps, device, with_noise):
        self.b, self.cfg_on, self.max_steps = b, cfg_on, max_steps
        c_c, c_in_c, extra_info = cond
        nb = 2 * b if cfg_on else b
        self.x = torch.empty((b,) + tuple(lat_shape), dtype=torch.float32, device=device)
        self.pred = torch.empty_like(self.x)
        self.noise = torch.empty_like(self.x) if with_noise else None
        self.x_in = torch.empty((nb,) + tuple(lat_shape), dtype=torch.float32, device=device)
        self.t_buf = torch.zeros(nb, dtype=torch.float32, device=device)
        self.step_idx = torch.zeros(1, dtype=torch.int32, device=device)
        self.coef_table = torch.zeros(max_steps, 8, dtype=torch.float32, device=device)
        self.t_table = torch.zeros(max_steps, dtype=torch.float32, device=device)
        n_c = c_c.shape[0]
        rows = n_c + (uncond[0].shape[0] if cfg_on else 0)
        self.ctx = torch.empty((rows,) + tuple(c_c.shape[1:]), dtype=torch.float32, device=device)
        self.n_c = n_c
        self.extra_info = dict(extra_info) if extra_info is not None else None
        prompts = list(c_in_c) + (list(uncond[1]) if cfg_on else [])
        self.c2 = (self.ctx, prompts, self.extra_info)
        # time-embedding rows of every step (model.time_embedding_rows): filled per run(), indexed in the graph
        self.emb_table = None
        self.emb_rows = None
        self.graph = None
        self.kvs = None
        self.kernels = 0
        self.pack_epoch = -1
        self._sampler = sampler

    def load_conditioning(self, cond, uncond):
        self.ctx[: self.n_c].copy_(cond[0])
        if self.cfg_on:
            self.ctx[self.n_c:].copy_(uncond[0])
        refresh = getattr(self._sampler.model, "refresh_conditioning", None)
        if refresh is not None:
            kvs = refresh(self.c2, self.x_in.shape[0])
            if self.graph is not None and kvs is not self.kvs:
                self.graph = None  # the model re-allocated its K / V^T buffers: the captured pointers are stale
            self.kvs = kvs

    def load_time_embedding(self, t_values: torch.Tensor):
        """emb_layers(time_embed(t)) of all steps in one pass (t alone decides them); the graph picks the row of its step."""
        fn = getattr(self._sampler.model, "time_embedding_rows", None)
        if fn is None or self.extra_info is None:
            return
        rows = fn(t_values)
        if self.emb_table is None or self.emb_table.shape[1] != rows.shape[1]:
            self.emb_table = torch.zeros(self.max_steps, rows.shape[1], dtype=torch.float32, device=rows.device)
            self.emb_rows = torch.zeros(self.x_in.shape[0], rows.shape[1], dtype=torch.float32, device=rows.device)
            self.extra_info["emb_rows"] = self.emb_rows
            self.graph = None
        self.emb_table[: rows.shape[0]].copy_(rows)

    def body(self, num_steps_tensorless):
        b = self.b
        self.x_in[:b].copy_(self.x)
        if self.cfg_on:
            self.x_in[b:].copy_(self.x)
        if self.emb_rows is not None:      # all samples of a step share its timestep (ddim.py:222: ts = full((b,), step))
            torch.index_select(self.emb_table, 0, self.step_idx.long().repeat(self.emb_rows.shape[0]), out=self.emb_rows)
        eps = self._sampler.model.apply_model(self.x_in, self.t_buf, self.c2)
        ops.cfg_ddim_update(self.x, eps, self.coef_table, self.x, self.pred, has_uncond=self.cfg_on, noise=self.noise,
                            step_idx=self.step_idx)
        ops.advance_step(self.step_idx, self.t_table, self.t_buf, self.max_steps)

    def capture(self):
        """Warm-up on a side stream (packs weights, fills caches, warms the allocator), then capture."""
        saved = (self.x.clone(), self.step_idx.clone(), self.t_buf.clone())
        self.step_idx.zero_()   # body() indexes the coefficient tables by the counter: a re-capture after a full run
        #                          would otherwise read row `total` (== max_steps: out of bounds)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.body(None)
        torch.cuda.current_stream().wait_stream(s)
        self.x.copy_(saved[0]); self.step_idx.copy_(saved[1]); self.t_buf.copy_(saved[2])
        g = torch.cuda.CUDAGraph()
        n0 = _lib.TRACE.count
        self.step_idx.zero_()
        with torch.cuda.graph(g):
            self.body(None)
        self.kernels = _lib.TRACE.count - n0
        self.pack_epoch = PackedModule.PACK_EPOCH
        self.x.copy_(saved[0]); self.step_idx.copy_(saved[1]); self.t_buf.copy_(saved[2])
        self.graph = g


def run(sampler, img, cond, uncond, steps, scales, temperature, log_every_t, intermediates):
    """DDIM loop of ddim.py:182-218 on captured graphs.  Returns (samples, intermediates)."""
    device = img.device
    b = img.shape[0]
    lat_shape = tuple(img.shape[1:])
    total = len(steps)
    rows, tvals, use_cfg = [], [], []
    for i in range(total):
        index = total - i - 1
        rows.append(sampler._coef_row(index, scales[i], False, temperature))
        tvals.append(float(steps[i]))
        use_cfg.append(not (uncond is None or scales[i] == 1.))
    with_noise = any(r[5] != 0.0 for r in rows)
    coef_host = torch.tensor(rows, dtype=torch.float32)
    t_host = torch.tensor(tvals, dtype=torch.float32)
    ekey = _extra_key(cond[2])
    cache = sampler.__dict__.setdefault("_step_graphs", {})

    def get(cfg_on):
        key = None
        if ekey is not None and hasattr(sampler.model, "refresh_conditioning"):
            key = (b, lat_shape, cfg_on, tuple(cond[0].shape), tuple(uncond[0].shape) if cfg_on else None, ekey,
                   with_noise, str(device))
        st = cache.get(key) if key is not None else None
        if st is not None and st.max_steps < total:
            st = None
        fresh = st is None
        if fresh:
            st = _StepGraph(sampler, b, lat_shape, cfg_on, cond, uncond, max(total, 64), device, with_noise)
            if key is not None:
                cache[key] = st
                if len(cache) > 4:
                    cache.pop(next(iter(cache)))
        st.coef_table[:total].copy_(coef_host, non_blocking=True)
        st.t_table[:total].copy_(t_host, non_blocking=True)
        st.load_time_embedding(st.t_table[:total])
        st.load_conditioning(cond, uncond)
        if st.graph is not None and st.pack_epoch != PackedModule.PACK_EPOCH:
            st.graph = None     # some module repacked its weights since the capture: the graph holds stale pointers
        if st.graph is None:
            st.capture()
        if cond[2] is not None:
            cond[2]["ca_layers_activations"] = {k: {} for k in ("outfeat", "attn", "attnscore", "q")}
        return st

    states = {}
    x_cur = img.float()
    prev = None
    for i in range(total):
        index = total - i - 1
        cfg_on = use_cfg[i]
        st = states.get(cfg_on)
        if st is None:
            st = states[cfg_on] = get(cfg_on)
        if st is not prev:  # (re-)enter this graph: hand over the latent, position the step counter
            st.x.copy_(x_cur if prev is None else prev.x)
            st.step_idx.fill_(i)
            st.t_buf.fill_(tvals[i])
            prev = st
        unscaled = noise_like(st.x.shape, device, False)                 # keeps the RNG stream in step (ddim.py:286)
        if st.noise is not None:
            st.noise.copy_(unscaled)
        st.graph.replay()
        sampler.graph_kernel_launches += st.kernels
        if index % log_every_t == 0 or index == total - 1:
            intermediates["x_inter"].append(st.x.clone())
            intermediates["pred_x0"].append(st.pred.clone())
    sampler._graphs = states
    return prev.x.clone(), intermediates
