"""Host mirror of ldm/models/diffusion/ddim.py (DDIMSampler) on the B200 C ABI.

Same constructor / sample() / ddim_sampling() / p_sample_ddim() / stochastic_encode() / decode()
signatures and the same quirks:
  * guidance_scale must be a (max, min) tuple; it is annealed linearly by repeated Python-float
    subtraction (ddim.py:168-180,215-218);
  * the UNet batch is [conditional ; unconditional] (conditional FIRST, ddim.py:238-243) and extra_info is
    taken from the conditional tuple (:235,:247);
  * `noise_like` is drawn every step even at eta = 0 (:286) so the global RNG stream advances exactly
    as in the reference.
The CFG combine + x0 / x_prev update is one fused kernel (af_cfg_ddim_update) that follows the reference's
fp32 operation order bit for bit.  When the call is "plain" (no mask / callbacks / score corrector) the
whole denoising step - UNet + update - is captured once in a CUDA graph and replayed per step with the
timestep, guidance scale and alpha coefficients read from device tables.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from .diffusion_util import make_ddim_sampling_parameters, make_ddim_timesteps, noise_like


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_cuda_graph = kwargs.get("use_cuda_graph", True)
        self.verbose_progress = kwargs.get("progress", False)
        self._graphs = {}
        self.graph_kernel_launches = 0   # kernels executed through CUDA-graph replays (bench.py: gpu_launches)

    def register_buffer(self, name, attr):
        if type(attr) == torch.Tensor:
            if attr.device != torch.device("cuda"):
                attr = attr.to(torch.device("cuda"))
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """ddim.py:28-68."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize,
                                                  num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        self.register_buffer("betas", to_torch(self.model.betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(self.model.alphas_cumprod_prev))
        ac = alphas_cumprod.cpu()
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(ac)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1. - ac)))
        self.register_buffer("log_one_minus_alphas_cumprod", to_torch(np.log(1. - ac)))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_torch(np.sqrt(1. / ac)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_torch(np.sqrt(1. / ac - 1)))
        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=ac, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.register_buffer("ddim_sigmas", ddim_sigmas)
        self.register_buffer("ddim_alphas", ddim_alphas)
        self.register_buffer("ddim_alphas_prev", ddim_alphas_prev)
        self.register_buffer("ddim_sqrt_one_minus_alphas", np.sqrt(1. - ddim_alphas))
        sigmas_for_original_sampling_steps = ddim_eta * torch.sqrt(
            (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * (
                    1 - self.alphas_cumprod / self.alphas_cumprod_prev))
        self.register_buffer("ddim_sigmas_for_original_num_steps", sigmas_for_original_sampling_steps)

    # ------------------------------------------------------------------ per-step scalar coefficients
    def _coef_row(self, index, guidance_scale, use_original_steps=False, temperature=1.):
        """[g, sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma*temperature, 0, 0] computed
        with the same fp32 torch scalar ops as ddim.py:267-283."""
        alphas = self.model.alphas_cumprod if use_original_steps else self.ddim_alphas
        alphas_prev = self.model.alphas_cumprod_prev if use_original_steps else self.ddim_alphas_prev
        s1m = self.model.sqrt_one_minus_alphas_cumprod if use_original_steps else self.ddim_sqrt_one_minus_alphas
        sigmas = self.ddim_sigmas_for_original_num_steps if use_original_steps else self.ddim_sigmas
        a_t = torch.full((1,), float(alphas[index]))
        a_prev = torch.full((1,), float(alphas_prev[index]))
        sigma_t = torch.full((1,), float(sigmas[index]))
        s1m_t = torch.full((1,), float(s1m[index]))
        dcoef = (1. - a_prev - sigma_t ** 2).sqrt()
        g = torch.full((1,), float(guidance_scale))  # python float -> fp32, as in `guidance_scale * (e_t - e_u)`
        return [float(g), float(s1m_t), float(a_t.sqrt()), float(a_prev.sqrt()), float(dcoef),
                float(sigma_t * temperature), 0.0, 0.0]

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, guidance_scale=1., unconditional_conditioning=None, **kwargs):
        """ddim.py:71-132."""
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for DDIM sampling is {size}, eta {eta}")
        samples, intermediates = self.ddim_sampling(
            conditioning, size, callback=callback, img_callback=img_callback, quantize_denoised=quantize_x0,
            mask=mask, x0=x0, ddim_use_original_steps=False, noise_dropout=noise_dropout, temperature=temperature,
            score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
            guidance_scale=guidance_scale, unconditional_conditioning=unconditional_conditioning, **kwargs)
        return samples, intermediates

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      guidance_scale=1., unconditional_conditioning=None, **kwargs):
        """ddim.py:135-220."""
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif timesteps is not None and not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = reversed(range(0, timesteps)) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]

        if isinstance(guidance_scale, (list, tuple)):
            max_guide_scale, min_guide_scale = guidance_scale
        else:
            # the reference leaves max_guide_scale unbound here (ddim.py:169-173): same failure mode
            raise UnboundLocalError("guidance_scale must be a (max, min) list/tuple (reference ddim.py:169-173)")
        max_guide_anneal_steps = total_steps - 1
        guide_scale_step_delta = (max_guide_scale - min_guide_scale) / max_guide_anneal_steps
        guide_scale = max_guide_scale
        steps = [int(s) for s in time_range]

        plain = (mask is None and callback is None and img_callback is None and score_corrector is None
                 and not quantize_denoised and noise_dropout == 0. and not ddim_use_original_steps
                 and isinstance(cond, tuple) and img.is_cuda)
        if plain and self.use_cuda_graph:
            scales = []
            g = guide_scale
            for i in range(total_steps):
                scales.append(g)
                g = g - guide_scale_step_delta if i <= max_guide_anneal_steps else 1
            from . import graph_sampler
            return graph_sampler.run(self, img, cond, unconditional_conditioning, steps, scales, temperature,
                                     log_every_t, intermediates)

        for i, step in enumerate(steps):
            index = total_steps - i - 1
            ts = torch.full((b,), step, device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            outs = self.p_sample_ddim(img, cond, ts, index=index, use_original_steps=ddim_use_original_steps,
                                      quantize_denoised=quantize_denoised, temperature=temperature,
                                      noise_dropout=noise_dropout, score_corrector=score_corrector,
                                      corrector_kwargs=corrector_kwargs, guidance_scale=guide_scale,
                                      unconditional_conditioning=unconditional_conditioning, **kwargs)
            img, pred_x0 = outs
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
            if i <= max_guide_anneal_steps:
                guide_scale = guide_scale - guide_scale_step_delta
            else:
                guide_scale = 1
        return img, intermediates

    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      guidance_scale=1., unconditional_conditioning=None):
        """ddim.py:222-296 (one fused update kernel instead of ~12 elementwise launches)."""
        b, device = x.shape[0], x.device
        has_uncond = not (unconditional_conditioning is None or guidance_scale == 1.)
        if not has_uncond:
            eps = self.model.apply_model(x, t, c)
        else:
            x_in = torch.cat([x] * 2)
            t_in = torch.cat([t] * 2)
            if isinstance(c, tuple):
                c_c, c_in_c, extra_info = c
                c_u, c_in_u, _ = unconditional_conditioning
                c2 = (self._twin(c_c, c_u), sum([c_in_c, c_in_u], []), extra_info)
            else:
                c2 = self._twin(c, unconditional_conditioning)
            eps = self.model.apply_model(x_in, t_in, c2)
        if score_corrector is not None:
            raise NotImplementedError("score_corrector")
        if quantize_denoised:
            raise NotImplementedError("quantize_denoised (VQ first stage) is not part of the SD-1.5 path")
        row = self._coef_row(index, guidance_scale, use_original_steps, temperature)
        coef = torch.tensor([row], dtype=torch.float32, device=device)
        unscaled_noise = noise_like(x.shape, device, repeat_noise)          # drawn even when sigma == 0 (:286)
        sigma = row[5]
        noise = None
        if sigma != 0.0:
            noise = unscaled_noise
            if noise_dropout > 0.:
                noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        x_prev = torch.empty_like(x)
        pred_x0 = torch.empty_like(x)
        ops.cfg_ddim_update(x.float().contiguous(), eps.contiguous(), coef, x_prev, pred_x0, has_uncond=has_uncond,
                            noise=noise)
        return x_prev, pred_x0

    def _twin(self, c_c, c_u):
        """torch.cat([c_c, c_u]) (ddim.py:243) - memoised so the UNet's per-prompt K/V cache hits on every step."""
        key = (id(c_c), c_c.data_ptr(), c_c._version, id(c_u), c_u.data_ptr(), c_u._version)
        ent = self.__dict__.get("_twin_cache")
        if ent is not None and ent[0] == key:
            return ent[3]
        twin = torch.cat([c_c, c_u])
        self.__dict__["_twin_cache"] = (key, c_c, c_u, twin)
        return twin

    # ------------------------------------------------------------------ CUDA-graph fast path
    def _graph_sampling(self, img, cond, uncond, steps, scales, temperature, log_every_t, intermediates):
        """One captured graph per (shape, cfg on/off): [x ; x] -> UNet -> fused CFG + DDIM update -> advance."""
        device = img.device
        b = img.shape[0]
        total = len(steps)
        c_c, c_in_c, extra_info = cond
        rows, tvals, use_cfg = [], [], []
        for i in range(total):
            index = total - i - 1
            rows.append(self._coef_row(index, scales[i], False, temperature))
            tvals.append(float(steps[i]))
            use_cfg.append(not (uncond is None or scales[i] == 1.))
        coef_table = torch.tensor(rows, dtype=torch.float32, device=device)
        t_table = torch.tensor(tvals, dtype=torch.float32, device=device)
        sigma_nonzero = any(r[5] != 0.0 for r in rows)

        x = img.float().clone().contiguous()
        pred = torch.empty_like(x)
        step_idx = torch.zeros(1, dtype=torch.int32, device=device)
        noise = torch.empty_like(x) if sigma_nonzero else None
        graphs = {}
        kernels_in_graph = {}

        def build(cfg_on):
            nb = 2 * b if cfg_on else b
            t_buf = torch.full((nb,), tvals[0], dtype=torch.float32, device=device)
            x_in = torch.empty((nb,) + tuple(x.shape[1:]), dtype=torch.float32, device=device)
            if cfg_on:
                c_u, c_in_u, _ = uncond
                c2 = (self._twin(c_c, c_u), sum([c_in_c, c_in_u], []), extra_info)
            else:
                c2 = cond

            def body():
                x_in[:b].copy_(x)
                if cfg_on:
                    x_in[b:].copy_(x)
                eps = self.model.apply_model(x_in, t_buf, c2)
                ops.cfg_ddim_update(x, eps, coef_table, x, pred, has_uncond=cfg_on, noise=noise, step_idx=step_idx)
                ops.advance_step(step_idx, t_table, t_buf, total)

            # warm-up on a side stream (fills the K/V cache, packs weights, warms the allocator), then capture
            saved = (x.clone(), step_idx.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                body()
            torch.cuda.current_stream().wait_stream(s)
            x.copy_(saved[0])
            step_idx.copy_(saved[1])
            t_buf.fill_(tvals[0])
            g = torch.cuda.CUDAGraph()
            n0 = _lib.TRACE.count
            with torch.cuda.graph(g):
                body()
            kernels_in_graph[cfg_on] = _lib.TRACE.count - n0
            x.copy_(saved[0])
            step_idx.copy_(saved[1])
            # the graph holds raw pointers: everything it reads or writes must stay referenced with it
            return g, t_buf, (x_in, c2, coef_table, t_table, step_idx, x, pred, noise)

        for i in range(total):
            index = total - i - 1
            cfg_on = use_cfg[i]
            if cfg_on not in graphs:
                graphs[cfg_on] = build(cfg_on)
            g, t_buf, _keepalive = graphs[cfg_on]
            if i == 0 or use_cfg[i - 1] != cfg_on:
                t_buf.fill_(tvals[i])
            unscaled = noise_like(x.shape, device, False)                   # keeps the RNG stream in step (:286)
            if noise is not None:
                noise.copy_(unscaled)
            g.replay()
            self.graph_kernel_launches += kernels_in_graph[cfg_on]
            if index % log_every_t == 0 or index == total - 1:
                intermediates["x_inter"].append(x.clone())
                intermediates["pred_x0"].append(pred.clone())
        self._graphs = graphs  # keep alive until the next call
        return x.clone(), intermediates

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """ddim.py:299-312."""
        if use_original_steps:
            sqrt_alphas_cumprod = self.sqrt_alphas_cumprod
            sqrt_one_minus_alphas_cumprod = self.sqrt_one_minus_alphas_cumprod
        else:
            sqrt_alphas_cumprod = torch.sqrt(self.ddim_alphas)
            sqrt_one_minus_alphas_cumprod = self.ddim_sqrt_one_minus_alphas
        if noise is None:
            noise = torch.randn_like(x0)

        def extract(a, t, x_shape):
            out = a.to(t.device).gather(-1, t)
            return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))

        return (extract(sqrt_alphas_cumprod, t, x0.shape) * x0 +
                extract(sqrt_one_minus_alphas_cumprod, t, x0.shape) * noise)

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, guidance_scale=1.0, unconditional_conditioning=None,
               use_original_steps=False):
        """ddim.py:315-350."""
        timesteps = np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps
        timesteps = timesteps[:t_start]
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        max_guide_scale = guidance_scale
        min_guide_scale = min(2.0, max_guide_scale)
        max_guide_anneal_steps = total_steps - 1
        guide_scale_step_delta = (max_guide_scale - min_guide_scale) / max_guide_anneal_steps
        guide_scale = max_guide_scale
        x_dec = x_latent
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((x_latent.shape[0],), int(step), device=x_latent.device, dtype=torch.long)
            x_dec, _ = self.p_sample_ddim(x_dec, cond, ts, index=index, use_original_steps=use_original_steps,
                                          guidance_scale=guide_scale,
                                          unconditional_conditioning=unconditional_conditioning)
            if i <= max_guide_anneal_steps:
                guide_scale = guide_scale - guide_scale_step_delta
            else:
                guide_scale = 1
        return x_dec
