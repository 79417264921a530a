"""Host-side schedule arithmetic of the sampler (scalar / 1000-element vector work; numpy float64 like the reference's
host code, cast to fp32 where the reference casts).

What the SD-1.5 / AdaFace configuration actually uses (configs/stable-diffusion/v1-inference-ada.yaml:5-9 and
DDIMSampler.make_schedule, ldm/models/diffusion/ddim.py:28-68):
  * beta schedule "linear" (linear in sqrt(beta), ldm/modules/diffusionmodules/util.py:23-25),
  * "uniform" DDIM time-step selection (util.py:48-50,57),
  * sigma_t = eta * sqrt((1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)) (util.py:63-69).
Other schedule names the reference's helper accepts are not part of this path and raise.
"""
from __future__ import annotations

import numpy as np
import torch


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """float64 numpy betas.  Only schedule == "linear": equally spaced in sqrt(beta), then squared."""
    if schedule != "linear":
        raise NotImplementedError(f"beta schedule '{schedule}': the AdaFace / SD-1.5 path uses 'linear' only")
    root = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64)
    return (root * root).numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """Every (T // S)-th training step, shifted by one: 1, 21, ..., 981 for S = 50, T = 1000."""
    if ddim_discr_method != "uniform":
        raise NotImplementedError(f"ddim discretisation '{ddim_discr_method}': only 'uniform' is on this path")
    stride = num_ddpm_timesteps // num_ddim_timesteps
    return np.arange(0, num_ddpm_timesteps, stride) + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """(sigmas, alphas, alphas_prev) of the selected steps.  alphacums: CPU fp32 tensor (as make_schedule passes it);
    alphas stays an fp32 tensor, alphas_prev / sigmas become float64 numpy arrays of fp32 values - the mixed types are
    what the reference's arithmetic sees, and the fp32 coefficient rows are derived from exactly these."""
    alphas = alphacums[ddim_timesteps]
    shifted = alphacums[ddim_timesteps[:-1]].tolist()
    alphas_prev = np.asarray([alphacums[0]] + shifted)
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def noise_like(shape, device, repeat=False):
    """One standard-normal draw per call from the global generator of `device` (util.py:267-270); with `repeat` the
    first sample's noise is shared by the batch."""
    if not repeat:
        return torch.randn(shape, device=device)
    one = torch.randn((1, *shape[1:]), device=device)
    return one.expand(shape[0], *shape[1:]).contiguous()


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """util.py:154-174 on the device kernel (af_timestep_embedding)."""
    from . import ops
    if repeat_only or max_period != 10000:
        raise NotImplementedError
    return ops.timestep_embedding(timesteps.float().contiguous(), dim)
