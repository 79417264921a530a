"""Multi-GPU sampling: DDIM trajectories are independent per image, so the image batch is sharded
across ranks - one process per GPU, each image's cond/uncond CFG pair on the same rank, weights and the per-prompt
K/V cache replicated - with NO data-path collective.  The only communication is the optional gather of the
finished latents.  The reference itself is single-GPU for inference (stable_txt2img.py:226,330)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` owned by `rank`; the first n_items % world_size ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_conditioning(cond, n_images: int, world_size: int, rank: int, layers: int = 16):
    """Slices a (c [layers*B, N, D], prompts, extra_info) conditioning tuple to this rank's images.  The layer index is
    minor to the batch index ('(b l)' order, embedding_manager.py:1349), so image i owns rows [i*layers, (i+1)*layers)."""
    c, prompts, extra = cond
    if c.shape[0] != n_images * layers:
        raise ValueError(f"context has {c.shape[0]} rows, expected {n_images}*{layers}")
    b, e = shard_range(n_images, world_size, rank)
    return (c[b * layers:e * layers], list(prompts[b:e]), _shard_extra(extra, n_images, b, e))


def _shard_extra(extra, n_images: int, b: int, e: int):
    """Per-image entries of extra_info follow the shard: tensors whose leading dimension is the global batch (img_mask
    [B,1,H,W], prompt_emb_mask [B,77,1], ...) are sliced, the placeholder index pairs (indices_B, indices_N) of
    `placeholder2indices` (embedding_manager.py:1292-1330) keep the entries of this rank's images, re-based to the local
    batch; scalars and flags are shared."""
    if extra is None:
        return None
    out = {}
    for k, v in extra.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == n_images:
            out[k] = v[b:e]
        elif k == "placeholder2indices" and isinstance(v, dict):
            local = {}
            for name, pair in v.items():
                if pair is None:
                    local[name] = None
                    continue
                iB, iN = pair
                keep = (iB >= b) & (iB < e)
                local[name] = (iB[keep] - b, iN[keep])
            out[k] = local
        else:
            out[k] = v
    return out


def sample_sharded(sampler, S: int, n_images: int, shape: Sequence[int], conditioning, unconditional_conditioning,
                   guidance_scale, x_T: torch.Tensor, eta: float = 0.0, gather: bool = True, **kw):
    """Runs `sampler.sample` on this rank's slice of the global batch (every rank passes the SAME global arguments)
    and, if `gather`, returns the global [n_images, ...] latents on every rank (all_gather of padded shards)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    b, e = shard_range(n_images, world, rank)
    local = None
    if e > b:
        local, _ = sampler.sample(S, e - b, list(shape), conditioning=shard_conditioning(conditioning, n_images, world, rank),
                                  unconditional_conditioning=shard_conditioning(unconditional_conditioning, n_images, world, rank),
                                  guidance_scale=guidance_scale, eta=eta, x_T=x_T[b:e], verbose=False, **kw)
    if not gather or world == 1:
        return local
    per = (n_images + world - 1) // world
    pad = torch.zeros((per,) + tuple(x_T.shape[1:]), dtype=torch.float32, device=x_T.device)
    if local is not None:
        pad[: e - b] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = [parts[r][: shard_range(n_images, world, r)[1] - shard_range(n_images, world, r)[0]] for r in range(world)]
    return torch.cat(out, 0)
