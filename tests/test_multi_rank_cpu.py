"""world_size-2 gloo run on CPU of the multi-GPU sampling host logic: shard ownership, CFG pairs and
layerwise context rows staying with their image, gather order.  The sampler is a stand-in (no CUDA here); the
sharding code is the product code (adaprompt_b200/parallel_sampling.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeSampler:
    """sample() returns x_T + mean(context rows of the image) - mean(uncond rows): exposes any mis-sharding."""

    def sample(self, S, batch_size, shape, conditioning=None, unconditional_conditioning=None, guidance_scale=None,
               eta=0., x_T=None, verbose=False, **kw):
        c, prompts, _ = conditioning
        uc, _, _ = unconditional_conditioning
        assert c.shape[0] == 16 * batch_size == uc.shape[0] and len(prompts) == batch_size == x_T.shape[0]
        cm = c.reshape(batch_size, 16, -1).mean(dim=(1, 2)) - uc.reshape(batch_size, 16, -1).mean(dim=(1, 2))
        return x_T + cm.reshape(-1, 1, 1, 1), {}


def _worker(rank, world, port, n_images, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaprompt_b200.parallel_sampling import sample_sharded
    g = torch.Generator().manual_seed(3)
    c = torch.randn(16 * n_images, 7, 5, generator=g)
    uc = torch.randn(16 * n_images, 7, 5, generator=g)
    x_T = torch.randn(n_images, 4, 8, 8, generator=g)
    prompts = [f"p{i}" for i in range(n_images)]
    out = sample_sharded(_FakeSampler(), 50, n_images, [4, 8, 8], (c, prompts, {}), (uc, [""] * n_images, {}), (4.0, 1.0), x_T)
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from adaprompt_b200.parallel_sampling import shard_range
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_world2_gloo_sharded_sampling_matches_single_process(tmp_path):
    n_images, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(3)
    c = torch.randn(16 * n_images, 7, 5, generator=g)
    uc = torch.randn(16 * n_images, 7, 5, generator=g)
    x_T = torch.randn(n_images, 4, 8, 8, generator=g)
    ref, _ = _FakeSampler().sample(50, n_images, [4, 8, 8], conditioning=(c, ["p"] * n_images, {}),
                                   unconditional_conditioning=(uc, [""] * n_images, {}), x_T=x_T)
    for r in range(world):
        out = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(out, ref)


def test_wrapper_prompt_rewriting():
    from adaprompt_b200.adaface_wrapper import AdaFaceWrapper
    w = object.__new__(AdaFaceWrapper)
    w.subject_string = "z"
    w.placeholder_tokens_str = " ".join(f"z_{i}" for i in range(16))
    assert w.update_prompt("a photo of z in a park") == f"a photo of {w.placeholder_tokens_str} in a park"
    assert w.update_prompt("a zebra") == w.placeholder_tokens_str + " a zebra"        # 'z' must be a whole word
    assert w.update_prompt(w.placeholder_tokens_str + " x") == w.placeholder_tokens_str + " x"
