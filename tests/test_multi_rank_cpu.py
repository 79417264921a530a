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


def test_shard_conditioning_slices_per_image_entries_of_extra_info():
    """img_mask / prompt_emb_mask rows and the placeholder index pairs follow the rank's images (they were
    shared unsliced); the union over the ranks is the global conditioning."""
    import torch
    from adaprompt_b200.parallel_sampling import shard_conditioning, shard_range
    n, layers = 5, 16
    c = torch.arange(n * layers * 2 * 3, dtype=torch.float32).reshape(n * layers, 2, 3)
    iB = torch.tensor([0, 0, 2, 2, 3, 4, 4])
    iN = torch.tensor([5, 6, 5, 6, 9, 5, 6])
    extra = {"use_layerwise_context": True, "img_mask": torch.arange(n * 4.).reshape(n, 1, 2, 2),
             "prompt_emb_mask": torch.ones(n, 77, 1), "placeholder2indices": {"z": (iB, iN), "y": None},
             "use_conv_attn_kernel_size": -1}
    seen_rows, seen_pairs = [], []
    for rank in range(2):
        b, e = shard_range(n, 2, rank)
        cc, prompts, ex = shard_conditioning((c, [f"p{i}" for i in range(n)], extra), n, 2, rank)
        assert cc.shape[0] == (e - b) * layers and prompts == [f"p{i}" for i in range(b, e)]
        assert torch.equal(ex["img_mask"], extra["img_mask"][b:e]) and ex["prompt_emb_mask"].shape[0] == e - b
        lb, ln = ex["placeholder2indices"]["z"]
        assert ex["placeholder2indices"]["y"] is None and bool((lb >= 0).all()) and bool((lb < e - b).all())
        seen_pairs += [(int(x) + b, int(y)) for x, y in zip(lb, ln)]
        seen_rows.append(cc)
        assert ex["use_layerwise_context"] is True and ex["use_conv_attn_kernel_size"] == -1
    assert torch.equal(torch.cat(seen_rows), c)
    assert seen_pairs == list(zip(iB.tolist(), iN.tolist()))
    assert extra["placeholder2indices"]["z"][0] is iB            # the caller's dict is untouched


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


# ------------------------------------------------------------------------------------------------ training all-reduce
def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaprompt_b200.train_cond import allreduce_gradients, clip_grad_norm
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    g = torch.Generator().manual_seed(100 + rank)
    params[0].grad = torch.randn(5, 3, generator=g)
    params[1].grad = torch.randn(7, generator=g)          # params[2] has no gradient on any rank: zeros in the bucket
    if rank == 1:
        params[1].grad = None                             # ... and params[1] none on rank 1: the bucket size must not change
    flat = allreduce_gradients(params, world)
    for p_, o in zip(params, (0, 64, 128)):              # flat_layout(): 64-element aligned starts
        p_.grad = flat[o:o + p_.numel()].view(p_.shape)
    total = clip_grad_norm(params, 0.5)
    torch.save({"g0": params[0].grad, "g1": params[1].grad, "flat": flat, "total": total}, os.path.join(out_dir, f"g{rank}.pt"))
    dist.destroy_process_group()


def test_world2_gloo_gradient_allreduce_is_the_mean(tmp_path):
    """Stage-1 step: one flat bucket, all-reduce SUM / world, then clip by norm 0.5."""
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    g0 = [torch.randn(5, 3, generator=g) for g in gens]
    g1 = [torch.randn(7, generator=g) for g in gens]
    g1[1] = torch.zeros(7)
    m0, m1 = sum(g0) / world, sum(g1) / world
    total = float(torch.cat([m0.reshape(-1), m1.reshape(-1)]).norm())
    scale = min(1.0, 0.5 / (total + 1e-6))
    for r in range(world):
        o = torch.load(os.path.join(str(tmp_path), f"g{r}.pt"))
        assert o["flat"].numel() == 192 and abs(o["total"] - total) < 1e-5     # fixed size on every rank (3 aligned slots)
        assert torch.allclose(o["g0"], m0 * scale, atol=1e-6) and torch.allclose(o["g1"], m1 * scale, atol=1e-6)


def _bucket_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaprompt_b200.train_cond import GradBucket
    w = torch.nn.Parameter(torch.ones(4, 3))
    b = torch.nn.Parameter(torch.zeros(3))
    unused = torch.nn.Parameter(torch.zeros(2))
    bucket = GradBucket([w, b, unused])
    bucket.begin_step()
    x = torch.arange(8.0).reshape(2, 4) + rank
    for _ in range(2):                                   # two accumulated micro-batches
        ((x @ w + b).sum() / 2).backward()
    assert w.grad.data_ptr() == bucket.views[0].data_ptr()          # autograd accumulated IN the bucket
    work = bucket.allreduce(world, async_op=True)
    bucket.wait()
    total = bucket.clip_(0.5)
    torch.save({"flat": bucket.flat.clone(), "total": float(total)}, os.path.join(out_dir, f"b{rank}.pt"))
    dist.destroy_process_group()


def test_world2_gloo_grad_bucket_accumulates_allreduces_and_clips(tmp_path):
    world = 2
    mp.spawn(_bucket_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gw = sum((torch.arange(8.0).reshape(2, 4) + r).sum(0)[:, None].expand(4, 3) for r in range(world)) / world
    gb = torch.full((3,), 2.0)
    ref = torch.zeros(192)
    ref[:12], ref[64:67] = gw.reshape(-1), gb
    scale = min(1.0, 0.5 / (float(ref.norm()) + 1e-6))
    for r in range(world):
        o = torch.load(os.path.join(str(tmp_path), f"b{r}.pt"))
        assert abs(o["total"] - float(ref.norm())) < 1e-4
        assert torch.allclose(o["flat"], ref * scale, atol=1e-6)
