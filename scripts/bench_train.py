"""Stage-1 distillation step timing (BASELINE.json configs[3]): per GPU batch 4 x 2 gradient-accumulation micro-steps,
64x64 latent, UNet forward + activation backward, SubjBasisGenerator (12-layer trainable CLIP text model) forward /
backward with weight gradients, frozen CLIP, one flat-bucket all-reduce of the gradients per optimizer step (NCCL over
NVLink when launched with torchrun, N > 1), clip by norm 0.5.  Synthetic data, random-init weights; the teacher's eps is
a fixed random tensor.  Prints one JSON line on rank 0.

    python scripts/bench_train.py [--steps 3] [--warmup 2] [--bs 4] [--accum 2]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py
"""
import argparse
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


class StubTokenizer:
    pad_token_id = 49407
    vocab = {"photo": 1125, "of": 539, "a": 320, "id": 1014, "person": 2533, ",": 267, "z": 345}

    def _ids(self, text):
        return [self.vocab[w] for w in text.replace(",", " , ").split()]

    def encode(self, text, add_special_tokens=False):
        return self._ids(text)

    def __call__(self, text, truncation=True, padding="max_length", max_length=77, return_tensors="pt", **kw):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [([49406] + self._ids(t) + [49407] * max_length)[:max_length] for t in texts]
        return types.SimpleNamespace(input_ids=torch.tensor(rows))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--bs", type=int, default=4)
    ap.add_argument("--accum", type=int, default=2)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--breakdown", action="store_true", help="per-entry-point CUDA-event times of one optimizer step (stderr)")
    args = ap.parse_args()
    import torch.distributed as dist
    from adaprompt_b200 import _lib
    from adaprompt_b200.clip_text import CLIPTextModelWrapper
    from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG
    from adaprompt_b200.subj_basis_generator import SubjBasisGenerator
    from adaprompt_b200.train_cond import DistillStep, allreduce_gradients, clip_grad_norm, trainable_parameters
    from adaprompt_b200.unet import UNetModel
    from adaprompt_b200.weights import spec_of, synth_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    with torch.device("meta"):
        unet = UNetModel(**SD15_UNET_CONFIG)
    unet = unet.to_empty(device=dev)
    unet.load_state_dict(synth_state_dict(spec_of(unet), 1234))
    unet.eval().prepare()
    for p in unet.parameters():
        p.requires_grad = False

    def clip(seed, train):
        m = CLIPTextModelWrapper().to(dev)
        sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
        for k in sd:
            if "embedding" in k:
                sd[k] = torch.randn(sd[k].shape, generator=torch.Generator().manual_seed(seed)) * 0.02
        m.load_state_dict(sd)
        for p in m.parameters():
            p.requires_grad = train
        return m

    tok = StubTokenizer()
    sbg = SubjBasisGenerator(num_out_embs_per_layer=16, clip_tokenizer=tok)
    sbg.prompt2token_proj = clip(41, True)
    sbg = sbg.to(dev).train()
    frozen, arc2face = clip(42, False), clip(43, False).eval()
    frozen.text_model.last_layers_skip_weights = [0.5, 0.5]
    acp = torch.linspace(0.9991, 0.0047, 1000)
    step = DistillStep(unet, frozen.text_model, sbg, arc2face, tok, acp, 345)
    params = trainable_parameters(sbg)
    n_params = sum(p.numel() for p in params)
    g = torch.Generator().manual_seed(100 + rank)
    H = args.latent

    def batch():
        b = args.bs
        prompt = [49406, 320, 1125, 539, 320, 345] + [267] * 15 + [49407] * 56
        return {"x0": torch.randn(b, 4, H, H, generator=g).to(dev), "noise": torch.randn(b, 4, H, H, generator=g).to(dev),
                "t": torch.randint(0, 1000, (b,), generator=g).to(dev),
                "teacher_eps": torch.randn(b, 4, H, H, generator=g).to(dev),
                "face_embs": torch.nn.functional.normalize(torch.randn(b, 512, generator=g), dim=-1).to(dev),
                "tokens": torch.tensor([prompt] * b).to(dev)}

    def optimizer_step():
        for p in params:
            p.grad = None
        losses = [step.micro_step(batch(), accum=args.accum) for _ in range(args.accum)]
        allreduce_gradients(params, world)
        gn = clip_grad_norm(params, 0.5)
        return sum(losses) / len(losses), gn

    for _ in range(args.warmup):
        loss, gn = optimizer_step()
    torch.cuda.synchronize()
    if args.breakdown and rank == 0:
        with _lib.profile() as prof:
            optimizer_step()
        summ = prof.summary()
        tot = sum(v["ms"] for v in summ.values())
        print(f"one optimizer step: {sum(v['launches'] for v in summ.values())} C-ABI calls, {tot:.1f} ms of kernel time", file=sys.stderr)
        for n, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:16]:
            print(f"  {n:28s} x{v['launches']:5d} {v['ms']:8.2f} ms  {v['flops'] / max(v['ms'], 1e-9) / 1e9:7.1f} TFLOP/s", file=sys.stderr)
        for n, v in sorted(prof.by_shape.items(), key=lambda kv: -kv[1]["ms"])[:24]:
            print(f"    {n:60s} x{v['launches']:4d} {v['ms']:8.2f} ms", file=sys.stderr)
    if world > 1:
        dist.barrier()
    n0 = _lib.TRACE.count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss, gn = optimizer_step()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    launches = (_lib.TRACE.count - n0) // args.steps
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    if rank == 0:
        print(json.dumps({"metric": "stage1_distill_samples_per_sec", "value": world * args.bs * args.accum / (ms_step / 1e3),
                          "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_optimizer_step": ms_step, "wall_ms_per_step": wall * 1e3 / args.steps,
                          "config": {"workload": "stage1_distill_bs4x2accum_64x64", "micro_batch": args.bs,
                                     "grad_accum": args.accum, "latent": [4, H, H], "trainable_params": n_params,
                                     "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0},
                          "loss": loss, "grad_norm": gn, "kernel_launches_per_step": launches,
                          "peak_memory_gib": peak_gb, "dtype": "bf16 operands / fp32 residual stream and gradients",
                          "data": "synthetic", "scaling": "weak"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
