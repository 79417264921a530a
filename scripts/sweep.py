"""BASELINE.json config #5: batch / resolution sweep - UNet batch 1..32 at 512^2 (64x64 latents) and 768^2 (96x96
latents, 9216 self-attention tokens): one eager UNet step per point with CUDA events around every C-ABI launch, reported
as step time, self-attention TFLOP/s at the top resolution and GroupNorm(+SiLU) GB/s over the 61 norm launches.
Usage: sweep.py [out.md]   (peaks from MEASURED_PEAKS.json if present)"""
import json, os, sys
import torch
sys.path.insert(0, ".")
from adaprompt_b200 import _lib
from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG
from adaprompt_b200.unet import UNetModel
from adaprompt_b200.weights import spec_of, synth_state_dict

peaks = {"tf": 1392.0, "gbs": 6530.0}
if os.path.exists("MEASURED_PEAKS.json"):
    try:
        mp = json.load(open("MEASURED_PEAKS.json"))
        peaks["tf"] = float(mp.get("bf16_tflops_sustained", mp.get("bf16_dense_tflops_sustained", peaks["tf"])))
        peaks["gbs"] = float(mp.get("hbm_gbs", mp.get("hbm_copy_gbs", peaks["gbs"])))
    except Exception:
        pass
with torch.device("meta"):
    unet = UNetModel(**SD15_UNET_CONFIG)
unet = unet.to_empty(device="cuda")
unet.load_state_dict(synth_state_dict(spec_of(unet), 1234))
unet.eval().prepare()
extra = {"use_layerwise_context": True, "use_conv_attn_kernel_size": -1, "placeholder2indices": None, "is_training": False}
lines = ["| latent | image | UNet batch | step ms | samples/s | UNet TFLOP/s (alg.) | self-attn top level TFLOP/s | all attention TFLOP/s | GroupNorm GB/s | frac of HBM peak |",
         "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
for L in (64, 96):
    for B in (1, 2, 4, 8, 16, 32):
        g = torch.Generator().manual_seed(B * 100 + L)
        x = torch.randn(B, 4, L, L, generator=g).cuda()
        t = torch.full((B,), 501.0, device="cuda")
        ctx = torch.randn(16 * B, 77, 768, generator=g).cuda()
        with torch.no_grad():
            for _ in range(2):
                eps = unet(x, t, context=ctx, extra_info=dict(extra))
            with _lib.profile() as prof:
                unet(x, t, context=ctx, extra_info=dict(extra))
            summ = prof.summary()
        assert torch.isfinite(eps).all()
        ms = sum(v["ms"] for v in summ.values())
        fl = sum(v["flops"] for v in summ.values())
        top = [v for k, v in prof.by_shape.items() if k.startswith("attention_bf16") and f"Nq{L * L} Nk{L * L} " in k]
        at = summ.get("af_attention_bf16")
        gn = summ.get("af_groupnorm_apply")
        top_tf = sum(v["flops"] for v in top) / (sum(v["ms"] for v in top) * 1e-3) / 1e12 if top else 0.0
        gbs = gn["bytes"] / (gn["ms"] * 1e-3) / 1e9
        lines.append(f"| {L}x{L} | {8 * L}^2 | {B} | {ms:.2f} | {B / ms * 1e3:.1f} | {fl / ms / 1e9:.0f} | {top_tf:.0f} | "
                     f"{at['flops'] / at['ms'] / 1e9:.0f} | {gbs:.0f} | {gbs / peaks['gbs']:.2f} |")
        print(lines[-1], flush=True)
        del x, ctx, eps
        torch.cuda.empty_cache()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.md"
hdr = (f"# r01 - batch / resolution sweep (BASELINE.json config #5)\n\nOne eager UNet step per point, CUDA events around every "
       f"launch (sum of kernel times; the CUDA-graph sampler runs the same kernels back to back). Peaks: {peaks['tf']:.0f} TFLOP/s "
       f"bf16 sustained, {peaks['gbs']:.0f} GB/s HBM copy.\n\n")
open(out, "w").write(hdr + "\n".join(lines) + "\n")
