"""One eager UNet denoising step at the headline geometry (batch 16 = 8 CFG pairs, 64x64 latent), for ncu.
Usage: unet_step.py [reps] [batch]   (prints the number of kernel launches per step)"""
import sys
import torch
sys.path.insert(0, ".")
from adaprompt_b200 import _lib
from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG
from adaprompt_b200.unet import UNetModel
from adaprompt_b200.weights import spec_of, synth_state_dict

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
with torch.device("meta"):
    unet = UNetModel(**SD15_UNET_CONFIG)
unet = unet.to_empty(device="cuda")
unet.load_state_dict(synth_state_dict(spec_of(unet), 1234))
unet.eval().prepare()
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 4, 64, 64, generator=g).cuda()
t = torch.full((B,), 501.0, device="cuda")
ctx = torch.randn(16 * B, 77, 768, generator=g).cuda()
extra = {"use_layerwise_context": True, "use_conv_attn_kernel_size": -1, "placeholder2indices": None,
         "is_training": False}
import json
with torch.no_grad():
    with _lib.profile() as prof:
        unet(x, t, context=ctx, extra_info=dict(extra))
    prof.summary()
    json.dump([(r[0], r[5], _lib.KERNELS_PER_CALL[r[0]], r[3], r[4]) for r in prof.records],
              open("gpurun_out/step_calls.json", "w"))
    for i in range(reps):
        n0 = _lib.TRACE.count
        eps = unet(x, t, context=ctx, extra_info=dict(extra))
        torch.cuda.synchronize()
        print(f"step {i}: {_lib.TRACE.count - n0} launches, |eps| {eps.abs().mean().item():.4f}", flush=True)
