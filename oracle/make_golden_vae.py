"""TEST INFRASTRUCTURE - generates tests/golden/vae.pt by running the UNMODIFIED reference first-stage decoder
(ldm.modules.diffusionmodules.model.Decoder from /root/reference, build container only) plus a plain
torch.nn.Conv2d post_quant_conv (ldm/models/autoencoder.py:306,330-333; AutoencoderKL itself needs pytorch_lightning,
which is not installed) on the synthetic-weight recipe.  Run:  python oracle/make_golden_vae.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True

from adaprompt_b200.weights import synth_state_dict  # noqa: E402
from oracle.vae_oracle import VAESpec, vae_latents  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    sys.path.insert(0, "/root/reference")
    with contextlib.redirect_stdout(io.StringIO()):
        from ldm.modules.diffusionmodules.model import Decoder
        dec = Decoder(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                      num_res_blocks=2, attn_resolutions=[], dropout=0.0)          # v1-inference-ada.yaml:58-72
    pq = torch.nn.Conv2d(4, 4, 1)                                                   # autoencoder.py:306
    spec = VAESpec()
    sd = synth_state_dict(spec.state_spec(), 4321)
    ref_keys = ["decoder." + k for k in dec.state_dict().keys()] + ["post_quant_conv.weight", "post_quant_conv.bias"]
    assert ref_keys == list(sd.keys()), "oracle state_spec must list the reference's parameters in its order"
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    pq.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
    dec.eval()
    out = {"seed": 4321}
    with torch.no_grad():
        for name in ("b2_8", "b1_16"):
            z = vae_latents(name)
            img = dec(pq(1. / 0.18215 * z))                                         # ddpm.py:1267, autoencoder.py:330-333
            out[name] = {"z_sum": float(z.double().abs().sum()), "image": img.clone()}
            print(name, tuple(img.shape), "|img|", float(img.abs().mean()))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "vae.pt"))


if __name__ == "__main__":
    main()
