"""B200 parity of the first-stage decoder (adaprompt_b200/vae.py -> C ABI -> sm_100a kernels) against the reference's
own outputs (tests/golden/vae.pt) and against the fp32 CPU oracle at larger latents.  Tolerance: north_star states
budgets for eps (1e-2 per step) and final latents (2e-2) only; the decoder is 30 bf16-operand conv layers deep (fp32
accumulation, fp32 residual stream) behind those latents, so the image is held to the same end-of-pipeline budget:
rel-L2 < 2e-2 against the fp32 reference (measured: 1.05e-2 at 8x8 latents with random-init weights)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae.pt")
TOL = 2e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def vae_and_sd():
    from adaprompt_b200.vae import AutoencoderKL
    from adaprompt_b200.weights import synth_state_dict
    from oracle.vae_oracle import VAESpec
    gold = torch.load(GOLD)
    sd = synth_state_dict(VAESpec().state_spec(), gold["seed"])
    with torch.device("meta"):
        vae = AutoencoderKL()
    vae = vae.to_empty(device="cuda")
    vae.load_state_dict(sd)
    vae.eval()
    return vae, sd, gold


def test_softmax_rows_matches_torch():
    from adaprompt_b200 import ops
    g = torch.Generator().manual_seed(3)
    for rows, n in ((64, 64), (37, 256), (16, 4096), (5, 9216)):
        x = (torch.randn(rows, n, generator=g) * 30).cuda()
        p = ops.softmax_rows(x, 0.0442)
        ref = torch.softmax(x.cpu() * 0.0442, dim=1)
        assert (p.float().cpu() - ref).abs().max() < 4e-3 * ref.max() + 1e-6
        assert (p.float().sum(1).cpu() - 1).abs().max() < 5e-3


def test_channel_mix4_matches_torch():
    from adaprompt_b200 import ops
    g = torch.Generator().manual_seed(4)
    x, w, b = torch.randn(3, 4, 9, 7, generator=g), torch.randn(4, 4, generator=g), torch.randn(4, generator=g)
    y = ops.channel_mix4(x.cuda(), w.cuda(), b.cuda(), 1. / 0.18215)
    ref = torch.nn.functional.conv2d(x / 0.18215, w.view(4, 4, 1, 1), b)
    assert _rel(y, ref) < 1e-6


@pytest.mark.parametrize("name", ["b2_8", "b1_16"])
def test_vae_decode_matches_reference_golden(vae_and_sd, name):
    from adaprompt_b200 import _lib
    from adaprompt_b200.vae import decode_first_stage
    from oracle.vae_oracle import vae_latents
    vae, _, gold = vae_and_sd
    n0 = _lib.TRACE.count
    img = decode_first_stage(vae, vae_latents(name).cuda())
    torch.cuda.synchronize()
    assert _lib.TRACE.count - n0 > 100                      # the CUDA path ran (no fallback exists)
    assert img.shape == gold[name]["image"].shape and img.dtype == torch.float32
    e = _rel(img, gold[name]["image"])
    print(f"vae {name}: image rel-L2 vs reference fp32 = {e:.3e}")
    assert e < TOL


def test_vae_decode_matches_oracle_256px(vae_and_sd):
    """32x32 latents -> 256x256 RGB: covers the 64-wide conv tiles, 1024-token AttnBlock and batch chunking."""
    from adaprompt_b200.vae import decode_first_stage
    from oracle.vae_oracle import VAESpec, decode_first_stage as oracle_decode, vae_latents
    vae, sd, _ = vae_and_sd
    z = vae_latents("b2_32")
    from adaprompt_b200 import ops
    img = decode_first_stage(vae, z.cuda())
    with ops.launch_options(split_k=1):        # whole-tile GEMM schedule: the arithmetic of a sample is independent of the batch
        img_whole = decode_first_stage(vae, z.cuda())
        img_chunked = decode_first_stage(vae, z.cuda(), max_batch=1)
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = oracle_decode(sd, VAESpec(), z)
    e = _rel(img, ref)
    print(f"vae b2_32: image rel-L2 vs fp32 oracle = {e:.3e}")
    assert e < TOL
    assert torch.equal(img_whole, img_chunked)              # per-sample arithmetic: chunking cannot change a bit
    # the split-K schedule (default) depends on the tile count, i.e. on the batch: same values up to bf16 rounding flips
    assert _rel(img, img_whole) < TOL and _rel(img_whole, ref) < TOL


def test_vae_decode_512px_properties(vae_and_sd):
    """Full size (64x64 latents -> 512x512, the BASELINE config's image size; the 4096-token AttnBlock and the
    64x2-pixel conv tiles only occur here): shape, finiteness and run-to-run bit reproducibility."""
    from adaprompt_b200.vae import decode_first_stage
    from oracle.vae_oracle import vae_latents
    vae, _, _ = vae_and_sd
    z = vae_latents("b1_64").cuda()
    a = decode_first_stage(vae, z)
    b = decode_first_stage(vae, z)
    torch.cuda.synchronize()
    assert a.shape == (1, 3, 512, 512) and torch.isfinite(a).all()
    assert torch.equal(a, b)


def test_txt2img_cli_end_to_end_synthetic(tmp_path):
    """Row N3: prompts -> conditioning (ID tokens spliced) -> DDIM with annealed CFG -> decode_first_stage -> PNGs,
    through the command-line entry point on random-init weights of the real architecture."""
    from PIL import Image
    from adaprompt_b200 import _lib, txt2img
    n0 = _lib.TRACE.count
    txt2img.main(["--synthetic", "--synthetic_clip_layers", "2", "--prompt", "a photo of a z in a park", "--ddim_steps", "4",
                  "--n_samples", "2", "--H", "256", "--W", "256", "--scale", "4", "1", "--outdir", str(tmp_path),
                  "--no_cuda_graph", "--save_latents"])
    assert _lib.TRACE.count - n0 > 1000
    files = sorted(os.listdir(tmp_path / "samples"))
    assert [f for f in files if f.endswith(".png")] == ["r0-00000.png", "r0-00001.png"]
    assert Image.open(str(tmp_path / "samples" / "r0-00000.png")).size == (256, 256)
    assert Image.open(str(tmp_path / "grid-r0.png")).size == (512, 256)
    lat = torch.load(str(tmp_path / "samples" / "r0-00000-latents.pt"))
    assert lat.shape == (2, 4, 32, 32) and torch.isfinite(lat).all()


def test_txt2img_cli_output_matches_the_oracle_pipeline(tmp_path):
    """Row N3, checked against the checker: with the conditioning the CLI built (saved by --save_conditioning), the fp32
    oracle chain - DDIM with annealed CFG on the UNet oracle, then the first-stage decoder oracle - must reproduce the
    latents the CLI saved (rel-L2 < 2e-2, north_star) and the PNG it wrote (mean |diff| < 2 grey levels)."""
    import numpy as np
    from PIL import Image
    from adaprompt_b200 import txt2img
    from adaprompt_b200.weights import synth_state_dict
    from oracle.golden_inputs import EXTRA_INFO
    from oracle.unet_oracle import UNetSpec, ddim_sample, unet_forward
    from oracle.vae_oracle import VAESpec, decode_first_stage as oracle_decode
    txt2img.main(["--synthetic", "--synthetic_clip_layers", "2", "--prompt", "a photo of a z", "--ddim_steps", "4",
                  "--n_samples", "1", "--H", "256", "--W", "256", "--scale", "4", "1", "--outdir", str(tmp_path),
                  "--save_latents", "--save_conditioning", "--seed_weights", "1234"])
    lat = torch.load(str(tmp_path / "samples" / "r0-00000-latents.pt"))
    cond = torch.load(str(tmp_path / "samples" / "r0-00000-cond.pt"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    uspec, vspec = UNetSpec(), VAESpec()
    sd_u = synth_state_dict(uspec.state_spec(), 1234)
    sd_v = synth_state_dict(vspec.state_spec(), 1235)
    apply = lambda x, t, c: unet_forward(sd_u, uspec, x, t, c[0], dict(c[2]))
    with torch.no_grad():
        ref_lat, _ = ddim_sample(apply, 4, [1, 4, 32, 32], (cond["c"], ["p"], dict(EXTRA_INFO)),
                                 (cond["uc"], [""], dict(EXTRA_INFO)), (4.0, 1.0), cond["x_T"])
        ref_img = torch.clamp((oracle_decode(sd_v, vspec, ref_lat) + 1.0) / 2.0, 0.0, 1.0)
    e = _rel(lat, ref_lat)
    png = np.asarray(Image.open(str(tmp_path / "samples" / "r0-00000.png"))).astype(np.float32)
    ref_png = (255.0 * ref_img[0].permute(1, 2, 0).numpy()).round().clip(0, 255)
    mad = float(np.abs(png - ref_png).mean())
    print(f"txt2img vs oracle pipeline: latent rel-L2 {e:.3e}, PNG mean abs diff {mad:.2f} / 255")
    assert e < 2e-2 and mad < 2.0
