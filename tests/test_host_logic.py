"""CPU checks of the host-side mirror logic that needs no kernel: DDIM schedule / coefficient rows of the
product sampler against the oracle restatement, flag side channel, K/V cache keying."""
import numpy as np
import pytest
import torch


class _FakeModel:
    """Schedule-only stand-in (no UNet) so DDIMSampler.make_schedule can run on CPU."""

    def __init__(self):
        from adaprompt_b200.ldm_lite import LatentDiffusionLite
        lite = LatentDiffusionLite.__new__(LatentDiffusionLite)
        torch.nn.Module.__init__(lite)
        lite.register_schedule("linear", 1000, 0.00085, 0.012)
        self.__dict__.update(betas=lite.betas, alphas_cumprod=lite.alphas_cumprod,
                             alphas_cumprod_prev=lite.alphas_cumprod_prev, num_timesteps=1000,
                             device=torch.device("cpu"),
                             sqrt_one_minus_alphas_cumprod=lite.sqrt_one_minus_alphas_cumprod)


def _cpu_sampler():
    from adaprompt_b200.ddim import DDIMSampler

    class S(DDIMSampler):
        def register_buffer(self, name, attr):   # the reference forces cuda here (ddim.py:22-26)
            setattr(self, name, attr)

    return S(_FakeModel())


def test_schedule_matches_oracle():
    from oracle.unet_oracle import ddim_schedule, make_alphas_cumprod
    s = _cpu_sampler()
    s.make_schedule(50, ddim_eta=0.0, verbose=False)
    ts, alphas, alphas_prev, sigmas, s1m = ddim_schedule(50)
    assert np.array_equal(s.ddim_timesteps, ts)
    assert torch.equal(s.ddim_alphas, alphas)
    assert np.array_equal(np.asarray(s.ddim_alphas_prev), np.asarray(alphas_prev))
    assert torch.equal(torch.as_tensor(s.ddim_sqrt_one_minus_alphas), torch.as_tensor(s1m))
    ac = torch.tensor(make_alphas_cumprod(), dtype=torch.float32)
    assert torch.equal(s.alphas_cumprod, ac)


def test_coef_rows_match_reference_scalar_ops():
    """Row = the fp32 scalars p_sample_ddim builds with torch.full (ddim.py:273-283)."""
    from oracle.unet_oracle import ddim_schedule, guidance_schedule
    s = _cpu_sampler()
    s.make_schedule(50, ddim_eta=0.0, verbose=False)
    ts, alphas, alphas_prev, sigmas, s1m = ddim_schedule(50)
    gs = guidance_schedule(50, (4.0, 1.0))
    for i in (0, 7, 49):
        index = 50 - i - 1
        row = s._coef_row(index, gs[i])
        a_t = torch.full((1,), alphas[index])
        a_prev = torch.full((1,), alphas_prev[index])
        sig = torch.full((1,), sigmas[index])
        assert row[0] == float(torch.tensor(gs[i], dtype=torch.float32))
        assert row[1] == float(torch.full((1,), s1m[index]))
        assert row[2] == float(a_t.sqrt()) and row[3] == float(a_prev.sqrt())
        assert row[4] == float((1. - a_prev - sig ** 2).sqrt()) and row[5] == 0.0


def test_set_cross_attn_flags_roundtrip():
    """openaimodel.py:723-824: layerwise flags land on attn2 of the 16 CA layers and are restorable."""
    from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG
    from adaprompt_b200.unet import ALL_CA_LAYER_INDICES, UNetModel
    with torch.device("meta"):
        m = UNetModel(**SD15_UNET_CONFIG)
    sizes = np.arange(16)
    old, _ = m.set_cross_attn_flags(ca_flag_dict={"use_conv_attn_kernel_size:layerwise": sizes, "is_training": False})
    mods = m._layer_modules()
    assert len(mods) == 25
    for ca_idx, layer_idx in enumerate(ALL_CA_LAYER_INDICES):
        attn2 = mods[layer_idx][1].transformer_blocks[0].attn2
        assert attn2.use_conv_attn_kernel_size == ca_idx and attn2.is_training is False
    assert old == {"use_conv_attn_kernel_size:layerwise": [-1] * 16, "is_training": True}
    m.set_cross_attn_flags(ca_flag_dict=old)
    assert mods[1][1].transformer_blocks[0].attn2.use_conv_attn_kernel_size == -1
    old2, _ = m.set_cross_attn_flags(ca_flag_dict={"save_attn_vars": True}, ca_layer_indices=[7, 8, 12])
    assert mods[7][1].transformer_blocks[0].attn2.save_attn_vars and not mods[1][1].transformer_blocks[0].attn2.save_attn_vars
    m.set_cross_attn_flags(ca_flag_dict=old2, ca_layer_indices=[7, 8, 12])
    assert m.set_cross_attn_flags() == (None, None)


def test_unet_block_layout_matches_survey():
    """22 ResBlocks, 16 SpatialTransformers, 3 Downsample, 3 Upsample; head dims 40/80/160."""
    from adaprompt_b200.attention import SpatialTransformer
    from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG
    from adaprompt_b200.unet import Downsample, ResBlock, UNetModel, Upsample
    with torch.device("meta"):
        m = UNetModel(**SD15_UNET_CONFIG)
    count = lambda t: sum(isinstance(x, t) for x in m.modules())
    assert (count(ResBlock), count(SpatialTransformer), count(Downsample), count(Upsample)) == (22, 16, 3, 3)
    mods = m._layer_modules()
    heads = {i: mods[i][1].transformer_blocks[0].attn1.dim_head for i in m._ca_modules()}
    assert all(heads[i] == 40 for i in (1, 2, 22, 23, 24))
    assert all(heads[i] == 80 for i in (4, 5, 19, 20, 21))
    assert all(heads[i] == 160 for i in (7, 8, 12, 16, 17, 18))


def test_packing_layouts():
    from adaprompt_b200.packing import pack_conv3x3, pack_geglu, pack_qk
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = pack_conv3x3(w)
    assert p.shape == (2, 3, 3, 3) and float(p[1, 2, 0, 1]) == float(w[1, 1, 2, 0])
    wg = torch.arange(512 * 2, dtype=torch.float32).reshape(512, 2)
    bg = torch.arange(512, dtype=torch.float32)
    wp, bp = pack_geglu(wg, bg)
    assert torch.equal(bp[:128], bg[:128]) and torch.equal(bp[128:256], bg[256:384]) and torch.equal(bp[256:384], bg[128:256])
    assert torch.equal(wp[128:256], wg[256:384])
    wq = torch.randn(320, 320)
    pq = pack_qk(wq, wq, 8)
    assert pq.shape == (2 * 8 * 48, 320)
    assert float(pq.view(2, 8, 48, 320)[:, :, 40:].abs().max()) == 0.0   # zero pad rows -> zero pad columns


# ------------------------------------------------------------------------------------------------ txt2img CLI (row N3)
def test_txt2img_cli_host_logic(tmp_path):
    from adaprompt_b200 import txt2img as t2i
    a = t2i.parse_args(["--synthetic", "--scale", "4", "1", "--n_samples", "3"])
    assert a.scale == [4.0, 1.0] and a.ddim_steps == 50 and a.subject_string == "z"
    assert t2i.parse_args(["--synthetic", "--scale", "5"]).scale == [5.0, 5.0]
    with pytest.raises(SystemExit):
        t2i.parse_args(["--prompt", "x"])                      # neither --ckpt nor --synthetic
    assert t2i.expand_prompt("a photo of a z in a park", "z", 4) == "a photo of a z , , ,  in a park".replace("  ", " ") \
        or t2i.expand_prompt("a photo of a z in a park", "z", 4).split() == "a photo of a z , , , in a park".split()
    tok = t2i.HashTokenizer()
    ids = tok(["a photo of a z, , ,", "hello world"]).input_ids
    assert ids.shape == (2, 77) and ids[0, 0] == t2i.BOS and ids[0, 5] == t2i.TOK_Z and ids[0, 6] == t2i.TOK_COMMA
    assert (ids[1, 3:] == t2i.EOS).all() and tok.encode("z")[0] == t2i.TOK_Z
    sd = {"state_dict": {"model.diffusion_model.out.2.bias": torch.zeros(4), "first_stage_model.decoder.conv_in.bias": torch.zeros(512),
                         "cond_stage_model.transformer.text_model.final_layer_norm.bias": torch.zeros(768), "model_ema.decay": torch.zeros(())}}
    parts = t2i.split_sd15_checkpoint(sd)
    assert list(parts["unet"]) == ["out.2.bias"] and list(parts["vae"]) == ["decoder.conv_in.bias"]
    assert list(parts["clip"]) == ["text_model.final_layer_norm.bias"]
    imgs = torch.rand(3, 3, 16, 16)
    paths = t2i.save_images(imgs, str(tmp_path / "s"), 7, "r0")
    t2i.save_grid(imgs, str(tmp_path / "g.png"), 2)
    from PIL import Image
    assert [p.split("-")[-1] for p in paths] == ["00007.png", "00008.png", "00009.png"]
    assert Image.open(str(tmp_path / "g.png")).size == (32, 32)


# ------------------------------------------------------------------------------------------------ Arc2Face teacher (row N4)
def test_diffusers_unet_key_map_is_a_bijection_onto_the_unet():
    from adaprompt_b200.arc2face_teacher import (convert_diffusers_unet_state_dict,
                                                 convert_ldm_unet_state_dict_to_diffusers)
    from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG
    from adaprompt_b200.unet import UNetModel
    with torch.device("meta"):
        unet = UNetModel(**SD15_UNET_CONFIG)
    sd = unet.state_dict()
    hf = convert_ldm_unet_state_dict_to_diffusers(sd)
    assert len(hf) == len(sd) == len(set(hf))
    for k in ("conv_in.weight", "time_embedding.linear_2.bias", "down_blocks.0.resnets.1.time_emb_proj.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "down_blocks.3.resnets.0.norm1.weight",
              "mid_block.attentions.0.transformer_blocks.0.attn2.to_k.weight", "mid_block.resnets.1.conv2.bias",
              "up_blocks.0.upsamplers.0.conv.weight", "up_blocks.1.resnets.0.conv_shortcut.weight",
              "up_blocks.3.attentions.2.proj_out.weight", "conv_norm_out.weight", "conv_out.bias"):
        assert k in hf, k
    assert not any(k.startswith(("input_blocks", "output_blocks", "middle_block", "out.", "time_embed.")) for k in hf)
    back = convert_diffusers_unet_state_dict(hf)
    assert list(back) == list(sd) and all(back[k].shape == sd[k].shape for k in sd)
    # Linear-style 1x1 projections of newer diffusers checkpoints are reshaped to the reference's Conv2d layout
    k = "mid_block.attentions.0.proj_in.weight"
    hf2 = dict(hf); hf2[k] = torch.empty(1280, 1280, device="meta")
    assert convert_diffusers_unet_state_dict(hf2)["middle_block.1.proj_in.weight"].shape == (1280, 1280, 1, 1)
    with pytest.raises(KeyError):
        convert_diffusers_unet_state_dict({"class_embedding.weight": torch.empty(1)})


def test_arc2face_teacher_multi_step_logic():
    """ddpm.py:5434-5480 restated by hand for 3 denoising steps with a stand-in UNet (same RNG call order)."""
    from adaprompt_b200.arc2face_teacher import Arc2FaceTeacher
    from adaprompt_b200.ldm_lite import LatentDiffusionLite

    class FakeUNet(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(0.3))
            self.calls = []

        def forward(self, x, t, context=None, extra_info=None):
            self.calls.append((tuple(context.shape), t.clone()))
            return self.w * x + 0.01 * t.view(-1, 1, 1, 1).float() / 1000 + context.mean()

    ldm = LatentDiffusionLite(unet=torch.nn.Identity())
    fake = FakeUNet()
    teacher = Arc2FaceTeacher(fake)
    g = torch.Generator().manual_seed(5)
    x0, noise = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    t, ctx = torch.tensor([800, 500]), torch.randn(2, 21, 768, generator=g)
    torch.manual_seed(11)
    preds, x0s, noises, ts = teacher(ldm, x0, noise, t, ctx, num_denoising_steps=3)
    assert len(preds) == 3 and len(x0s) == 3 and len(noises) == 3 and len(ts) == 3
    assert fake.calls[0][0] == (32, 21, 768)                      # one copy of the context per cross-attention layer
    assert all((ts[i + 1] < ts[i]).all() for i in range(2))
    # hand restatement
    torch.manual_seed(11)
    acp = ldm.alphas_cumprod
    xs, ns, tt = [x0], [noise], [t]
    for i in range(3):
        a = acp[tt[i]].view(-1, 1, 1, 1)
        xn = a.sqrt() * xs[i] + (1 - a).sqrt() * ns[i]
        pr = 0.3 * xn + 0.01 * tt[i].view(-1, 1, 1, 1).float() / 1000 + ctx.repeat_interleave(16, 0).mean()
        assert torch.allclose(pr, preds[i], atol=1e-6)
        xs.append(torch.sqrt(1.0 / a) * xn - torch.sqrt(1.0 / a - 1) * pr)
        assert torch.allclose(xs[-1], x0s[i], atol=1e-5)
        if i < 2:
            r = torch.rand_like(tt[i].float())
            lb, ub = tt[i] * np.power(0.5, np.power(2, -0.3)), tt[i] * np.power(0.7, np.power(2, -0.3))
            tt.append(((ub - lb) * r + lb).long())
            assert torch.equal(tt[-1], ts[i + 1])
            ns.append(torch.randn_like(xs[-1]))


# ------------------------------------------------------------------------------------------------ checkpoint boundary
def test_reference_format_checkpoint_loads_without_the_reference_package():
    """tests/golden/adaface_ckpt_tiny.pt was written by the UNMODIFIED reference classes (oracle/make_golden_ckpt.py):
    pickled adaface.subj_basis_generator.SubjBasisGenerator inside an nn.ModuleDict, reference CLIPTextModelWrapper,
    transformers CLIP modules, one CLIPAttentionMKV layer (embedding_manager.py:1824-1838).  It must load here with
    neither `adaface` nor `ldm` importable, and come out as the native mirror with identical tensors / attributes."""
    import os
    import sys
    from adaprompt_b200.checkpoint import load_adaface_ckpt
    from adaprompt_b200.clip_text import CLIPAttentionMKV, CLIPTextModelWrapper
    from adaprompt_b200.subj_basis_generator import SubjBasisGenerator
    gold = os.path.join(os.path.dirname(__file__), "golden")
    assert "adaface" not in sys.modules and "ldm" not in sys.modules
    ckpt = load_adaface_ckpt(os.path.join(gold, "adaface_ckpt_tiny.pt"))
    exp = torch.load(os.path.join(gold, "adaface_ckpt_tiny_expected.pt"))
    assert exp["class_of_sbg"] == "adaface.subj_basis_generator.SubjBasisGenerator"
    assert ckpt["subject_strings"] == ["z"] and ckpt["do_zero_shot"] is True and ckpt["token2num_vectors"] == {"z": 16}
    sbg = ckpt["string_to_subj_basis_generator_dict"]["z"]
    assert type(sbg) is SubjBasisGenerator and type(sbg.prompt2token_proj) is CLIPTextModelWrapper
    sd = sbg.prompt2token_proj.state_dict()
    for k, v in exp["prompt2token_proj"].items():
        if k.endswith("position_ids"):
            continue
        assert torch.equal(sd[k], v), k
    layers = sbg.prompt2token_proj.text_model.encoder.layers
    assert isinstance(layers[1].self_attn, CLIPAttentionMKV) and layers[1].self_attn.multiplier == 2
    assert layers[1].self_attn.k_proj.weight.shape == (256, 128) and layers[0].self_attn.k_proj.weight.shape == (128, 128)
    assert torch.equal(sbg.hidden_state_layer_weights.detach(), exp["hidden_state_layer_weights"])
    assert torch.equal(sbg.pos_embs.detach(), exp["pos_embs"]) and torch.equal(sbg.pad_embeddings, exp["pad_embeddings"])
    assert (sbg.num_out_layers, sbg.num_out_embs_per_layer, sbg.prompt2token_proj_attention_multiplier) == (16, 16, 2)
    assert sbg.prompt2token_proj_grad_scale == 0.4 and sbg.placeholder_is_bg is False
    # AdaFaceWrapper.load_subj_basis_generator(adaface_ckpt_path) without an injected generator (adaface_wrapper.py:49-59)
    from adaprompt_b200.adaface_wrapper import AdaFaceWrapper
    w = object.__new__(AdaFaceWrapper)
    torch.nn.Module.__init__(w)
    w.subject_string, w.device, w.is_training = "z", "cpu", False
    w._injected = {"subj_basis_generator": None, "tokenizer": None}
    w.load_subj_basis_generator(os.path.join(gold, "adaface_ckpt_tiny.pt"))
    assert type(w.subj_basis_generator) is SubjBasisGenerator and w.subj_basis_generator.num_out_layers == 1   # :59
    assert not w.subj_basis_generator.training
    w.subject_string = "y"
    with pytest.raises(KeyError):
        w.load_subj_basis_generator(os.path.join(gold, "adaface_ckpt_tiny.pt"))
    # the 0.4 / 5 gradient scalers are real: identity forward, scaled backward
    x = torch.ones(3, requires_grad=True)
    sbg.prompt2token_proj_grad_scaler(x).sum().backward()
    assert torch.allclose(x.grad, torch.full((3,), 0.4))
    w = sbg.hidden_state_layer_weights
    sbg.hidden_state_layer_weights_grad_scaler(w).sum().backward()
    assert torch.allclose(w.grad, torch.full_like(w, 5.0))


def test_import_path_aliases_resolve_to_the_mirrors():
    """Drop-in boundary: the reference's dotted paths (v1-inference-ada.yaml:36,
    embedding_manager.py:7-8 legacy names) resolve to the B200 mirrors after install_import_aliases()."""
    import importlib
    import sys
    from adaprompt_b200.checkpoint import install_import_aliases
    before = {k for k in sys.modules if k == "ldm" or k == "adaface" or k.startswith(("ldm.", "adaface."))}
    try:
        installed = install_import_aliases()
        assert "ldm.modules.diffusionmodules.openaimodel" in installed
        from adaprompt_b200.ddim import DDIMSampler
        from adaprompt_b200.subj_basis_generator import SubjBasisGenerator
        from adaprompt_b200.unet import UNetModel
        assert importlib.import_module("ldm.modules.diffusionmodules.openaimodel").UNetModel is UNetModel
        assert importlib.import_module("ldm.models.diffusion.ddim").DDIMSampler is DDIMSampler
        assert importlib.import_module("adaface.subj_basis_generator").SubjBasisGenerator is SubjBasisGenerator
        assert importlib.import_module("ldm.modules.subj_basis_generator").SubjBasisGenerator is SubjBasisGenerator
        import ldm.modules.attention as att
        assert att.CrossAttention.__module__ == "adaprompt_b200.attention"
        # instantiate_from_config-style lookup (ldm/util.py:104-111)
        module, cls = "ldm.modules.diffusionmodules.openaimodel.UNetModel".rsplit(".", 1)
        assert getattr(importlib.import_module(module), cls) is UNetModel
    finally:
        for k in [k for k in sys.modules if (k == "ldm" or k == "adaface" or k.startswith(("ldm.", "adaface."))) and k not in before]:
            del sys.modules[k]
