"""TEST INFRASTRUCTURE - seeded synthetic inputs shared by oracle/make_golden.py and the tests.
No dependency on /root/reference: the tests regenerate the inputs and verify their checksums
against the ones stored next to the golden outputs."""
from __future__ import annotations

import torch

EXTRA_INFO = {"use_layerwise_context": True, "use_conv_attn_kernel_size": -1, "placeholder2indices": None,
              "is_training": False}


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def _g(seed):
    return torch.Generator().manual_seed(seed)


def unet_inputs(name: str):
    """-> (x [B,4,H,W], t [B] int64, context [16B,77,768], extra_info)."""
    extra = dict(EXTRA_INFO)
    if name == "b1_t501_64":
        g = _g(7)
        return torch.randn(1, 4, 64, 64, generator=g), torch.full((1,), 501, dtype=torch.long), \
            torch.randn(16, 77, 768, generator=g), extra
    if name == "b2_t981_21_64":
        g = _g(8)
        return torch.randn(2, 4, 64, 64, generator=g), torch.tensor([981, 21]), \
            torch.randn(32, 77, 768, generator=g), extra
    if name == "b2_t501_32":
        g = _g(9)
        return torch.randn(2, 4, 32, 32, generator=g), torch.tensor([501, 501]), \
            torch.randn(32, 77, 768, generator=g), extra
    if name == "b1_t261_mask_32":
        g = _g(10)
        x = torch.randn(1, 4, 32, 32, generator=g)
        ctx = torch.randn(16, 77, 768, generator=g)
        mask = (torch.rand(1, 1, 32, 32, generator=g) > 0.4).float()
        mask[:, :, :2] = 1.0
        extra["img_mask"] = mask
        return x, torch.full((1,), 261, dtype=torch.long), ctx, extra
    if name == "b16_t501_64":          # the benchmarked UNet batch (8 images x [cond ; uncond]) at 64x64
        g = _g(11)
        return torch.randn(16, 4, 64, 64, generator=g), torch.full((16,), 501, dtype=torch.long), \
            torch.randn(256, 77, 768, generator=g), extra
    if name == "b2_t741_hijk_32":      # iter_type mix_hijk (openaimodel.py:885-892): context = [v_ctx | k_ctx] along tokens
        g = _g(12)
        extra["iter_type"] = "mix_hijk"
        return torch.randn(2, 4, 32, 32, generator=g), torch.tensor([741, 741]), \
            torch.randn(32, 154, 768, generator=g), extra
    if name == "b2_t341_compel_32":    # compel-style CFG on the context (openaimodel.py:898-916), inference batch mask
        g = _g(13)
        extra.update(apply_compel_cfg_prob=0.7, compel_cfg_weight_level_range=(1, 3),
                     empty_context=torch.randn(1, 77, 768, generator=g), python_random_seed=1234)
        return torch.randn(2, 4, 32, 32, generator=g), torch.tensor([341, 341]), \
            torch.randn(32, 77, 768, generator=g), extra
    if name == "b2_t601_convattn_capture_32":
        # conv attention (use_conv_attn_kernel_size 3: attention.py:208-216, ldm/util.py:700-878) on sample 0 only (its
        # prompt holds the 16 subject tokens at positions 5..20; sample 1 has none) + attention / feature capture
        # (capture_distill_attn: openaimodel.py:947-952,984-988,1031-1035)
        g = _g(14)
        extra.update(use_conv_attn_kernel_size=3, capture_distill_attn=True,
                     placeholder2indices={"z": (torch.zeros(16, dtype=torch.long), torch.arange(5, 21))})
        return torch.randn(2, 4, 32, 32, generator=g), torch.tensor([601, 601]), \
            torch.randn(32, 77, 768, generator=g), extra
    raise KeyError(name)


def module_inputs():
    g = _g(21)
    r = lambda *s: torch.randn(*s, generator=g)
    mask = (torch.rand(1, 1, 32, 32, generator=g) > 0.35).float()
    mask[:, :, 0] = 1.0
    return {
        "res_out5": {"x": r(1, 1920, 8, 8), "emb": r(1, 1280)},
        "res_in1": {"x": r(1, 320, 16, 16), "emb": r(1, 1280)},
        "st_in4": {"x": r(1, 640, 16, 16), "ctx": r(1, 77, 768), "mask": mask},
        "st_in1": {"x": r(1, 320, 32, 32), "ctx": r(1, 77, 768)},
        "st_mid": {"x": r(3, 1280, 8, 8), "ctx": r(3, 77, 768)},
        "ca_in4": {"x": r(2, 128, 640), "ctx": r(2, 77, 768)},
        "down_in3": {"x": r(1, 320, 32, 32)},
        "up_out2": {"x": r(1, 1280, 4, 4)},
    }


def ddim_inputs(name: str):
    """-> (S, shape (b,4,H,W), cond tuple, uncond tuple, guidance_scale, x_T)."""
    if name == "s10_32_g4_1":
        g, S, hw = _g(31), 10, 32
    elif name == "s50_64_g4_1":
        g, S, hw = _g(32), 50, 64
    elif name == "s50_32_g10_4":       # the CLI's default scale (stable_txt2img.py:152), x_inter logged at EVERY step
        g, S, hw = _g(33), 50, 32
    else:
        raise KeyError(name)
    b = 1
    c = torch.randn(16 * b, 77, 768, generator=g)
    uc = torch.randn(16 * b, 77, 768, generator=g)
    x_T = torch.randn(b, 4, hw, hw, generator=g)
    cond = (c, ["a photo of a z"] * b, dict(EXTRA_INFO))
    uncond = (uc, [""] * b, dict(EXTRA_INFO))
    return S, (b, 4, hw, hw), cond, uncond, (10.0, 4.0) if name.endswith("g10_4") else (4.0, 1.0), x_T
