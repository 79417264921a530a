"""B200: the conditioning path through the C ABI against the CPU oracle and the
reference golden vectors (tests/golden/text_path.pt, made by oracle/make_golden_text.py from the unmodified
reference).  Floating-point tolerance: bf16 tensor-core operands with fp32 accumulation and fp32 residual stream ->
rel-L2 < 1e-2 (north_star tolerance for a network forward); splice rows / indices / masks: torch.equal."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "text_path.pt")
TOL = 1e-2


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


class StubTokenizer:
    """Fixed id rows for the prompt templates of the path (no CLIP vocabulary offline) - same as the golden script."""
    pad_token_id = 49407

    def __init__(self):
        from oracle import text_oracle as to
        self.to = to
        self.vocab = {"photo": to.TOK_PHOTO, "of": to.TOK_OF, "a": to.TOK_A, "id": to.TOK_ID, "person": to.TOK_PERSON,
                      ",": to.TOK_COMMA, "z": to.TOK_Z, "y": to.TOK_Y}

    def _ids(self, text):
        return [self.vocab[w] for w in text.replace(",", " , ").split()]

    def encode(self, text, add_special_tokens=False):
        return self._ids(text)

    def __call__(self, text, truncation=True, padding="max_length", max_length=77, return_tensors="pt", **kw):
        texts = [text] if isinstance(text, str) else list(text)
        return types.SimpleNamespace(input_ids=torch.tensor([self.to.pad_ids(self._ids(t), max_length) for t in texts]))


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def _clip(seed, kv_mult=None):
    from adaprompt_b200.clip_text import CLIPTextModelWrapper
    from oracle import text_oracle as to
    sd = to.clip_synth_state_dict(seed, kv_mult=kv_mult)
    m = CLIPTextModelWrapper()
    if kv_mult:
        m.extend_clip_attention_MKV_multiplier(-1, -1, kv_mult[0], noise_std=0)
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


@pytest.fixture(scope="module")
def models(gold):
    s = gold["seeds"]
    return {k: _clip(s[k]) for k in ("arc2face", "sbg", "frozen")}


def test_state_dict_keys_are_hf_names(models):
    from oracle import text_oracle as to
    m, _ = models["sbg"]
    assert list(m.state_dict().keys()) == list(to.clip_state_spec().keys())


@pytest.mark.parametrize("mult", [1, 2, 4])
def test_attention_small_mkv_vs_reference_module(gold, mult):
    """af_attention_small (+ the kv row repack) against the reference CLIPAttentionMKV output."""
    from adaprompt_b200 import ops
    from adaprompt_b200.weights import synth_state_dict
    g = gold[f"mkv_attn_m{mult}"]
    spec = {}
    for n, o in (("k_proj", 768 * mult), ("v_proj", 768 * mult), ("q_proj", 768), ("out_proj", 768)):
        spec[f"{n}.weight"] = (o, 768)
        spec[f"{n}.bias"] = (o,)
    sd = {k: v.cuda() for k, v in synth_state_dict(spec, g["seed"]).items()}
    x = g["x"].cuda()
    B, L, E, H, hd = 2, 22, 768, 12, 64
    kv = lambda w: w.reshape(mult, H, hd, *w.shape[1:]).transpose(0, 1).reshape(mult * E, *w.shape[1:])
    wqkv = torch.cat([sd["q_proj.weight"], kv(sd["k_proj.weight"]), kv(sd["v_proj.weight"])]).to(torch.bfloat16).contiguous()
    bqkv = torch.cat([sd["q_proj.bias"], kv(sd["k_proj.bias"]), kv(sd["v_proj.bias"])]).contiguous()
    xb = x.reshape(B * L, E).to(torch.bfloat16).contiguous()
    qkv = torch.empty(B * L, E * (1 + 2 * mult), device="cuda", dtype=torch.bfloat16)
    ops.gemm(xb, wqkv, qkv, bias=bqkv)
    o = torch.empty(B * L, E, device="cuda", dtype=torch.bfloat16)
    ops.attention_small(qkv, o, B=B, heads=H, L=L, k_off=E, v_off=E + E * mult, mult=mult, scale=hd ** -0.5)
    y = torch.empty(B * L, E, device="cuda", dtype=torch.float32)
    ops.gemm(o, sd["out_proj.weight"].to(torch.bfloat16).contiguous(), y, bias=sd["out_proj.bias"])
    assert _rel(y.reshape(B, L, E), g["y"]) < TOL


def test_arc2face_forward_vs_reference(gold, models):
    from adaprompt_b200.adaface_util import arc2face_forward_face_embs
    g = gold["arc2face_forward"]
    enc, _ = models["arc2face"]
    pe, core = arc2face_forward_face_embs(StubTokenizer(), enc, g["face_embs"].cuda(), 77)
    print(f"arc2face fwd rel-L2 {_rel(pe, g['prompt_embeds']):.3e}")
    assert _rel(pe, g["prompt_embeds"]) < TOL and torch.equal(core, pe[:, 4:20])
    pe22, _ = arc2face_forward_face_embs(StubTokenizer(), enc, g["face_embs"].cuda(), 22)
    assert _rel(pe22, gold["arc2face_forward_len22"]["prompt_embeds"]) < TOL


def _sbg(models, key="sbg"):
    from adaprompt_b200.subj_basis_generator import SubjBasisGenerator
    s = SubjBasisGenerator(num_out_embs_per_layer=16, clip_tokenizer=StubTokenizer())
    s.prompt2token_proj = models[key][0]
    return s.cuda().eval()


def test_subj_basis_generator_vs_reference(gold, models):
    core = gold["arc2face_forward"]["prompt_embeds"][:, 4:20].cuda()
    s = _sbg(models)
    subj, prompt = s(core, None, None, 1.0, True, False, "full_half_pad")
    g = gold["sbg_full_half_pad"]
    print(f"sbg subj rel-L2 {_rel(subj[:, 0], g['subj']):.3e} prompt {_rel(prompt, g['prompt']):.3e}")
    assert tuple(subj.shape) == (2, 16, 16, 768) and torch.equal(subj[:, 5], subj[:, 0])
    assert _rel(subj[:, 0], g["subj"]) < TOL and _rel(prompt, g["prompt"]) < TOL
    assert torch.allclose(s.pad_embeddings.cpu(), gold["pad_embeddings"], atol=1e-7)
    assert torch.equal(prompt[:, 22:49].cpu(), gold["pad_embeddings"][22:49].expand(2, -1, -1)) or \
        torch.allclose(prompt[:, 22:49].cpu(), gold["pad_embeddings"][22:49].expand(2, -1, -1), atol=1e-7)
    subj2, prompt2 = s(core, None, None, 0.8, True, True, "full_half_pad")
    g2 = gold["sbg_training_scale0p8"]
    assert _rel(subj2[:, 0], g2["subj"]) < TOL and _rel(prompt2[:, :24], g2["prompt_head"]) < TOL


def test_subj_basis_generator_mkv2_vs_reference(gold):
    from adaprompt_b200.subj_basis_generator import SubjBasisGenerator
    m, _ = _clip(gold["seeds"]["sbg_mkv2"], kv_mult={i: 2 for i in range(12)})
    s = SubjBasisGenerator(num_out_embs_per_layer=16, clip_tokenizer=StubTokenizer())
    s.prompt2token_proj = m
    s = s.cuda().eval()
    core = gold["arc2face_forward"]["prompt_embeds"][:, 4:20].cuda()
    subj, prompt = s(core)
    g = gold["sbg_mkv2"]
    assert _rel(subj[:, 0], g["subj"]) < TOL and _rel(prompt[:, :24], g["prompt_head"]) < TOL


class _FixedSBG(torch.nn.Module):
    """Returns given subject embeddings: isolates the splice so that it can be compared bit-exactly."""

    def __init__(self, subj):
        super().__init__()
        self.subj = subj

    def forward(self, *a, **k):
        return self.subj, None


def _manager(models, sbg):
    from adaprompt_b200.embedding_manager import EmbeddingManagerLite
    em = EmbeddingManagerLite(StubTokenizer(), arc2face_text_encoder=models["arc2face"][0])
    em.string_to_subj_basis_generator_dict["z"] = sbg
    return em.cuda().eval()


def test_splice_bit_exact_vs_reference(gold, models):
    g = gold["splice_3prompts"]
    tokens = g["tokens"].cuda()
    table = models["frozen"][1]["text_model.embeddings.token_embedding.weight"].cuda()
    em = _manager(models, _FixedSBG(g["subj_used"].cuda()[None, None].repeat(1, 16, 1, 1)))
    em.set_zs_image_features(None, gold["arc2face_forward"]["face_embs"][:1].cuda())
    static = em(tokens, table[tokens])
    assert static.shape == (48, 77, 768)
    assert torch.equal(static[g["sel"].cuda()][:, :24].cpu(), g["static_rows"])
    assert abs(float(static.double().abs().sum()) - g["static_sum"]) <= 1e-9 * g["static_sum"]
    iB, iN = em.placeholder2indices["z"]
    assert torch.equal(iB.cpu(), g["indices_B"]) and torch.equal(iN.cpu(), g["indices_N"])
    assert torch.equal(em.prompt_emb_mask.cpu(), g["prompt_emb_mask"])
    assert em.layer_copies_identical
    g2 = gold["splice_2ids"]
    from oracle import text_oracle as to
    tokens2 = torch.tensor([to.subject_prompt_ids(77)] * 2).cuda()
    em2 = _manager(models, _FixedSBG(g2["subj_used"].cuda()[:, None].repeat(1, 16, 1, 1)))
    em2.set_zs_image_features(None, gold["arc2face_forward"]["face_embs"].cuda())
    static2 = em2(tokens2, table[tokens2])
    assert torch.equal(static2[g2["sel"].cuda()][:, :24].cpu(), g2["static_rows"])


def test_get_learned_conditioning_end_to_end(gold, models):
    """C7: id embeddings -> C3 -> C1 -> C5 -> C6 -> (c, prompts, extra_info), against the reference-produced frozen
    CLIP rows (which used the reference's own SBG output) and against the full CPU oracle chain."""
    from adaprompt_b200.clip_text import FrozenCLIPEmbedder
    from adaprompt_b200.ldm_lite import LatentDiffusionLite
    from oracle import text_oracle as to
    fr = FrozenCLIPEmbedder(tokenizer=StubTokenizer())
    fr.transformer.text_model = models["frozen"][0].text_model
    fr.set_last_layers_skip_weights([0.5, 0.5])
    fr = fr.cuda().eval()
    em = _manager(models, _sbg(models))
    ldm = LatentDiffusionLite(unet=torch.nn.Identity(), cond_stage_model=fr, embedding_manager=em)
    g = gold["splice_3prompts"]
    tokens = g["tokens"].cuda()
    face = gold["arc2face_forward"]["face_embs"][:1].cuda()
    c, prompts, extra = ldm.get_learned_conditioning(tokens, None, face)
    assert c.shape == (48, 77, 768) and extra["use_layerwise_context"] and extra["use_conv_attn_kernel_size"] == -1
    f = gold["frozen_clip_rows"]
    err = _rel(c[f["sel"].cuda()][:, :32], f["c_head"])
    print(f"get_learned_conditioning rel-L2 vs reference rows {err:.3e}")
    assert err < TOL
    iB, iN = extra["placeholder2indices"]["z"]
    assert torch.equal(iB.cpu(), g["indices_B"]) and torch.equal(iN.cpu(), g["indices_N"])
    # de-duplicated encode (one of the 16 identical layer copies) == full 16x encode: bit for bit with the whole-tile
    # GEMM schedule; the split-K schedule depends on the row count, which re-associates the K sums
    from adaprompt_b200 import ops
    with ops.launch_options(split_k=1):
        c_dedup, _, _ = ldm.get_learned_conditioning(tokens, None, face)
        fr.dedup_layer_copies = False
        c_full, _, _ = ldm.get_learned_conditioning(tokens, None, face)
    assert torch.equal(c_dedup, c_full)
    assert _rel(c, c_full) < TOL          # different bf16 rounding flips through 12 layers, same accuracy class
    # config #2 of BASELINE.json: batch 32 identities -> 16 ID tokens each, spliced into 32 prompts
    face32 = torch.nn.functional.normalize(torch.randn(32, 512, generator=torch.Generator().manual_seed(11)), dim=-1).cuda()
    tokens32 = torch.tensor([to.subject_prompt_ids(77)] * 32).cuda()
    fr.dedup_layer_copies = True
    c32, _, extra32 = ldm.get_learned_conditioning(tokens32, None, face32)
    assert c32.shape == (512, 77, 768) and torch.isfinite(c32).all()
    iB, iN = extra32["placeholder2indices"]["z"]
    assert torch.equal(iB.cpu(), torch.arange(32).repeat_interleave(16))
    assert torch.equal(iN.cpu(), (5 + torch.arange(16)).repeat(32))
    # oracle chain on 2 of the 32 identities
    sds = {k: models[k][1] for k in ("arc2face", "sbg", "frozen")}
    c_ref, _, _ = to.get_learned_conditioning(sds["frozen"], sds["arc2face"], sds["sbg"], torch.tensor([[1.0], [2.0], [4.0]]),
                                              tokens32[:2].cpu(), face32[:2].cpu(), ["p"] * 2)
    err2 = _rel(c32[:32], c_ref)
    print(f"batch-32 conditioning rel-L2 vs oracle {err2:.3e}")
    assert err2 < TOL


def test_adaface_wrapper_generates_and_installs_subject_embeddings(gold, models):
    """AdaFaceWrapper surface (adaface_wrapper.py:207-254): id embedding -> 22-token Arc2Face pass -> SubjBasisGenerator
    with num_out_layers=1 -> [16, 768] -> rows z_0..z_15 of the text encoder's embedding table."""
    from adaprompt_b200.adaface_wrapper import AdaFaceWrapper
    from adaprompt_b200.clip_text import CLIPTextModelWrapper
    from oracle import text_oracle as to
    te = CLIPTextModelWrapper()
    te.load_state_dict(models["frozen"][1])
    w = AdaFaceWrapper(None, "unused", "unused", "cuda", text_encoder=te, tokenizer=StubTokenizer(),
                       subj_basis_generator=_sbg(models), arc2face_text_encoder=models["arc2face"][0])
    assert w.placeholder_token_ids == list(range(49408, 49424))
    face = gold["arc2face_forward"]["face_embs"][:1].cuda()
    embs = w.generate_adaface_embeddings(None, pre_face_embs=face)
    assert embs.shape == (16, 768)
    table = w.text_encoder.text_model.embeddings.token_embedding.weight
    assert torch.equal(table[49408:49424], embs)
    # oracle: 22-token Arc2Face pass + SBG
    sds = {k: models[k][1] for k in ("arc2face", "sbg")}
    _, core = to.arc2face_forward_face_embs(sds["arc2face"], face.cpu(), 22)
    subj, _ = to.subj_basis_generator_forward(sds["sbg"], core, torch.tensor([[1.0], [2.0], [4.0]]), num_out_layers=1)
    assert _rel(embs, subj[0, 0]) < TOL
