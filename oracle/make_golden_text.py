"""TEST INFRASTRUCTURE - generates tests/golden/text_*.pt by running the UNMODIFIED reference conditioning code
(/root/reference, build container only) on the synthetic-weight recipe.  Run:

    python oracle/make_golden_text.py

What is executed from the reference, as is:
  * adaface.util.arc2face_forward_face_embs / arc2face_inverse_face_prompt_embs   (C3, C2)
  * adaface.subj_basis_generator.SubjBasisGenerator.forward / generate_pad_embeddings (C1; the instance is built
    with object.__new__ because __init__ downloads pretrained CLIP - SURVEY.md section 8(c))
  * ldm.modules.embedding_manager.EmbeddingManager.forward / get_static_embedding / update_placeholder_indices /
    update_prompt_masks and StaticLayerwiseEmbedding.forward (C5; instance built the same way)
  * adaface.arc2face_models.CLIPAttentionMKV.forward and extend_clip_attention_MKV_multiplier (C4, MKV)
  * transformers' CLIPTextEmbeddings / CLIPEncoderLayer / final LayerNorm modules of a reference
    CLIPTextModelWrapper instance, driven layer by layer with an explicit causal mask by `ByHandTextEncoder`
    below, which follows CLIPTextModelWrapper.forward (arc2face_models.py:204-248) because that forward itself is
    broken under the installed transformers 5.5.0 (causal mask dropped; hidden_states indexing raises).
The CLIP tokenizer has no vocabulary files offline: `StubTokenizer` returns the fixed id rows of the three prompt
templates the path uses (ids from oracle/text_oracle.py).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle import text_oracle as to  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED_ARC2FACE, SEED_SBG, SEED_FROZEN = 2101, 2102, 2103


def import_text_reference():
    rh.import_reference()
    import ldm.util  # noqa: F401  (must be fully imported before adaface.* aliases sys.modules['ldm'])
    ldm_pkg = sys.modules["ldm"]
    importlib.import_module("ldm.modules.embedding_manager")
    sys.modules["ldm"] = ldm_pkg
    return types.SimpleNamespace(em=sys.modules["ldm.modules.embedding_manager"], au=sys.modules["adaface.util"],
                                 am=sys.modules["adaface.arc2face_models"], sbg=sys.modules["adaface.subj_basis_generator"],
                                 lu=sys.modules["ldm.util"])


class StubTokenizer:
    """Fixed-vocabulary stand-in for CLIPTokenizer (no vocab files offline)."""
    pad_token_id = to.PAD
    VOCAB = {"photo": to.TOK_PHOTO, "of": to.TOK_OF, "a": to.TOK_A, "id": to.TOK_ID, "person": to.TOK_PERSON,
             ",": to.TOK_COMMA, "z": to.TOK_Z, "y": to.TOK_Y}

    def _ids(self, text):
        return [self.VOCAB[w] for w in text.replace(",", " , ").split()]

    def encode(self, text, add_special_tokens=False):
        return self._ids(text)

    def __call__(self, text, truncation=True, padding="max_length", max_length=77, return_tensors="pt", **kw):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [to.pad_ids(self._ids(t), max_length) for t in texts]
        return types.SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.long))


def build_ref_clip(ns, seed, num_layers=12, mkv=None):
    """Reference CLIPTextModelWrapper (random-init architecture) loaded with the synthetic recipe; optional MKV
    extension through the reference's own extend_clip_attention_MKV_multiplier (noise-free so the weights stay a
    pure function of the recipe: the extended k/v weights are then re-loaded from the recipe's widened tensors)."""
    from transformers import CLIPTextConfig
    cfg = CLIPTextConfig(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=num_layers,
                         vocab_size=49408, max_position_embeddings=77, hidden_act="quick_gelu")
    m = ns.am.CLIPTextModelWrapper(cfg).eval()
    if mkv:
        m.extend_clip_attention_MKV_multiplier(-1, -1, mkv, noise_std=0)
    sd = to.clip_synth_state_dict(seed, num_layers=num_layers, kv_mult={i: mkv for i in range(num_layers)} if mkv else None)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    assert not unexpected, unexpected
    return m, sd


class ByHandTextEncoder:
    """Call protocol of CLIPTextModelWrapper.forward (arc2face_models.py:178-248) on the reference instance's own
    sub-modules, with the causal mask applied explicitly."""

    def __init__(self, m):
        self.m = m
        self.text_model = m.text_model
        self.config = m.config
        self.dtype = torch.float32

    def __call__(self, input_ids=None, input_token_embs=None, hidden_state_layer_weights=None,
                 return_token_embs=False, **kw):
        tm = self.m.text_model
        if return_token_embs:
            return tm.embeddings.token_embedding(input_ids)
        h = tm.embeddings(input_ids=input_ids, inputs_embeds=input_token_embs)                    # :210
        B, L = input_ids.shape
        causal = torch.full((L, L), torch.finfo(h.dtype).min).triu(1)[None, None].expand(B, 1, L, L)  # :214
        states = [h]
        for layer in tm.encoder.layers:
            h = layer(h, causal)
            h = h[0] if isinstance(h, tuple) else h
            states.append(h)
        if hidden_state_layer_weights is not None:                                                 # :236-246
            n = len(hidden_state_layer_weights)
            w = hidden_state_layer_weights.to(h.dtype)
            w = w / w.sum(dim=0, keepdim=True)
            w = w.unsqueeze(1).unsqueeze(1)
            h = (torch.stack(states[-n:], dim=0) * w).sum(dim=0)
        return (tm.final_layer_norm(h),)                                                            # :248


def build_ref_sbg(ns, enc, tok):
    s = object.__new__(ns.sbg.SubjBasisGenerator)
    nn.Module.__init__(s)
    s.placeholder_is_bg = False
    s.num_out_layers = 16
    s.num_out_embs_per_layer = 16
    s.num_out_embs = 256
    s.output_dim = 768
    s.zs_extra_words_scale = 0.5
    s.clip_tokenizer = tok
    s.__dict__["prompt2token_proj"] = enc
    s.prompt2token_proj_grad_scale = 0.4
    s.prompt2token_proj_grad_scaler = ns.au.gen_gradient_scaler(0.4)
    s.prompt2token_proj_attention_multiplier = -1
    s.initialize_hidden_state_layer_weights("per-layer", "cpu")
    s.pad_embeddings = None
    s.eval()
    return s


def build_ref_embedding_manager(ns, sbg_obj, arc_enc, tok):
    e = object.__new__(ns.em.EmbeddingManager)
    nn.Module.__init__(e)
    se = object.__new__(ns.em.StaticLayerwiseEmbedding)
    nn.Module.__init__(se)
    se.do_zero_shot, se.device_type = True, "cpu"
    se.basis_rand_weights = torch.zeros(1)
    se.basis_comm_weights = torch.zeros(1)
    e.string_to_token_dict = {"z": to.TOK_Z}
    e.background_string_dict = {}
    e.placeholder_strings = ["z"]
    e.string_to_static_embedder_dict = nn.ModuleDict({"z": se})
    e.__dict__["string_to_subj_basis_generator_dict"] = {"z": sbg_obj}
    e.token2num_vectors = {"z": 16}
    e.use_layerwise_embedding = True
    e.num_unet_ca_layers = 16
    e.num_layers_per_embedder = 16
    e.do_zero_shot = True
    e.curr_subj_is_face = True
    e.__dict__["arc2face_text_encoder"] = arc_enc
    e.tokenizer = tok
    e.zs_out_id_embs_scale_range = (1.0, 1.0)
    e.zs_adaface_prompt_embs_inf_type = "full_half_pad"
    e.iter_type = None
    e.CLS_DELTA_STRING_MAX_SEARCH_SPAN = 0
    e.current_subj_name_to_cls_delta_tokens = {}
    e.training_begin_add_noise_std_range = None
    e.zs_image_feat_dict = {}
    e.eval()
    return e


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    os.makedirs(OUT, exist_ok=True)
    ns = import_text_reference()
    tok = StubTokenizer()
    out = {}
    with torch.no_grad():
        # ---- MKV attention module: reference CLIPAttentionMKV vs restated attention ---------------------------------
        from transformers import CLIPTextConfig
        cfg = CLIPTextConfig(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=1)
        g = torch.Generator().manual_seed(5)
        for mult in (1, 2, 4):
            att = ns.am.CLIPAttentionMKV(cfg, mult).eval()
            from adaprompt_b200.weights import spec_of, synth_state_dict
            att.load_state_dict(synth_state_dict(spec_of(att), 77 + mult))
            x = torch.randn(2, 22, 768, generator=g)
            causal = torch.full((22, 22), torch.finfo(torch.float32).min).triu(1)[None, None].expand(2, 1, 22, 22)
            y, _ = att(x, None, causal)
            out[f"mkv_attn_m{mult}"] = {"x": x, "y": y, "seed": 77 + mult}
            print(f"mkv m={mult}: |y|={y.abs().mean():.4f}")

        # ---- C3 / C1 / C2 / C5 / C6 chain ----------------------------------------------------------------------------
        arc_m, arc_sd = build_ref_clip(ns, SEED_ARC2FACE)
        sbg_m, sbg_sd = build_ref_clip(ns, SEED_SBG)
        arc_enc, sbg_enc = ByHandTextEncoder(arc_m), ByHandTextEncoder(sbg_m)
        # causality sanity check of the by-hand driver
        ids = torch.tensor([to.inverse_template_ids(77)])
        e0 = sbg_enc(input_ids=ids)[0]
        ids2 = ids.clone(); ids2[0, 30] = 1234
        e1 = sbg_enc(input_ids=ids2)[0]
        assert torch.equal(e0[0, :30], e1[0, :30]) and not torch.equal(e0[0, 30:], e1[0, 30:]), "driver is not causal"

        face = torch.nn.functional.normalize(torch.randn(2, 512, generator=torch.Generator().manual_seed(11)), dim=-1)
        pe, core = ns.au.arc2face_forward_face_embs(tok, arc_enc, face, input_max_length=77)
        out["arc2face_forward"] = {"face_embs": face, "prompt_embeds": pe.clone()}  # core == prompt_embeds[:, 4:20]
        assert torch.equal(core, pe[:, 4:20])
        pe22, core22 = ns.au.arc2face_forward_face_embs(tok, arc_enc, face, input_max_length=22)
        out["arc2face_forward_len22"] = {"prompt_embeds": pe22.clone()}
        print(f"arc2face fwd: |core|={core.abs().mean():.4f}")

        sbg_obj = build_ref_sbg(ns, sbg_enc, tok)
        subj, prompt = sbg_obj(core, None, None, 1.0, True, False, "full_half_pad")
        out["sbg_full_half_pad"] = {"subj": subj[:, 0].clone(), "subj_shape": tuple(subj.shape), "prompt": prompt.clone(),
                                    "layers_identical": bool((subj == subj[:, :1]).all())}
        subj2, prompt2 = sbg_obj(core, None, None, 0.8, True, True, "full_half_pad")
        out["sbg_training_scale0p8"] = {"subj": subj2[:, 0].clone(), "prompt_head": prompt2[:, :24].clone(),
                                        "prompt_sum": to_sum(prompt2)}
        subj3, prompt3 = sbg_obj(core, None, None, 1.0, True, False, "full_pad")
        out["sbg_full_pad"] = {"prompt_20_30": prompt3[:, 20:30].clone(), "prompt_sum": to_sum(prompt3)}
        out["pad_embeddings"] = sbg_obj.pad_embeddings.clone()
        print(f"sbg: subj {tuple(subj.shape)} |subj|={subj.abs().mean():.4f}")

        # MKV x2 prompt2token_proj (README.md:72 --extend_prompt2token_proj_attention_multiplier)
        mkv_m, _ = build_ref_clip(ns, SEED_SBG + 7, mkv=2)
        sbg_mkv = build_ref_sbg(ns, ByHandTextEncoder(mkv_m), tok)
        subj_m, prompt_m = sbg_mkv(core, None, None, 1.0, True, False, "full_half_pad")
        out["sbg_mkv2"] = {"subj": subj_m[:, 0].clone(), "prompt_head": prompt_m[:, :24].clone(), "prompt_sum": to_sum(prompt_m)}

        # C5: EmbeddingManager.forward on 3 prompts (two with the placeholder at different positions, one without)
        frozen_m, frozen_sd = build_ref_clip(ns, SEED_FROZEN)
        em_obj = build_ref_embedding_manager(ns, sbg_obj, arc_enc, tok)
        rows = [to.subject_prompt_ids(77),
                to.pad_ids([to.TOK_A, to.TOK_PERSON, to.TOK_Z] + [to.TOK_COMMA] * 15, 77),
                to.pad_ids([to.TOK_A, to.TOK_PHOTO], 77)]
        tokens = torch.tensor(rows)
        embedded = frozen_m.text_model.embeddings.token_embedding(tokens)
        em_obj.set_zs_image_features(torch.zeros(1, 514, 8), face[:1], (1.0, 1.0))
        static = em_obj(tokens, embedded)
        p2i = em_obj.placeholder2indices["z"]
        sel = torch.tensor([0, 1, 15, 16, 31, 32, 47])
        out["splice_3prompts"] = {"tokens": tokens, "sel": sel, "static_rows": static[sel][:, :24].clone(), "static_sum": to_sum(static),
                                  "subj_used": em_obj.static_subj_embs_dict["z"][0].clone(),   # [K, D]: layer copies are identical
                                  "subj_layers_identical": bool((em_obj.static_subj_embs_dict["z"] == em_obj.static_subj_embs_dict["z"][:1]).all()),
                                  "indices_B": p2i[0].clone(), "indices_N": p2i[1].clone(),
                                  "prompt_emb_mask": em_obj.prompt_emb_mask.clone()}
        print("splice:", tuple(static.shape), p2i[0].tolist()[:18], p2i[1].tolist()[:18],
              em_obj.prompt_emb_mask.sum(1).flatten().tolist())
        # batch of 2 identities x same prompt (config 2 style, small)
        tokens2 = torch.tensor([to.subject_prompt_ids(77)] * 2)
        em_obj.set_zs_image_features(torch.zeros(2, 514, 8), face, (1.0, 1.0))
        static2 = em_obj(tokens2, frozen_m.text_model.embeddings.token_embedding(tokens2))
        sel2 = torch.tensor([0, 15, 16, 31])
        out["splice_2ids"] = {"sel": sel2, "static_rows": static2[sel2][:, :24].clone(), "static_sum": to_sum(static2),
                              "subj_used": em_obj.static_subj_embs_dict["z"][[0, 16]].clone()}

        # C6: frozen CLIP on the spliced embeddings (by-hand driver + skip weights 0.5/0.5), first layer copies only
        fr_enc = ByHandTextEncoder(frozen_m)
        ids16 = tokens.unsqueeze(1).repeat(1, 16, 1).view(48, 77)
        c = fr_enc(input_ids=ids16[sel], input_token_embs=static[sel], hidden_state_layer_weights=torch.tensor([[0.5], [0.5]]))[0]
        out["frozen_clip_rows"] = {"sel": sel, "c_head": c[:, :32].clone(), "c_sum": to_sum(c)}
        print(f"frozen clip: |c|={c.abs().mean():.4f}")
    out["seeds"] = {"arc2face": SEED_ARC2FACE, "sbg": SEED_SBG, "sbg_mkv2": SEED_SBG + 7, "frozen": SEED_FROZEN}
    # keep the fixture small: fp32 -> store as is (about 3 MB)
    torch.save(out, os.path.join(OUT, "text_path.pt"))
    print("saved", os.path.getsize(os.path.join(OUT, "text_path.pt")) / 1e6, "MB")


def to_sum(t):
    return float(t.double().abs().sum())


if __name__ == "__main__":
    main()
