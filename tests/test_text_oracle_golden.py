"""CPU: the conditioning-path oracle (oracle/text_oracle.py) against golden vectors produced by the UNMODIFIED
reference code (oracle/make_golden_text.py)."""
import os

import pytest
import torch

from oracle import text_oracle as to

GOLD = os.path.join(os.path.dirname(__file__), "golden", "text_path.pt")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.fixture(scope="module")
def sds(gold):
    s = gold["seeds"]
    return {k: to.clip_synth_state_dict(s[k]) for k in ("arc2face", "sbg", "frozen")}


@pytest.mark.parametrize("mult", [1, 2, 4])
def test_mkv_attention_matches_reference_module(gold, mult):
    from adaprompt_b200.weights import synth_state_dict
    g = gold[f"mkv_attn_m{mult}"]
    spec = {}
    for n, o in (("k_proj", 768 * mult), ("v_proj", 768 * mult), ("q_proj", 768), ("out_proj", 768)):
        spec[f"{n}.weight"] = (o, 768)
        spec[f"{n}.bias"] = (o,)
    sd = synth_state_dict(spec, g["seed"])
    y = to.clip_attention(sd, "", g["x"], heads=12)
    assert _rel(y, g["y"]) < 2e-5


def test_arc2face_forward_matches_reference(gold, sds):
    g = gold["arc2face_forward"]
    pe, core = to.arc2face_forward_face_embs(sds["arc2face"], g["face_embs"], 77)
    assert _rel(pe, g["prompt_embeds"]) < 5e-5
    assert torch.equal(core, pe[:, 4:20])
    pe22, _ = to.arc2face_forward_face_embs(sds["arc2face"], g["face_embs"], 22)
    assert _rel(pe22, gold["arc2face_forward_len22"]["prompt_embeds"]) < 5e-5


def test_subj_basis_generator_matches_reference(gold, sds):
    core = gold["arc2face_forward"]["prompt_embeds"][:, 4:20]
    w = torch.tensor([[1.0], [2.0], [4.0]])
    subj, prompt = to.subj_basis_generator_forward(sds["sbg"], core, w)
    g = gold["sbg_full_half_pad"]
    assert tuple(subj.shape) == g["subj_shape"] == (2, 16, 16, 768) and g["layers_identical"]
    assert _rel(subj[:, 0], g["subj"]) < 5e-5 and torch.equal(subj[:, 3], subj[:, 0])
    assert _rel(prompt, g["prompt"]) < 5e-5
    pad = to.generate_pad_embeddings(sds["sbg"])
    assert torch.allclose(pad, gold["pad_embeddings"], atol=1e-7)
    assert torch.equal(prompt[:, 22:49], pad[22:49].expand(2, -1, -1))            # 'full_half_pad' rows are exact copies
    subj2, prompt2 = to.subj_basis_generator_forward(sds["sbg"], core, w, out_id_embs_scale=0.8, is_training=True)
    g2 = gold["sbg_training_scale0p8"]
    assert _rel(subj2[:, 0], g2["subj"]) < 5e-5 and _rel(prompt2[:, :24], g2["prompt_head"]) < 5e-5
    assert abs(float(prompt2.double().abs().sum()) - g2["prompt_sum"]) < 1e-4 * g2["prompt_sum"]
    _, prompt3 = to.subj_basis_generator_forward(sds["sbg"], core, w, adaface_prompt_embs_inf_type="full_pad")
    assert _rel(prompt3[:, 20:30], gold["sbg_full_pad"]["prompt_20_30"]) < 5e-5


def test_subj_basis_generator_mkv2_matches_reference(gold):
    sd = to.clip_synth_state_dict(gold["seeds"]["sbg_mkv2"], kv_mult={i: 2 for i in range(12)})
    core = gold["arc2face_forward"]["prompt_embeds"][:, 4:20]
    subj, prompt = to.subj_basis_generator_forward(sd, core, torch.tensor([[1.0], [2.0], [4.0]]))
    g = gold["sbg_mkv2"]
    assert _rel(subj[:, 0], g["subj"]) < 5e-5 and _rel(prompt[:, :24], g["prompt_head"]) < 5e-5


def test_splice_is_bit_exact_and_indices_match_reference(gold, sds):
    g = gold["splice_3prompts"]
    tokens = g["tokens"]
    embedded = sds["frozen"]["text_model.embeddings.token_embedding.weight"][tokens]
    assert g["subj_layers_identical"]
    subj = g["subj_used"][None, None].repeat(1, 16, 1, 1)                          # [BS=1, 16, K, D]
    static, indices, mask = to.splice_subject_embeddings(tokens, embedded, subj)
    assert static.shape == (48, 77, 768)
    assert torch.equal(static[g["sel"]][:, :24], g["static_rows"])                 # bit-exact rows
    assert abs(float(static.double().abs().sum()) - g["static_sum"]) <= 1e-9 * g["static_sum"]
    assert torch.equal(indices[0], g["indices_B"]) and torch.equal(indices[1], g["indices_N"])
    assert indices[0].tolist() == [0] * 16 + [1] * 16
    assert indices[1].tolist() == list(range(5, 21)) + list(range(3, 19))          # SURVEY.md 8(c) known answer
    assert torch.equal(mask, g["prompt_emb_mask"]) and mask.sum(1).flatten().tolist() == [20.0, 18.0, 2.0]
    g2 = gold["splice_2ids"]
    tokens2 = torch.tensor([to.subject_prompt_ids(77)] * 2)
    emb2 = sds["frozen"]["text_model.embeddings.token_embedding.weight"][tokens2]
    subj2 = g2["subj_used"][:, None].repeat(1, 16, 1, 1)                           # [BS=2, 16, K, D]
    static2, _, _ = to.splice_subject_embeddings(tokens2, emb2, subj2)
    assert torch.equal(static2[g2["sel"]][:, :24], g2["static_rows"])


def test_frozen_clip_matches_reference(gold, sds):
    g = gold["splice_3prompts"]
    tokens = g["tokens"]
    embedded = sds["frozen"]["text_model.embeddings.token_embedding.weight"][tokens]
    subj = g["subj_used"][None, None].repeat(1, 16, 1, 1)
    static, _, _ = to.splice_subject_embeddings(tokens, embedded, subj)
    c = to.frozen_clip_encode(sds["frozen"], tokens, static)
    f = gold["frozen_clip_rows"]
    assert _rel(c[f["sel"]][:, :32], f["c_head"]) < 5e-5
    assert torch.equal(c[0], c[7])                                                 # the 16 layer copies are identical (N2)
