"""TEST INFRASTRUCTURE - CPU oracle of the AdaFace conditioning path, fp32, plain torch.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file, and only as the
checker.  The product path (adaprompt_b200/*) never does.

Restatement (not a copy), each function citing the reference lines it follows.  Pinning status:
  * C1 SubjBasisGenerator.forward, C2 arc2face_inverse_face_prompt_embs, C3 arc2face_forward_face_embs, C5
    EmbeddingManager splice / placeholder indices / prompt mask: pinned against the UNMODIFIED reference
    functions executed in the build container (oracle/make_golden_text.py -> tests/golden/text_*.pt).
  * CLIPAttentionMKV (adaface/arc2face_models.py:87-173): pinned against the reference module itself.
  * C4 / C6 CLIP text transformer arithmetic: the algorithm lives in the un-vendored third-party dependency
    HuggingFace `transformers` (requirements.txt:14 `transformers>=4.32.0`, code comments cite v4.34.1).  The
    reference wrappers (arc2face_models.py:178-280, ldm/modules/encoders/modules.py:234-371) silently break under
    the installed transformers 5.5.0 (the causal mask is dropped - SURVEY.md section 8(c)), so they cannot be run as
    the oracle; the published CLIP text-encoder algorithm is restated below (pre-LN layers, causal mask,
    quick_gelu MLP, weighted sum of the last hidden states before the final LayerNorm) and pinned on the
    installed transformers' own CLIPEncoderLayer / CLIPTextEmbeddings modules driven layer by layer with an
    explicit causal mask, following the reference call sites arc2face_models.py:204-248.  End-to-end behaviour of
    the reference wrapper classes under their pinned transformers version is therefore "parity unpinned".
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# token ids the reference hard-codes or documents (embedding_manager.py:1062,1645-1646; adaface/util.py:87)
BOS, EOS, PAD = 49406, 49407, 49407
TOK_COMMA, TOK_Z, TOK_Y, TOK_ID = 267, 345, 344, 1014
N_CA_LAYERS = 16


# ---------------------------------------------------------------------------------------------------
# CLIP text transformer (HF CLIPTextModel as driven by arc2face_models.py:204-248 / modules.py:260-283,361-370)
# ---------------------------------------------------------------------------------------------------
def clip_state_spec(num_layers=12, hidden=768, inter=3072, vocab=49408, max_pos=77, kv_mult: Optional[Dict[int, int]] = None):
    """key -> shape with HF names (text_model.*).  kv_mult: {layer: multiplier} for CLIPAttentionMKV layers."""
    from collections import OrderedDict
    spec = OrderedDict()
    spec["text_model.embeddings.token_embedding.weight"] = (vocab, hidden)
    spec["text_model.embeddings.position_embedding.weight"] = (max_pos, hidden)
    for i in range(num_layers):
        p = f"text_model.encoder.layers.{i}."
        m = (kv_mult or {}).get(i, 1)
        for n, o in (("k_proj", hidden * m), ("v_proj", hidden * m), ("q_proj", hidden), ("out_proj", hidden)):
            spec[p + f"self_attn.{n}.weight"] = (o, hidden)
            spec[p + f"self_attn.{n}.bias"] = (o,)
        spec[p + "layer_norm1.weight"] = (hidden,)
        spec[p + "layer_norm1.bias"] = (hidden,)
        spec[p + "mlp.fc1.weight"] = (inter, hidden)
        spec[p + "mlp.fc1.bias"] = (inter,)
        spec[p + "mlp.fc2.weight"] = (hidden, inter)
        spec[p + "mlp.fc2.bias"] = (hidden,)
        spec[p + "layer_norm2.weight"] = (hidden,)
        spec[p + "layer_norm2.bias"] = (hidden,)
    spec["text_model.final_layer_norm.weight"] = (hidden,)
    spec["text_model.final_layer_norm.bias"] = (hidden,)
    return spec


def clip_synth_state_dict(seed: int, **kw) -> SD:
    """Synthetic CLIP-text weights: the shared recipe, with embeddings ~ N(0, 0.02) like the HF init
    (SURVEY.md section 8(d) config 2)."""
    from adaprompt_b200.weights import _gen, synth_state_dict
    spec = clip_state_spec(**kw)
    sd = synth_state_dict(spec, seed)
    for k in ("text_model.embeddings.token_embedding.weight", "text_model.embeddings.position_embedding.weight"):
        sd[k] = torch.randn(spec[k], generator=_gen(k, seed), dtype=torch.float32) * 0.02
    return sd


def num_layers_of(sd: SD, prefix="text_model.") -> int:
    n = 0
    while f"{prefix}encoder.layers.{n}.layer_norm1.weight" in sd:
        n += 1
    return n


def clip_attention(sd: SD, p: str, x: torch.Tensor, heads: int, causal: bool = True) -> torch.Tensor:
    """HF CLIPAttention / CLIPAttentionMKV (arc2face_models.py:87-173): q scaled by head_dim^-1/2; with a KV
    multiplier m, token t contributes m keys/values (columns [r*E,(r+1)*E) of the widened k/v projections), all
    masked like token t (:117-142)."""
    B, L, E = x.shape
    hd = E // heads
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * hd ** -0.5
    k = F.linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    m = k.shape[-1] // E
    q = q.view(B, L, heads, hd).transpose(1, 2)                       # [B, H, L, hd]
    k = k.view(B, L * m, heads, hd).transpose(1, 2)                   # key index = t*m + r
    v = v.view(B, L * m, heads, hd).transpose(1, 2)
    w = q @ k.transpose(-1, -2)                                       # [B, H, L, L*m]
    if causal:
        tok = torch.arange(L * m) // m
        mask = tok[None, :] > torch.arange(L)[:, None]
        w = w.masked_fill(mask, torch.finfo(w.dtype).min)
    a = w.softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, L, E)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def clip_layer(sd: SD, p: str, h: torch.Tensor, heads: int) -> torch.Tensor:
    """HF CLIPEncoderLayer: pre-LN attention and quick_gelu MLP, both residual."""
    E = h.shape[-1]
    x = F.layer_norm(h, (E,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], 1e-5)
    h = h + clip_attention(sd, p + "self_attn.", x, heads)
    x = F.layer_norm(h, (E,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], 1e-5)
    x = F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    x = x * torch.sigmoid(1.702 * x)                                  # QuickGELUActivation
    return h + F.linear(x, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def clip_text_forward(sd: SD, input_ids: torch.Tensor, input_token_embs: Optional[torch.Tensor] = None,
                      hidden_state_layer_weights: Optional[torch.Tensor] = None, heads: int = 12,
                      prefix: str = "text_model.") -> torch.Tensor:
    """CLIPTextModelWrapper.forward -> last_hidden_state (arc2face_models.py:204-248): embeddings (token or the
    given token embeddings, + position) -> causal encoder -> [normalised weighted sum of the last n hidden states
    (:236-246)] -> final LayerNorm (:248)."""
    tok_w = sd[prefix + "embeddings.token_embedding.weight"]
    pos_w = sd[prefix + "embeddings.position_embedding.weight"]
    L = input_ids.shape[-1]
    tok = tok_w[input_ids] if input_token_embs is None else input_token_embs
    h = tok + pos_w[:L]
    states = [h]
    for i in range(num_layers_of(sd, prefix)):
        h = clip_layer(sd, f"{prefix}encoder.layers.{i}.", h, heads)
        states.append(h)
    if hidden_state_layer_weights is not None:
        w = hidden_state_layer_weights.to(h.dtype)
        w = w / w.sum(dim=0, keepdim=True)
        w = w.unsqueeze(1).unsqueeze(1)
        h = (torch.stack(states[-w.shape[0]:], dim=0) * w).sum(dim=0)
    E = h.shape[-1]
    return F.layer_norm(h, (E,), sd[prefix + "final_layer_norm.weight"], sd[prefix + "final_layer_norm.bias"], 1e-5)


# ---------------------------------------------------------------------------------------------------
# fixed token rows (stub of the CLIP tokenizer for the three templates the path uses)
# ---------------------------------------------------------------------------------------------------
def pad_ids(ids: Sequence[int], length: int = 77) -> List[int]:
    row = [BOS] + list(ids) + [EOS]
    return row + [PAD] * (length - len(row))


# "photo", "of", "a", "person": ids of the openai/clip-vit-large-patch14 vocabulary; any fixed distinct ids give
# the same arithmetic, the stub only has to be the same on both sides of a comparison.
TOK_PHOTO, TOK_OF, TOK_A, TOK_PERSON = 1125, 539, 320, 2533


def arc2face_template_ids(length: int = 77) -> List[int]:
    """'photo of a id person' (adaface/util.py:90-96): 'id' sits at position 4."""
    return pad_ids([TOK_PHOTO, TOK_OF, TOK_A, TOK_ID, TOK_PERSON], length)


def inverse_template_ids(length: int = 77) -> List[int]:
    """'photo of a ' + ', ' * 16 (adaface/util.py:165): 16 commas at positions 4..19."""
    return pad_ids([TOK_PHOTO, TOK_OF, TOK_A] + [TOK_COMMA] * 16, length)


def subject_prompt_ids(length: int = 77, placeholder: int = TOK_Z) -> List[int]:
    """'a photo of a z' + ', ' * 15 (personalized.py:885-894 padding): z at position 5."""
    return pad_ids([TOK_A, TOK_PHOTO, TOK_OF, TOK_A, placeholder] + [TOK_COMMA] * 15, length)


# ---------------------------------------------------------------------------------------------------
# C3 / C2 / C1
# ---------------------------------------------------------------------------------------------------
def arc2face_forward_face_embs(sd: SD, face_embs: torch.Tensor, input_max_length: int = 77):
    """adaface/util.py:76-125: pad the 512-d ArcFace embedding to 768, put it at the 'id' token, CLIP pass;
    returns (prompt_embeds [N,L,768], core [N,16,768] = rows 4:20)."""
    N = face_embs.shape[0]
    ids = torch.tensor(arc2face_template_ids(input_max_length)).repeat(N, 1)
    tok = sd["text_model.embeddings.token_embedding.weight"][ids].clone()
    padded = F.pad(face_embs, (0, tok.shape[-1] - face_embs.shape[-1]))
    tok[ids == TOK_ID] = padded
    pe = clip_text_forward(sd, ids, tok)
    return pe, pe[:, 4:20]


def generate_pad_embeddings(sd: SD) -> torch.Tensor:
    """subj_basis_generator.py:587-602: token + position embedding of 77 pad tokens."""
    ids = torch.full((77,), PAD)
    return sd["text_model.embeddings.token_embedding.weight"][ids] + sd["text_model.embeddings.position_embedding.weight"][:77]


def arc2face_inverse_face_prompt_embs(sd: SD, face_prompt_embs: torch.Tensor, return_emb_types: Sequence[str],
                                      pad_embeddings: torch.Tensor, hidden_state_layer_weights=None,
                                      input_max_length: int = 77):
    """adaface/util.py:132-238 with list_extra_words=None."""
    BS = face_prompt_embs.shape[0]
    ids = torch.tensor(inverse_template_ids(input_max_length)).repeat(BS, 1)
    tok = sd["text_model.embeddings.token_embedding.weight"][ids].clone()
    tok[:, 4:20] = face_prompt_embs
    pe = clip_text_forward(sd, ids, tok, hidden_state_layer_weights)
    core = pe[:, 4:20]
    out = []
    for t in return_emb_types:
        if t == "full":
            out.append(pe)
        elif t == "full_half_pad":
            p2 = pe.clone()
            pads = p2.shape[1] - 23
            if pads >= 2:
                p2[:, 22:22 + pads // 2] = pad_embeddings[22:22 + pads // 2]
            out.append(p2)
        elif t == "full_pad":
            p2 = pe.clone()
            p2[:, 22:-1] = pad_embeddings[22:-1]
            out.append(p2)
        elif t == "core":
            out.append(core)
        elif t == "full_zeroed_extra":
            p2 = pe.clone()
            p2[:, 22:24] = pad_embeddings[22:24]
            p2[:, 24:-1] = 0
            out.append(p2)
        elif t == "b_core_e":
            out.append(torch.cat([pe[:, :22], pe[:, [-1]]], dim=1))
        else:
            raise ValueError(t)
    return out


def subj_basis_generator_forward(sd: SD, arc2face_id_embs: torch.Tensor, hidden_state_layer_weights: torch.Tensor,
                                 out_id_embs_scale: float = 1.0, is_training: bool = False,
                                 adaface_prompt_embs_inf_type: str = "full_half_pad", num_out_layers: int = 16,
                                 num_out_embs_per_layer: int = 16):
    """SubjBasisGenerator.forward, face branch (subj_basis_generator.py:470-567)."""
    types = ["full_pad", "core"] if is_training else [adaface_prompt_embs_inf_type, "core"]
    pad = generate_pad_embeddings(sd)
    prompt_embs, core = arc2face_inverse_face_prompt_embs(sd, arc2face_id_embs, types, pad, hidden_state_layer_weights)
    subj = core.unsqueeze(1).repeat(1, num_out_layers, 1, 1)
    if out_id_embs_scale != 1:
        pe = pad[4:4 + num_out_embs_per_layer].unsqueeze(0).unsqueeze(0)
        subj = subj * out_id_embs_scale + pe * (1 - out_id_embs_scale)
    return subj, prompt_embs


# ---------------------------------------------------------------------------------------------------
# C5 splice (embedding_manager.py:1292-1588, 1646-1648, 1699-1722; ldm/util.py:1874-1886)
# ---------------------------------------------------------------------------------------------------
def first_index_in_each_instance(rows: torch.Tensor, cols: torch.Tensor):
    """ldm/util.py:1874-1886 (indices come sorted by row from torch.where)."""
    keep = torch.ones_like(rows, dtype=torch.bool)
    keep[1:] = rows[1:] != rows[:-1]
    return rows[keep], cols[keep]


def splice_subject_embeddings(tokenized_text: torch.Tensor, embedded_text: torch.Tensor, adaface_subj_embs: torch.Tensor,
                              placeholder_token: int = TOK_Z, num_vectors: int = 16):
    """EmbeddingManager.forward for one zero-shot face placeholder.  tokenized_text [B,N]; embedded_text [B,N,D];
    adaface_subj_embs [BS,16,K,D] -> (static_embedded_text [16B,N,D], placeholder_indices (B_idx, N_idx) or None,
    prompt_emb_mask [B,N,1])."""
    B, N = tokenized_text.shape
    emb = embedded_text.clone().unsqueeze(1).repeat(1, N_CA_LAYERS, 1, 1).view(B * N_CA_LAYERS, N, -1)     # :1349
    tok = tokenized_text.unsqueeze(1).repeat(1, N_CA_LAYERS, 1).view(B * N_CA_LAYERS, N)                  # :1353
    rows, cols = torch.where(tok == placeholder_token)                                                    # :1359
    if rows.numel() > 0:
        r1, c1 = first_index_in_each_instance(rows, cols)                                                 # :1368
        occurs = r1.numel() // N_CA_LAYERS                                                                # :1383
        subj = adaface_subj_embs
        if subj.shape[0] < occurs:
            subj = subj.repeat(occurs // subj.shape[0], 1, 1, 1)                                          # :1449-1451
        static = subj.reshape(-1, subj.shape[2], subj.shape[3])                                           # 'b l k d -> (b l) k d' :511
        for k in range(num_vectors):
            sk = static[:, k]
            if sk.shape[0] == N_CA_LAYERS:
                sk = sk.repeat(occurs, 1)                                                                 # :1553
            emb[(r1, c1 + k)] = sk                                                                        # :1561-1562
    # update_placeholder_indices on the un-repeated tokens (:1699-1722)
    rows, cols = torch.where(tokenized_text == placeholder_token)
    if rows.numel() == 0:
        indices = None
    else:
        rb, cb = first_index_in_each_instance(rows, cols)
        bs = rb.shape[0]
        rb = rb.unsqueeze(1).repeat(1, num_vectors).view(-1)
        cb = cb.unsqueeze(1).repeat(1, num_vectors).view(-1) + torch.arange(num_vectors).repeat(bs)
        indices = (rb, cb)
    mask = ((tokenized_text != BOS) & (tokenized_text != EOS)).float().unsqueeze(2)                      # :1646-1648
    return emb, indices, mask


# ---------------------------------------------------------------------------------------------------
# C6 FrozenCLIPEmbedder (modules.py:195-223, 260-283, 361-370) and C7 conditioning tuple (ddpm.py:1065-1078)
# ---------------------------------------------------------------------------------------------------
def frozen_clip_encode(sd: SD, tokenized_text: torch.Tensor, static_embedded_text: torch.Tensor,
                       last_layers_skip_weights=(0.5, 0.5)) -> torch.Tensor:
    """tokens repeated x16 with the spliced token embeddings -> + position -> causal encoder -> normalised
    weighted sum of the last len(w) hidden states -> final LayerNorm."""
    B16 = static_embedded_text.shape[0]
    ids = tokenized_text.unsqueeze(1).repeat(1, B16 // tokenized_text.shape[0], 1).view(B16, -1)
    w = torch.tensor(last_layers_skip_weights, dtype=torch.float32).unsqueeze(1)
    return clip_text_forward(sd, ids, static_embedded_text, w)


def get_learned_conditioning(sd_frozen: SD, sd_arc2face: SD, sd_sbg: SD, layer_weights: torch.Tensor,
                             tokenized_text: torch.Tensor, zs_id_embs: torch.Tensor, prompts: List[str]):
    """ddpm.py:970-1085 for the zero-shot face path: id embs -> C3 -> C1 -> C5 -> C6 -> (c, prompts, extra_info)."""
    _, id_embs = arc2face_forward_face_embs(sd_arc2face, zs_id_embs)
    subj, _ = subj_basis_generator_forward(sd_sbg, id_embs, layer_weights)
    embedded = sd_frozen["text_model.embeddings.token_embedding.weight"][tokenized_text]
    static, indices, mask = splice_subject_embeddings(tokenized_text, embedded, subj)
    c = frozen_clip_encode(sd_frozen, tokenized_text, static)
    extra_info = {"use_layerwise_context": True, "use_conv_attn_kernel_size": -1,
                  "placeholder2indices": {"z": indices}, "prompt_emb_mask": mask, "is_training": False,
                  "compel_cfg_weight_level_range": None, "apply_compel_cfg_prob": 0, "empty_context": None,
                  "capture_distill_attn": False}
    return c, prompts, extra_info
