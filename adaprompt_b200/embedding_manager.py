"""Splice slice of ldm/modules/embedding_manager.py:EmbeddingManager (reference: askerlee/adaprompt) for the zero-shot
face path: forward :1292-1326, get_static_embedding :1329-1588 (live inference lines), update_placeholder_indices
:1699-1722, update_prompt_masks :1646-1648, set_zs_image_features :1790-1817, StaticLayerwiseEmbedding.forward :502-516.

Replaces placeholder-token rows of the CLIP token embeddings with the AdaFace ID tokens, expanding the batch x16 (one
copy per UNet cross-attention layer, layer index minor to batch) and records placeholder2indices / prompt_emb_mask.
Index work runs on the GPU (af_find_first_token) and the row replacement is an exact copy (af_splice_rows): bit-exact.
Training-only features (embedding noise, cls-delta strings, frozen-generator mixing, background tokens) are out of
scope and raise when requested.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import ops
from .adaface_util import arc2face_forward_face_embs

N_CA_LAYERS = 16


class StaticLayerwiseEmbedding(nn.Module):
    """Zero-shot mode is a pure rearrange 'b l k d -> (b l) k d' (embedding_manager.py:502-516)."""

    def __init__(self, num_layers=16, num_vectors_per_subj_token=16, out_emb_dim=768, do_zero_shot=True, **kwargs):
        super().__init__()
        if not do_zero_shot:
            raise NotImplementedError("only the zero-shot StaticLayerwiseEmbedding is implemented")
        self.num_layers, self.K, self.out_emb_dim, self.do_zero_shot = num_layers, num_vectors_per_subj_token, out_emb_dim, True
        self.bias, self.has_bias = None, False

    def forward(self, adaface_subj_embs=None):
        b, l, k, d = adaface_subj_embs.shape
        out = adaface_subj_embs.reshape(b * l, k, d)
        self.bias, self.has_bias = out, True
        return out


class EmbeddingManagerLite(nn.Module):
    def __init__(self, tokenizer, subject_strings=("z",), placeholder_tokens: Optional[Dict[str, int]] = None,
                 token2num_vectors: Optional[Dict[str, int]] = None, num_unet_ca_layers=N_CA_LAYERS,
                 arc2face_text_encoder=None, zs_adaface_prompt_embs_inf_type="full_half_pad"):
        super().__init__()
        self.tokenizer = tokenizer
        self.subject_strings = list(subject_strings)
        self.placeholder_strings = list(subject_strings)
        self.background_string_dict = {}
        if placeholder_tokens is None:
            placeholder_tokens = {s: tokenizer.encode(s, add_special_tokens=False)[0] for s in self.subject_strings}
        self.string_to_token_dict = dict(placeholder_tokens)
        self.token2num_vectors = dict(token2num_vectors or {s: 16 for s in self.subject_strings})
        self.use_layerwise_embedding = True
        self.num_unet_ca_layers = num_unet_ca_layers
        self.num_layers_per_embedder = num_unet_ca_layers
        self.do_zero_shot = True
        self.curr_subj_is_face = True
        self.string_to_static_embedder_dict = nn.ModuleDict(
            {s: StaticLayerwiseEmbedding(num_unet_ca_layers, self.token2num_vectors[s]) for s in self.subject_strings})
        self.string_to_subj_basis_generator_dict = nn.ModuleDict()
        self.arc2face_text_encoder = arc2face_text_encoder
        self.zs_adaface_prompt_embs_inf_type = zs_adaface_prompt_embs_inf_type
        self.zs_out_id_embs_scale_range = (1.0, 1.0)
        self.zs_image_feat_dict = {}
        self.iter_type = None
        self.layer_copies_identical = False
        self.clear_prompt_adhoc_info()

    # ------------------------------------------------------------------ ad-hoc state
    def clear_prompt_adhoc_info(self):
        self.placeholder2indices = {}
        self.img_mask = None
        self.prompt_emb_mask = None

    def set_zs_image_features(self, zs_clip_features, zs_id_embs, zs_out_id_embs_scale_range=(1.0, 1.0),
                              add_noise_to_zs_id_embs=False):
        """:1790-1817.  zs_clip_features may be None here (only the background branch reads them)."""
        if self.training and add_noise_to_zs_id_embs:
            raise NotImplementedError("training-time noise on the ID embeddings")
        subj = bg = None
        if zs_clip_features is not None:
            subj, bg = zs_clip_features.chunk(2, dim=1)
        self.zs_image_feat_dict = {"subj": subj, "bg": bg, "id": zs_id_embs}
        self.zs_out_id_embs_scale_range = zs_out_id_embs_scale_range
        for s in self.placeholder_strings:
            self.string_to_static_embedder_dict[s].bias = None

    # ------------------------------------------------------------------ forward
    def forward(self, tokenized_text, embedded_text):
        """tokenized_text int64 [B, N]; embedded_text fp32 [B, N, 768] -> static_embedded_text [16*B, N, 768]."""
        self.clear_prompt_adhoc_info()
        B, N = tokenized_text.shape
        static, tokens_rep, subj_dict = self.get_static_embedding(tokenized_text, embedded_text.clone(), self.zs_image_feat_dict,
                                                                  self.string_to_static_embedder_dict, B, N,
                                                                  self.num_unet_ca_layers, tokenized_text.device)
        self.static_subj_embs_dict = dict(subj_dict)
        self.update_prompt_masks(tokenized_text, tokens_rep)
        return static

    def get_static_embedding(self, tokenized_text, embedded_text, zs_image_feat_dict, embedder_dict, BS, N,
                             num_unet_ca_layers, device):
        if self.training:
            raise NotImplementedError("EmbeddingManagerLite implements the inference path only")
        orig_tokenized_text = tokenized_text
        static_subj_embs_dict = {}
        L = num_unet_ca_layers
        embedded_text = embedded_text.unsqueeze(1).repeat(1, L, 1, 1).view(BS * L, N, -1).contiguous()   # :1349
        tokenized_text = tokenized_text.unsqueeze(1).repeat(1, L, 1).view(BS * L, N).contiguous()       # :1353
        identical = True
        for placeholder_string, placeholder_token in self.string_to_token_dict.items():
            first = ops.find_first_token(tokenized_text, placeholder_token)                             # :1359,:1368
            occurs_rows = int((first >= 0).sum().item())
            if occurs_rows == 0:
                continue
            REAL_OCCURS_IN_BATCH = occurs_rows // self.num_layers_per_embedder                          # :1383
            zs_id_embs = zs_image_feat_dict["id"]
            sbg = self.string_to_subj_basis_generator_dict[placeholder_string]
            if self.arc2face_text_encoder is None:
                raise RuntimeError("EmbeddingManagerLite.arc2face_text_encoder is not set")
            with torch.no_grad():
                _, arc2face_id_embs = arc2face_forward_face_embs(self.tokenizer, self.arc2face_text_encoder, zs_id_embs,
                                                                 return_full_and_core_embs=True)        # :1424
            adaface_subj_embs, _ = sbg(arc2face_id_embs, zs_image_feat_dict.get("subj"), zs_id_embs,
                                       self.zs_out_id_embs_scale_range[0], is_face=self.curr_subj_is_face,
                                       is_training=False,
                                       adaface_prompt_embs_inf_type=self.zs_adaface_prompt_embs_inf_type)  # :1435-1442
            if adaface_subj_embs.shape[0] < REAL_OCCURS_IN_BATCH:                                       # :1449-1451
                adaface_subj_embs = adaface_subj_embs.repeat(REAL_OCCURS_IN_BATCH // adaface_subj_embs.shape[0], 1, 1, 1)
            identical = identical and adaface_subj_embs.shape[1] > 0 and bool(
                (adaface_subj_embs == adaface_subj_embs[:, :1]).all().item())
            subj_static_embedding = embedder_dict[placeholder_string](adaface_subj_embs.float())       # :1508 'b l k d -> (b l) k d'
            static_subj_embs_dict[placeholder_string] = subj_static_embedding
            K = self.token2num_vectors[placeholder_string]
            n_src = subj_static_embedding.shape[0]
            # the k-th vector of source row (occurrence j, layer l) goes to row r = the j-th row-group that holds the
            # placeholder (:1516-1562).  Rows are (b l) ordered, so the i-th row WITH a placeholder reads source row
            # i (or i % 16 when one identity serves the whole batch, :1553).
            has = first >= 0
            rank = (torch.cumsum(has.int(), 0) - 1).to(torch.int32)
            if n_src == L:
                src_index = rank % L
            elif n_src == L * REAL_OCCURS_IN_BATCH:
                src_index = rank
            else:
                raise ValueError(f"{n_src} subject embedding rows for {REAL_OCCURS_IN_BATCH} occurrences")
            src_index = torch.where(has, src_index, torch.zeros_like(src_index)).contiguous()
            ops.splice_rows(embedded_text, subj_static_embedding[:, :K].contiguous(), first, src_index)
            self.update_placeholder_indices(orig_tokenized_text, placeholder_string, placeholder_token, K,
                                            placeholder_is_bg=False)
        self.layer_copies_identical = identical
        return embedded_text, tokenized_text, static_subj_embs_dict

    def update_placeholder_indices(self, tokenized_text, placeholder_string, placeholder_token,
                                   num_vectors_per_subj_token, placeholder_is_bg):
        """:1699-1722 on the un-repeated tokens."""
        first = ops.find_first_token(tokenized_text.contiguous(), placeholder_token).long()
        B_idx = torch.nonzero(first >= 0).flatten()
        if B_idx.numel() == 0:
            self.placeholder2indices[placeholder_string] = None
            return
        N_idx = first[B_idx]
        if num_vectors_per_subj_token > 1:
            K, BS = num_vectors_per_subj_token, B_idx.shape[0]
            B_idx = B_idx.unsqueeze(1).repeat(1, K).view(-1)
            N_idx = N_idx.unsqueeze(1).repeat(1, K).view(-1) + torch.arange(K, device=N_idx.device).repeat(BS)
        self.placeholder2indices[placeholder_string] = (B_idx, N_idx)

    def update_prompt_masks(self, tokenized_text, tokenized_text_repeated=False):
        """:1646-1648."""
        mask = (tokenized_text != 49406) & (tokenized_text != 49407)
        self.prompt_emb_mask = mask.float().unsqueeze(2)
