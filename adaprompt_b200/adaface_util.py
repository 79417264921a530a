"""Host mirror of the Arc2Face helpers in adaface/util.py (reference: askerlee/adaprompt):
arc2face_forward_face_embs :76-125, arc2face_inverse_face_prompt_embs :132-238, get_b_core_e_embeddings :127-129,
gen_gradient_scaler :60-72.  Same signatures; the CLIP passes run on the C ABI through
adaprompt_b200.clip_text.CLIPTextModelWrapper, the 16-row splices are exact row copies (af_splice_rows)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import ops


class _ScaleGradFn(torch.autograd.Function):
    """adaface/util.py:28-47 ScaleGrad: identity forward, gradient times alpha."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = float(alpha)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.alpha, None


class GradientScaler(nn.Module):
    """What gen_gradient_scaler returns for alpha not in {0, 1}: identity forward, d/dx scaled by alpha."""

    def __init__(self, alpha):
        super().__init__()
        self.alpha = float(alpha)

    def forward(self, x):
        return _ScaleGradFn.apply(x, self.alpha) if torch.is_grad_enabled() and x.requires_grad else x


class _Detach(nn.Module):
    def forward(self, x):
        return x.detach()


def gen_gradient_scaler(alpha, debug=False):
    """adaface/util.py:60-72: alpha 1 -> identity, alpha 0 -> detach, otherwise a gradient scaler (forward identity)."""
    if alpha == 1:
        return nn.Identity()
    if alpha > 0:
        return GradientScaler(alpha)
    return _Detach()


def _tokenize(tokenizer, text, max_length, device):
    """Token ids on `device`.  The prompts on this path are constants ("photo of a id person", the 16-comma template), so
    the ids are uploaded once per (tokenizer, text) and kept on the tokenizer object: a pageable host-to-device copy
    synchronises the stream, which would stall the training step every micro-batch and cannot be captured into a CUDA
    graph.  Callers treat the result as read-only."""
    cache = getattr(tokenizer, "_af_device_ids", None)
    if cache is None:
        cache = {}
        try:
            tokenizer._af_device_ids = cache
        except AttributeError:          # a tokenizer without a __dict__: no caching
            pass
    key = (text if isinstance(text, str) else tuple(text), max_length, str(device))
    ids = cache.get(key)
    if ids is None:
        enc = tokenizer(text, truncation=True, padding="max_length", max_length=max_length, return_tensors="pt")
        ids = (enc["input_ids"] if isinstance(enc, dict) else enc.input_ids).to(device)
        if len(cache) >= 64:
            cache.clear()
        cache[key] = ids
    return ids


def arc2face_forward_face_embs(tokenizer, arc2face_text_encoder, face_embs, input_max_length=77,
                               return_full_and_core_embs=True):
    """face_embs: [N, 512] normalised ArcFace embeddings -> (prompt_embeds [N, L, 768], core [N, 16, 768])."""
    arcface_token_id = tokenizer.encode("id", add_special_tokens=False)[0]
    input_ids = _tokenize(tokenizer, "photo of a id person", input_max_length, face_embs.device)
    input_ids = input_ids.repeat(len(face_embs), 1).contiguous()
    face_embs_dtype = face_embs.dtype
    hidden = arc2face_text_encoder.config.hidden_size
    face_embs_padded = F.pad(face_embs.float(), (0, hidden - face_embs.shape[-1]), "constant", 0)   # :103
    token_embs = arc2face_text_encoder(input_ids=input_ids, return_token_embs=True)
    # token_embs[input_ids == arcface_token_id] = face_embs_padded (:107): first occurrence per row, 1 row each
    start = ops.find_first_token(input_ids, arcface_token_id)
    ops.splice_rows(token_embs, face_embs_padded.reshape(len(face_embs), 1, hidden).contiguous(), start)
    prompt_embeds = arc2face_text_encoder(input_ids=input_ids, input_token_embs=token_embs, return_token_embs=False)[0]
    prompt_embeds = prompt_embeds.to(face_embs_dtype)
    if return_full_and_core_embs:
        return prompt_embeds, prompt_embeds[:, 4:20]
    return prompt_embeds[:, 4:20]


def get_b_core_e_embeddings(prompt_embeds, length=22):
    return torch.cat([prompt_embeds[:, :length], prompt_embeds[:, [-1]]], dim=1)


def arc2face_inverse_face_prompt_embs(clip_tokenizer, inverse_text_encoder, face_prompt_embs, list_extra_words,
                                      return_emb_types, pad_embeddings, hidden_state_layer_weights=None,
                                      input_max_length=77, zs_extra_words_scale=0.5):
    """face_prompt_embs: [BS, 16, 768] core ID embeddings -> list of tensors per return_emb_types."""
    if list_extra_words is not None:
        if len(list_extra_words) != len(face_prompt_embs):
            if len(face_prompt_embs) > 1:
                if len(list_extra_words) == 1:
                    list_extra_words = list_extra_words * len(face_prompt_embs)
                else:
                    raise ValueError("list_extra_words has a different length from face_prompt_embs")
            else:
                list_extra_words = list_extra_words[:1]
        for extra_words in list_extra_words:
            assert len(extra_words.split()) <= 2, "Each extra_words string should consist of at most 2 words."
        prompt_templates = ["photo of a " + ", " * 16 + list_extra_words[i] for i in range(len(list_extra_words))]
    else:
        prompt_templates = ["photo of a " + ", " * 16 for _ in range(len(face_prompt_embs))]
    input_ids = _tokenize(clip_tokenizer, prompt_templates, input_max_length, face_prompt_embs.device).contiguous()
    face_prompt_embs_dtype = face_prompt_embs.dtype
    token_embs = inverse_text_encoder(input_ids=input_ids, return_token_embs=True)
    BS = token_embs.shape[0]
    start = torch.full((BS,), 4, dtype=torch.int32, device=token_embs.device)                       # token_embs[:, 4:20] = ... (:184)
    ops.splice_rows(token_embs, face_prompt_embs.float().contiguous(), start)
    prompt_embeds = inverse_text_encoder(input_ids=input_ids, input_token_embs=token_embs,
                                         hidden_state_layer_weights=hidden_state_layer_weights,
                                         return_token_embs=False)[0]
    prompt_embeds = prompt_embeds.to(face_prompt_embs_dtype)
    core_prompt_embs = prompt_embeds[:, 4:20]
    if list_extra_words is not None:
        extra_words_embs = prompt_embeds[:, 20:22] * zs_extra_words_scale
        core_prompt_embs = torch.cat([core_prompt_embs, extra_words_embs], dim=1)
    return_prompts = []
    for emb_type in return_emb_types:
        if emb_type == "full":
            return_prompts.append(prompt_embeds)
        elif emb_type == "full_half_pad":
            p2 = prompt_embeds.clone()
            PADS = p2.shape[1] - 23
            if PADS >= 2:
                p2[:, 22:22 + PADS // 2] = pad_embeddings[22:22 + PADS // 2]
            return_prompts.append(p2)
        elif emb_type == "full_pad":
            p2 = prompt_embeds.clone()
            p2[:, 22:-1] = pad_embeddings[22:-1]
            return_prompts.append(p2)
        elif emb_type == "core":
            return_prompts.append(core_prompt_embs)
        elif emb_type == "full_zeroed_extra":
            p2 = prompt_embeds.clone()
            p2[:, 22:24] = pad_embeddings[22:24]
            p2[:, 24:-1] = 0
            return_prompts.append(p2)
        elif emb_type == "b_core_e":
            return_prompts.append(get_b_core_e_embeddings(prompt_embeds, length=22))
        else:
            raise ValueError(f"unknown emb_type {emb_type!r}")   # the reference calls breakpoint() here
    return return_prompts
