"""Host mirror of adaface/subj_basis_generator.py:SubjBasisGenerator (reference: askerlee/adaprompt), face path.

Constructor arguments, attribute names (read by the reference after unpickling checkpoints: prompt2token_proj,
prompt2token_proj_attention_multiplier, hidden_state_layer_weights, pad_embeddings, num_out_layers, ...) and the
forward signature :470-471 are kept.  The background-token branch (placeholder_is_bg: latent queries cross-attending
CLIP-vision features, :540-556) and the DINO object branch (:529-533) are out of scope and raise.
`prompt2token_proj` is adaprompt_b200.clip_text.CLIPTextModelWrapper, i.e. the CLIP pass runs on the C ABI.
"""
from __future__ import annotations

import torch
from torch import nn

from .adaface_util import arc2face_inverse_face_prompt_embs, gen_gradient_scaler
from .clip_text import CLIPTextConfigLite, CLIPTextModelWrapper


class SubjBasisGenerator(nn.Module):
    def __init__(self, num_heads=6, num_id_vecs={"subj": 77, "bg": 257}, num_out_embs_per_layer=4, num_out_layers=16,
                 image_embedding_dim=768, dino_embedding_dim=384, output_dim=768, placeholder_is_bg: bool = False,
                 prompt2token_proj_grad_scale: float = 0.4, zs_extra_words_scale: float = 0.5,
                 learnable_hidden_state_weights_scheme: str = "per-layer",
                 bg_prompt_translator_has_to_out_proj: bool = False, clip_tokenizer=None, clip_config=None):
        super().__init__()
        if placeholder_is_bg:
            raise NotImplementedError("background-token SubjBasisGenerator (needs CLIP-vision features) is out of scope")
        self.placeholder_is_bg = placeholder_is_bg
        self.num_out_layers = num_out_layers
        self.num_out_embs_per_layer = num_out_embs_per_layer
        self.num_out_embs = num_out_layers * num_out_embs_per_layer
        self.output_dim = output_dim
        self.num_id_vecs = num_id_vecs["subj"]
        self.pos_embs = nn.Parameter(torch.randn(1, self.num_id_vecs, output_dim))
        self.pos_embs_ln = nn.LayerNorm(output_dim)
        self.zs_extra_words_scale = zs_extra_words_scale
        self.output_scale = output_dim ** -0.5
        self.clip_tokenizer = clip_tokenizer          # reference: CLIPTokenizer.from_pretrained (needs network)
        self.obj_proj_in = None                       # DINO object branch: out of scope
        self.prompt2token_proj = CLIPTextModelWrapper(clip_config or CLIPTextConfigLite())
        self.prompt2token_proj_grad_scale = prompt2token_proj_grad_scale
        self.prompt2token_proj_grad_scaler = gen_gradient_scaler(prompt2token_proj_grad_scale)
        if prompt2token_proj_grad_scale == 0:
            self.freeze_prompt2token_proj()
        self.prompt2token_proj_attention_multiplier = -1
        self.initialize_hidden_state_layer_weights(learnable_hidden_state_weights_scheme, "cpu")
        self.pad_embeddings = None
        self.bg_proj_in = None

    def forward(self, arc2face_id_embs, clip_features=None, raw_id_embs=None, out_id_embs_scale=1.0, is_face=True,
                is_training=False, adaface_prompt_embs_inf_type="full_half_pad"):
        """arc2face_id_embs [BS,16,768] -> (adaface_subj_embs [BS, num_out_layers, 16, 768], adaface_prompt_embs [BS,77,768])."""
        if not is_face:
            raise NotImplementedError("object (DINO) branch is out of scope")
        assert arc2face_id_embs is not None
        if self.clip_tokenizer is None:
            raise RuntimeError("SubjBasisGenerator.clip_tokenizer is not set (CLIP vocabulary files are unavailable offline)")
        hidden_state_layer_weights = self.hidden_state_layer_weights_grad_scaler(self.hidden_state_layer_weights)   # :498
        return_emb_types = ["full_pad", "core"] if is_training else [adaface_prompt_embs_inf_type, "core"]        # :501-505
        if self.pad_embeddings is None:
            self.generate_pad_embeddings()
        else:
            self.pad_embeddings = self.pad_embeddings.to(arc2face_id_embs.device)
        with torch.no_grad():
            adaface_prompt_embs, core_id_embs = arc2face_inverse_face_prompt_embs(
                self.clip_tokenizer, self.prompt2token_proj, arc2face_id_embs, list_extra_words=None,
                return_emb_types=return_emb_types, pad_embeddings=self.pad_embeddings,
                hidden_state_layer_weights=hidden_state_layer_weights, input_max_length=77,
                zs_extra_words_scale=self.zs_extra_words_scale)                                                    # :519-527
        adaface_prompt_embs = self.prompt2token_proj_grad_scaler(adaface_prompt_embs)
        core_id_embs = self.prompt2token_proj_grad_scaler(core_id_embs)
        adaface_subj_embs = core_id_embs.unsqueeze(1).repeat(1, self.num_out_layers, 1, 1)                         # :558
        if out_id_embs_scale != 1:                                                                                 # :561-565
            pad_embeddings = self.pad_embeddings[4:4 + self.num_out_embs_per_layer].unsqueeze(0).unsqueeze(0)
            adaface_subj_embs = adaface_subj_embs * out_id_embs_scale + pad_embeddings * (1 - out_id_embs_scale)
        return adaface_subj_embs, adaface_prompt_embs

    def initialize_hidden_state_layer_weights(self, learnable_hidden_state_weights_scheme, device):
        """:569-585."""
        if learnable_hidden_state_weights_scheme == "none":
            self.hidden_state_layer_weights = None
            self.hidden_state_layer_weights_grad_scaler = gen_gradient_scaler(1)
        elif learnable_hidden_state_weights_scheme == "per-layer":
            self.hidden_state_layer_weights = nn.Parameter(torch.tensor([[1.0], [2.0], [4.0]], device=device),
                                                           requires_grad=True)
            self.hidden_state_layer_weights_grad_scaler = gen_gradient_scaler(5)
        else:
            raise ValueError(learnable_hidden_state_weights_scheme)

    def generate_pad_embeddings(self):
        """:587-602: token + position embedding of 77 pad tokens (pad id 49407), detached."""
        emb = self.prompt2token_proj.text_model.embeddings
        pad_id = getattr(self.clip_tokenizer, "pad_token_id", 49407)
        pad_tokens = torch.full((1, 77), pad_id, dtype=torch.long, device=emb.token_embedding.weight.device)
        self.pad_embeddings = emb(pad_tokens)[0].detach()

    def extend_prompt2token_proj_attention(self, begin_layer_idx=-1, end_layer_idx=-1, multiplier=2, noise_std=0.1):
        if multiplier > 1:
            n = self.prompt2token_proj.extend_clip_attention_MKV_multiplier(begin_layer_idx, end_layer_idx, multiplier,
                                                                            noise_std)
            self.prompt2token_proj_attention_multiplier = multiplier
            return n
        return 0

    def freeze_prompt2token_proj(self):
        if self.prompt2token_proj is not None:
            for p in self.prompt2token_proj.parameters():
                p.requires_grad = False
