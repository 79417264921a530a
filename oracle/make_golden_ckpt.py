"""TEST INFRASTRUCTURE - writes tests/golden/adaface_ckpt_tiny.pt: a checkpoint in the reference's own on-disk format,
produced by the UNMODIFIED reference classes (/root/reference, build container only):

  * adaface.subj_basis_generator.SubjBasisGenerator (instance built with object.__new__ because __init__ downloads
    pretrained CLIP, as in make_golden_text.py) whose prompt2token_proj is a reference
    adaface.arc2face_models.CLIPTextModelWrapper (transformers' CLIP text modules), with layer 1 extended to 2 keys /
    values per token by the reference's extend_clip_attention_MKV_multiplier;
  * saved with the dict layout of EmbeddingManager.save (ldm/modules/embedding_manager.py:1824-1838):
    string_to_subj_basis_generator_dict is an nn.ModuleDict.
The CLIP tower is tiny (hidden 128, 2 heads, 2 layers, vocab 96) so the pickle stays ~1.5 MB; loading exercises exactly
the class paths and object structure of a real embeddings_gs-*.pt.  The plain tensors are stored next to it for the test.

    python oracle/make_golden_ckpt.py
"""
from __future__ import annotations

import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.make_golden_text import build_ref_sbg, import_text_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ns = import_text_reference()
    from transformers import CLIPTextConfig
    torch.manual_seed(77)
    cfg = CLIPTextConfig(hidden_size=128, intermediate_size=256, num_attention_heads=2, num_hidden_layers=2,
                         vocab_size=96, max_position_embeddings=77, hidden_act="quick_gelu")
    enc = ns.am.CLIPTextModelWrapper(cfg)
    enc.extend_clip_attention_MKV_multiplier(begin_layer_idx=1, end_layer_idx=2, multiplier=2, noise_std=0.1)
    sbg = build_ref_sbg(ns, enc, tok=None)
    sbg.output_dim = 128
    sbg.prompt2token_proj_attention_multiplier = 2
    sbg.pos_embs = nn.Parameter(torch.randn(1, 77, 128))
    sbg.pos_embs_ln = nn.LayerNorm(128)
    with torch.no_grad():
        sbg.hidden_state_layer_weights.copy_(torch.tensor([[0.7], [2.2], [3.9]]))
    sbg.pad_embeddings = torch.randn(77, 128)
    ckpt = {"string_to_token": {"z": torch.tensor([345])}, "string_to_static_embedder": nn.ParameterDict(),
            "string_to_subj_basis_generator_dict": nn.ModuleDict({"z": sbg}), "token2num_vectors": {"z": 16},
            "emb_global_scale_scores": None, "use_conv_attn_kernel_size": -1, "placeholder_strings": ["z"],
            "subject_strings": ["z"], "background_strings": [], "ca_q_bns": None, "ca_outfeat_lns": None,
            "do_zero_shot": True}
    path = os.path.join(OUT, "adaface_ckpt_tiny.pt")
    torch.save(ckpt, path)
    expected = {"prompt2token_proj": {k: v.clone() for k, v in enc.state_dict().items()},
                "hidden_state_layer_weights": sbg.hidden_state_layer_weights.detach().clone(),
                "pos_embs": sbg.pos_embs.detach().clone(), "pad_embeddings": sbg.pad_embeddings.clone(),
                "class_of_sbg": type(sbg).__module__ + "." + type(sbg).__name__,
                "class_of_attn_layer1": type(enc.text_model.encoder.layers[1].self_attn).__module__ + "." +
                                        type(enc.text_model.encoder.layers[1].self_attn).__name__}
    torch.save(expected, os.path.join(OUT, "adaface_ckpt_tiny_expected.pt"))
    print(path, os.path.getsize(path), expected["class_of_sbg"], expected["class_of_attn_layer1"])


if __name__ == "__main__":
    main()
