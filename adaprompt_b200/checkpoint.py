"""Loading the reference's AdaFace checkpoints ("embeddings_gs-*.pt").

The reference saves PICKLED nn.Module OBJECTS, not state_dicts (EmbeddingManager.save,
ldm/modules/embedding_manager.py:1824-1838): `string_to_subj_basis_generator_dict` is an nn.ModuleDict of
adaface.subj_basis_generator.SubjBasisGenerator instances whose `prompt2token_proj` is an
adaface.arc2face_models.CLIPTextModelWrapper holding transformers' CLIP modules (and CLIPAttentionMKV layers).  Consumers
index the dict and use the objects directly (adaface/adaface_wrapper.py:49-59, embedding_manager.py:1884-1911).  Older
checkpoints name the same classes ldm.modules.subj_basis_generator / ldm.modules.arc2face_models (:7-8).

Neither the reference package nor a matching transformers version can be assumed at load time, so nothing here imports
them: every class under adaface.* / ldm.* / transformers.* / diffusers.* found in the pickle is materialised as a
`PickledModule` shell (an nn.Module that just receives the pickled __dict__: parameters, buffers, sub-modules, scalar
attributes), and `to_native_subj_basis_generator` rebuilds the native adaprompt_b200 SubjBasisGenerator from it: the
HF-keyed state_dict loads unchanged (clip_text.py keeps the key names), multi-key/value layers are re-created with the
multiplier read off the k_proj shape, and the scalar attributes the reference reads after unpickling are copied.

  load_adaface_ckpt(path)        -> the checkpoint dict with native SubjBasisGenerators
  install_import_aliases()       -> `import adaface.subj_basis_generator`, `ldm.modules.arc2face_models`, ... resolve to
                                    the mirrors (drop-in for reference-side code and plain torch.load)
"""
from __future__ import annotations

import pickle
import sys
import types
from typing import Dict

import torch
from torch import nn

_SHELL_PREFIXES = ("adaface.", "ldm.", "transformers.", "diffusers.")
_SCALAR_ATTRS = ("placeholder_is_bg", "num_out_layers", "num_out_embs_per_layer", "num_out_embs", "output_dim",
                 "zs_extra_words_scale", "output_scale", "prompt2token_proj_grad_scale",
                 "prompt2token_proj_attention_multiplier", "num_id_vecs")


class PickledModule(nn.Module):
    """Shell for a pickled object whose class is not importable here; `pickled_class` names the original."""
    pickled_class = "?"

    def __init__(self, *a, **k):            # some classes pickle by constructor call (__reduce__): swallow the arguments
        super().__init__()
        self.pickled_args = (a, k)

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:                               # (dict, slots) form
            for part in state or ():
                if isinstance(part, dict):
                    self.__dict__.update(part)
        for name in ("_parameters", "_buffers", "_modules"):
            self.__dict__.setdefault(name, {})
        nn.Module.__setstate__(self, self.__dict__)

    def forward(self, *a, **k):
        raise RuntimeError(f"{self.pickled_class} was unpickled as a shell; convert it with adaprompt_b200.checkpoint")

    def extra_repr(self):
        return f"pickled_class={self.pickled_class}"


_shell_classes: Dict[str, type] = {}


def _shell(module: str, name: str) -> type:
    key = f"{module}.{name}"
    cls = _shell_classes.get(key)
    if cls is None:
        cls = _shell_classes[key] = type(name, (PickledModule,), {"pickled_class": key, "__module__": __name__})
    return cls


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith(_SHELL_PREFIXES):
            return _shell(module, name)
        return super().find_class(module, name)


_pickle_module = types.ModuleType("adaprompt_b200._ckpt_pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()
_pickle_module.__dict__.update({k: getattr(pickle, k) for k in ("dump", "dumps", "loads", "Pickler", "HIGHEST_PROTOCOL",
                                                                 "UnpicklingError", "PicklingError")})


def is_pickled_sbg(obj) -> bool:
    return isinstance(obj, PickledModule) and obj.pickled_class.endswith("subj_basis_generator.SubjBasisGenerator")


def to_native_subj_basis_generator(shell, clip_tokenizer=None):
    """PickledModule shell of a reference SubjBasisGenerator -> adaprompt_b200.subj_basis_generator.SubjBasisGenerator."""
    from .clip_text import CLIPAttentionMKV, CLIPTextConfigLite
    from .subj_basis_generator import SubjBasisGenerator
    d = shell.__dict__
    if d.get("placeholder_is_bg", False):
        raise NotImplementedError("background-token SubjBasisGenerator checkpoints (CLIP-vision branch) are out of scope")
    p2t = d["_modules"].get("prompt2token_proj") or d.get("prompt2token_proj")
    if p2t is None:
        raise ValueError("pickled SubjBasisGenerator has no prompt2token_proj")
    sd = {k: v for k, v in p2t.state_dict().items()}
    emb_w = sd["text_model.embeddings.token_embedding.weight"]
    pos_w = sd["text_model.embeddings.position_embedding.weight"]
    layers = sorted({int(k.split(".")[3]) for k in sd if k.startswith("text_model.encoder.layers.")})
    hidden = emb_w.shape[1]
    cfg = CLIPTextConfigLite(hidden_size=hidden, intermediate_size=sd["text_model.encoder.layers.0.mlp.fc1.weight"].shape[0],
                             num_attention_heads=hidden // 64, num_hidden_layers=len(layers),
                             vocab_size=emb_w.shape[0], max_position_embeddings=pos_w.shape[0])
    hw = d["_parameters"].get("hidden_state_layer_weights")
    num_id_vecs = d.get("num_id_vecs", 77)
    native = SubjBasisGenerator(num_id_vecs={"subj": num_id_vecs if isinstance(num_id_vecs, int) else 77, "bg": 257},
                                num_out_embs_per_layer=d.get("num_out_embs_per_layer", 16),
                                num_out_layers=d.get("num_out_layers", 16), output_dim=d.get("output_dim", hidden),
                                prompt2token_proj_grad_scale=d.get("prompt2token_proj_grad_scale", 0.4),
                                zs_extra_words_scale=d.get("zs_extra_words_scale", 0.5),
                                learnable_hidden_state_weights_scheme="per-layer" if hw is not None else "none",
                                clip_tokenizer=clip_tokenizer if clip_tokenizer is not None else d.get("clip_tokenizer"),
                                clip_config=cfg)
    # multi-key/value attention layers (arc2face_models.py:285-302): the multiplier is the k_proj fan-out
    tm = native.prompt2token_proj.text_model
    for i in layers:
        m = sd[f"text_model.encoder.layers.{i}.self_attn.k_proj.weight"].shape[0] // hidden
        if m > 1:
            old = tm.encoder.layers[i].self_attn
            new = CLIPAttentionMKV(old.config, m)
            new.q_proj, new.out_proj = old.q_proj, old.out_proj
            tm.encoder.layers[i].self_attn = new
    sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
    missing, unexpected = native.prompt2token_proj.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not k.endswith("position_ids")]
    if missing or unexpected:
        raise ValueError(f"prompt2token_proj keys do not line up: missing {missing[:4]}, unexpected {unexpected[:4]}")
    with torch.no_grad():
        if hw is not None:
            native.hidden_state_layer_weights.data = hw.detach().clone().float()
        for name in ("pos_embs",):
            t = d["_parameters"].get(name)
            if t is not None and t.shape == getattr(native, name).shape:
                getattr(native, name).data.copy_(t)
        ln = d["_modules"].get("pos_embs_ln")
        if ln is not None:
            native.pos_embs_ln.load_state_dict(ln.state_dict())
    for name in _SCALAR_ATTRS:
        if name in d and name != "num_id_vecs":
            setattr(native, name, d[name])
    pad = d.get("pad_embeddings")
    native.pad_embeddings = pad.detach().clone() if torch.is_tensor(pad) else None
    mult = d.get("prompt2token_proj_attention_multiplier", -1)
    native.prompt2token_proj_attention_multiplier = mult
    return native


def load_adaface_ckpt(path, clip_tokenizer=None, map_location="cpu") -> dict:
    """torch.load of a reference embedding-manager checkpoint without the reference package: every pickled
    SubjBasisGenerator (in `string_to_subj_basis_generator_dict`, a ModuleDict shell or a plain dict) is converted to the
    native class; everything else is returned as unpickled (tensors, dicts, shells)."""
    ckpt = torch.load(path, map_location=map_location, pickle_module=_pickle_module, weights_only=False)
    if not isinstance(ckpt, dict):
        raise ValueError(f"{path}: expected the dict EmbeddingManager.save writes")
    sbgs = ckpt.get("string_to_subj_basis_generator_dict")
    if sbgs is not None:
        items = sbgs._modules.items() if isinstance(sbgs, nn.Module) else sbgs.items()
        ckpt["string_to_subj_basis_generator_dict"] = {
            k: (to_native_subj_basis_generator(v, clip_tokenizer) if is_pickled_sbg(v) else v) for k, v in items}
    return ckpt


# ------------------------------------------------------------------------------------------------ import-path aliases
_ALIASES = {
    "adaface": None, "adaface.subj_basis_generator": "adaprompt_b200.subj_basis_generator",
    "adaface.arc2face_models": "adaprompt_b200.clip_text", "adaface.util": "adaprompt_b200.adaface_util",
    "adaface.adaface_wrapper": "adaprompt_b200.adaface_wrapper",
    "ldm": None, "ldm.modules": None, "ldm.modules.diffusionmodules": None, "ldm.models": None,
    "ldm.models.diffusion": None,
    "ldm.modules.subj_basis_generator": "adaprompt_b200.subj_basis_generator",      # legacy checkpoint names
    "ldm.modules.arc2face_models": "adaprompt_b200.clip_text",
    "ldm.modules.attention": "adaprompt_b200.attention",
    "ldm.modules.diffusionmodules.openaimodel": "adaprompt_b200.unet",
    "ldm.modules.diffusionmodules.util": "adaprompt_b200.diffusion_util",
    "ldm.modules.embedding_manager": "adaprompt_b200.embedding_manager",
    "ldm.models.diffusion.ddim": "adaprompt_b200.ddim",
    "ldm.prodigy": "adaprompt_b200.prodigy",
}


def install_import_aliases(force: bool = False):
    """Registers the reference's module paths in sys.modules as aliases of the mirrors, so reference-side code
    (`from ldm.models.diffusion.ddim import DDIMSampler`, `instantiate_from_config` on
    "ldm.modules.diffusionmodules.openaimodel.UNetModel", v1-inference-ada.yaml:36) picks up the B200 modules without
    edits.  Existing entries (a real reference checkout on sys.path) are left alone unless force=True.
    Returns the list of names installed."""
    import importlib
    done = []
    for name, target in _ALIASES.items():
        if name in sys.modules and not force:
            continue
        if target is None:
            mod = types.ModuleType(name)
            mod.__path__ = []          # a package
        else:
            mod = importlib.import_module(target)
        sys.modules[name] = mod
        done.append(name)
    for name in done:                   # attach children to parents so attribute access works too
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, sys.modules[name])
    return done
