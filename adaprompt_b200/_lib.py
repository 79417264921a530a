"""ctypes binding of libadaface_b200.so (include/adaface_b200.h).

The product path has no CPU fallback: if the library is missing, or there is no sm_100 device,
every op raises.  The library itself loads fine on a CPU-only box (static cudart, driver entry
points resolved lazily) so that symbol-export checks can run without a GPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libadaface_b200.so")

AF_DTYPE_F32 = 0
AF_DTYPE_BF16 = 1


class AfEpilogue(Structure):
    _fields_ = [
        ("bias", c_void_p),
        ("rowbias", c_void_p),
        ("rows_per_group", c_int),
        ("ld_rowbias", c_longlong),
        ("residual", c_void_p),
        ("ldr", c_longlong),
        ("out", c_void_p),
        ("ldo", c_longlong),
        ("out_dtype", c_int),
        ("geglu", c_int),
    ]


# symbol -> (restype, argtypes); mirrors include/adaface_b200.h one to one
SIGNATURES = {
    "af_version": (c_int, []),
    "af_last_error": (c_char_p, []),
    "af_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "af_gemm_bf16": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_int, c_void_p, c_int, c_int,
                             POINTER(AfEpilogue), c_int, c_void_p]),
    "af_conv3x3_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                POINTER(AfEpilogue), c_int, c_void_p]),
    "af_attention_bf16": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_groupnorm_workspace_bytes": (c_size_t, [c_int]),
    "af_groupnorm_silu": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "af_layernorm": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "af_conv_in": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_conv_out": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_timestep_embedding": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "af_linear_small": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_cast_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "af_upsample2x_cast": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "af_cfg_ddim_update": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_longlong, c_void_p]),
    "af_advance_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

_lib = None


class AdaFaceB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the shared library (building is the job of __graft_entry__.build / adaprompt_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdaFaceB200Error(
            f"{LIB_PATH} not found: build it with `python -m adaprompt_b200.build` (needs nvcc, sm_100a). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = load().af_last_error().decode(errors="replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (rc={rc})")
    raise AdaFaceB200Error(f"{what}: {msg} (cudaError {rc})")
