"""ctypes binding of ``include/mcm_b200.h`` (the C ABI of the sm_100a extension).

The product path has no CPU fallback: if the shared library is missing, or a compute entry point
is called without a B200, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 2

OK, EINVAL, ESTATE, ECUDA, ENOMEM, EUNSUPPORTED = range(6)

PROF_KINDS = ["patchify", "gemm_patch", "embed_finish", "gemm_qkv", "attention", "gemm_out", "layernorm",
              "gemm_fc1", "gemm_fc2", "tail", "gemm_other"]

OPT_CLS_SHORTCUT = 1
OPT_PRECISION = 2
OPT_CUDA_GRAPH = 3
PRECISION_FP16 = 0
PRECISION_SPLIT = 1
PRECISIONS = {"fp16": PRECISION_FP16, "split": PRECISION_SPLIT, 0: PRECISION_FP16, 1: PRECISION_SPLIT}

SCORE_KINDS = {"MCM": 0, "max-logit": 1, "energy": 2, "entropy": 3, "var": 4}


class McmConfig(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("patch", C.c_int32), ("width", C.c_int32), ("layers", C.c_int32),
        ("heads", C.c_int32), ("mlp", C.c_int32), ("proj", C.c_int32), ("eps", C.c_float),
        ("max_batch", C.c_int32), ("device", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/mcm_b200.h
_H = C.c_void_p
_P = C.c_void_p
SIGNATURES = {
    "mcm_abi_version": (C.c_int32, []),
    "mcm_create": (C.c_int, [C.POINTER(McmConfig), C.POINTER(_H)]),
    "mcm_destroy": (None, [_H]),
    "mcm_last_error": (C.c_char_p, [_H]),
    "mcm_load_weight": (C.c_int, [_H, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int32)]),
    "mcm_finalize_weights": (C.c_int, [_H]),
    "mcm_set_text_bank": (C.c_int, [_H, _P, C.c_int32, C.c_int32]),
    "mcm_image_features": (C.c_int, [_H, _P, C.c_int32, _P, _P]),
    "mcm_score": (C.c_int, [_H, _P, C.c_int32, C.c_float, C.c_int32, _P, _P]),
    "mcm_score_stream_host": (C.c_int, [_H, _P, C.c_int64, C.c_int32, C.c_float, C.c_int32, _P]),
    "mcm_set_normalization": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mcm_image_features_u8": (C.c_int, [_H, _P, C.c_int32, _P, _P]),
    "mcm_score_u8": (C.c_int, [_H, _P, C.c_int32, C.c_float, C.c_int32, _P, _P]),
    "mcm_score_stream_host_u8": (C.c_int, [_H, _P, C.c_int64, C.c_int32, C.c_float, C.c_int32, _P]),
    "mcm_resize_crop_u8": (C.c_int, [_H, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, _P, _P]),
    "mcm_dbg_resize_tables": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _P, _P, C.c_int32]),
    "mcm_set_maha": (C.c_int, [_H, _P, _P, C.c_int32, C.c_int32]),
    "mcm_maha_score": (C.c_int, [_H, _P, C.c_int32, _P, _P]),
    "mcm_dbg_maha_from_features": (C.c_int, [_H, _P, C.c_int32, _P, _P]),
    "mcm_score_stream_host_images": (C.c_int, [_H, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64,
                                              C.c_int32, C.c_float, C.c_int32, _P]),
    "mcm_launch_count": (C.c_int64, [_H]),
    "mcm_reset_launch_count": (None, [_H]),
    "mcm_set_option": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "mcm_allgather_scores": (C.c_int, [_H, _P, _P, C.c_int32, _P, _P]),
    "mcm_dbg_gemm_split": (C.c_int, [_H, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mcm_dbg_attention_split": (C.c_int, [_H, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mcm_profile_enable": (C.c_int, [_H, C.c_int32]),
    "mcm_profile_read": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "mcm_flops_per_image": (C.c_double, [C.POINTER(McmConfig), C.c_int32]),
    "mcm_dbg_gemm": (C.c_int, [_H, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mcm_dbg_fold_ln": (C.c_int, [_H, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mcm_dbg_gemm_ln": (C.c_int, [_H, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, _P]),
    "mcm_dbg_gemm_resid_ln": (C.c_int, [_H, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_int32), _P]),
    "mcm_dbg_gemm_resid_h2": (C.c_int, [_H, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _P]),
    "mcm_dbg_layernorm": (C.c_int, [_H, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P]),
    "mcm_dbg_attention": (C.c_int, [_H, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mcm_dbg_tail": (C.c_int, [_H, _P, C.c_int32, C.c_float, C.c_int32, _P, _P, _P]),
    "mcm_dbg_embed": (C.c_int, [_H, _P, C.c_int32, _P, _P]),
}

_lib = None
_lib_override = None


def use_library(path: str) -> None:
    """Development tools only (``tools/``): load an A/B variant built by ``build.build_variant`` instead of the
    in-tree library.  Must be called before the first engine is created; nothing in the product reads the environment."""
    global _lib_override
    if _lib is not None:
        raise RuntimeError("the mcm_b200 library is already loaded")
    _lib_override = str(path)


def lib_path() -> str:
    return _lib_override or _build.LIB_PATH


def load() -> C.CDLL:
    """Load (once) the in-tree shared library and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.isfile(path):
        raise RuntimeError(
            f"mcm_b200 CUDA extension not built ({path} is missing); run `python -m mcm_b200.build` "
            "(needs nvcc).  There is no CPU fallback for this path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mcm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.mcm_abi_version()} != {ABI_VERSION}; rebuild the extension")
    _lib = lib
    return lib


def error_text(handle) -> str:
    msg = load().mcm_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle=None) -> None:
    """Translate a non-zero return code into the exception the reference's callers would see
    (ValueError for bad arguments like HF's wrong-image-size check, RuntimeError otherwise)."""
    if rc == OK:
        return
    msg = error_text(handle) or f"mcm_b200 error {rc}"
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
