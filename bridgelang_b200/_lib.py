"""_lib.py — ctypes binding of libbridgelang_b200.so (include/bridgelang_b200.h).

There is deliberately no fallback: if the library is missing or a call fails the caller gets an exception.
Status translation follows SURVEY.md §8b (non-zero status → RuntimeError).
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

PKG_DIR = Path(__file__).resolve().parent
# BLB_LIB: development override (A/B of two builds on one box); the product path is the in-tree library
LIB_PATH = Path(os.environ["BLB_LIB"]) if os.environ.get("BLB_LIB") else PKG_DIR / "libbridgelang_b200.so"

EPI_BIAS, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_PATCH, EPI_BIAS_QGELU = 0, 1, 2, 3, 4
DTYPE_F32, DTYPE_BF16, DTYPE_F16, DTYPE_F64 = 0, 1, 2, 3

c_f32p = C.POINTER(C.c_float)


class Epilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("gamma", C.c_void_p), ("resid", C.c_void_p), ("ld_resid", C.c_int32),
        ("out", C.c_void_p), ("ld_out", C.c_int32), ("out_col_off", C.c_int32), ("pos", C.c_void_p),
        ("tok_in", C.c_int32), ("tok_out", C.c_int32), ("tok_shift", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p), ("ln_parts", C.c_int32), ("ln_eps", C.c_float),
        ("stats_out", C.c_void_p), ("xb_out", C.c_void_p), ("ld_xb", C.c_int32),
        ("shift_in", C.c_void_p), ("shift_out", C.c_void_p),
    ]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1",
        "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2", "qkv_colsum", "fc1_colsum")]


class VitWeights(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32), ("hidden_pad", C.c_int32),
        ("n_prefix", C.c_int32), ("n_blocks", C.c_int32), ("patch_ldk", C.c_int32), ("ln_eps", C.c_float),
        ("patch_w", C.c_void_p), ("patch_b", C.c_void_p), ("pos_embed", C.c_void_p), ("prefix", C.c_void_p),
        ("blocks_host", C.POINTER(BlockWeights)), ("ln_folded", C.c_int32),
        ("hidden", C.c_int32), ("grid", C.c_int32), ("img_size", C.c_int32), ("act", C.c_int32),
        ("norm_pre_w", C.c_void_p), ("norm_pre_b", C.c_void_p), ("patch_w_u8", C.c_void_p), ("patch_b_u8", C.c_void_p),
    ]


class ProjectorWeights(C.Structure):
    _fields_ = [
        ("in_dim", C.c_int32), ("hidden_dim", C.c_int32), ("out_dim", C.c_int32),
        ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p),
        ("fc3_w", C.c_void_p), ("fc3_b", C.c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/bridgelang_b200.h one to one
_SIGNATURES = {
    "blb_abi_version": (C.c_int, []),
    "blb_status_string": (C.c_char_p, [C.c_int]),
    "blb_launch_count": (C.c_longlong, []),
    "blb_set_gemm_cta_group": (None, [C.c_int]),
    "blb_debug_attention_trace": (None, [C.c_void_p]),
    "blb_timing_enable": (None, [C.c_int]),
    "blb_timing_reset": (None, []),
    "blb_timing_collect": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "blb_timing_records": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]),
    "blb_gemm_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(Epilogue), C.c_void_p]),
    "blb_gemm_stats_parts": (C.c_int, [C.c_int]),
    "blb_rowstats_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "blb_layernorm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_float, C.c_void_p]),
    "blb_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "blb_im2col_patch14": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "blb_u8_to_patches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "blb_resize_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "blb_vit_workspace_bytes": (C.c_size_t, [C.POINTER(VitWeights), C.c_int]),
    "blb_vit_tower_forward_u8": (C.c_int, [C.POINTER(VitWeights), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    "blb_patch_matrix_bytes": (C.c_size_t, [C.POINTER(VitWeights), C.c_int]),
    "blb_fused_featurize_project_forward_u8": (C.c_int, [C.POINTER(VitWeights), C.POINTER(VitWeights),
                                                         C.POINTER(ProjectorWeights), C.c_void_p, C.c_int, C.c_void_p,
                                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "blb_vit_tower_forward": (C.c_int, [C.POINTER(VitWeights), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "blb_projector_workspace_bytes": (C.c_size_t, [C.POINTER(ProjectorWeights), C.c_int]),
    "blb_projector_forward": (C.c_int, [C.POINTER(ProjectorWeights), C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "blb_fused_workspace_bytes": (C.c_size_t, [C.POINTER(VitWeights), C.POINTER(VitWeights),
                                               C.POINTER(ProjectorWeights), C.c_int]),
    "blb_fused_featurize_project_forward": (C.c_int, [C.POINTER(VitWeights), C.POINTER(VitWeights),
                                                      C.POINTER(ProjectorWeights), C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "blb_preprocess_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "blb_encode_actions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "blb_action_token_metrics": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "blb_argmax": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "blb_argmax_window": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "blb_argmax_window_detokenize_unnormalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int,
                                                           C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                           C.c_void_p]),
    "blb_detokenize_unnormalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "blb_argmax_detokenize_unnormalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int,
                                                    C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library (no torch types cross this boundary)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m bridgelang_b200.build` "
                "(there is no CPU or eager fallback for this path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().blb_status_string(status).decode()
        raise RuntimeError(f"bridgelang_b200 {what} failed with status {status}: {msg}")
