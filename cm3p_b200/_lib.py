"""ctypes loader for libcm3p_b200.so (the C ABI declared in include/cm3p_b200.h).

There is no fallback: if the shared library is missing `load()` raises, and every compute entry
point itself returns CM3P_ERR_ARCH on a non-sm_100 device, which `check()` turns into a
RuntimeError carrying `cm3p_last_error()`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# CM3P_LIB_PATH: load a differently-built copy of the same library (kernel ablation builds for profiling)
LIB_PATH = os.environ.get("CM3P_LIB_PATH") or os.path.join(HERE, "libcm3p_b200.so")

_P = c_void_p
_I = c_int
_L = c_int64
_F = c_float

# name -> (restype, argtypes); must list every symbol declared in include/cm3p_b200.h
SIGNATURES = {
    "cm3p_last_error": (c_char_p, []),
    "cm3p_version": (_I, []),
    "cm3p_num_sms": (_I, []),
    "cm3p_set_option": (_I, [_I, _I]),
    "cm3p_get_option": (_I, [_I]),
    "cm3p_attn_pack_groups": (_I, [_P, _I, _P, _P, _I, _P]),
    "cm3p_gemm_bf16": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _L, _L, _L, _I, _P, _L, _P, _L, _F, _I, _P, _P, _L, _P, _L, _L,
                             _P]),
    "cm3p_gemm_bf16_ln": (_I, [_P, _L, _P, _L, _P, _L, _L, _L, _L, _I, _P, _L, _P, _L, _P, _P, _L, _P, _P, _P, _F, _P]),
    "cm3p_attn_varlen_fwd": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "cm3p_layernorm_fwd": (_I, [_P, _P, _P, _P, _L, _I, _F, _P]),
    "cm3p_embed_gather_ln": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _P]),
    "cm3p_conv1d_k3_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cm3p_conv1d_k3_wgrad": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "cm3p_transpose_cast_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "cm3p_pool_project_normalize": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "cm3p_clip_loss_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "cm3p_attn_varlen_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "cm3p_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _F, _P]),
    "cm3p_embed_gather_ln_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _P]),
    "cm3p_geglu_bwd": (_I, [_P, _P, _P, _P, _L, _I, _P]),
    "cm3p_gelu_fwd": (_I, [_P, _P, _L, _P]),
    "cm3p_gelu_bwd": (_I, [_P, _P, _P, _L, _P]),
    "cm3p_colsum_f32": (_I, [_P, _P, _L, _I, _P]),
    "cm3p_pool_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "cm3p_l2norm_bwd": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "cm3p_clip_loss_bwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _P, _I, _I, _I, _P]),
    "cm3p_conv2_col2im_gelu_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "cm3p_vocab_ce_fwd": (_I, [_P, _L, _P, _P, _I, _P, _P, _P, _L, _I, _P]),
    "cm3p_vocab_ce_bwd": (_I, [_P, _L, _P, _P, _I, _P, _P, _L, _I, _P]),
    "cm3p_gather_rows": (_I, [_P, _P, _P, _L, _I, _P]),
    "cm3p_scatter_add_rows": (_I, [_P, _P, _P, _L, _I, _P]),
    "cm3p_segment_accumulate": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "cm3p_mean_renormalize": (_I, [_P, _P, _P, _I, _I, _P]),
    "cm3p_logmel_frames": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _P]),
    "cm3p_logmel_power_mel": (_I, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cm3p_logmel_finalize": (_I, [_P, _P, _I, _L, _P]),
    "cm3p_normalize_vectors": (_I, [_P, _P, _L, _I, _P]),
    "cm3p_knn_workspace_bytes": (_L, [_L, _I]),
    "cm3p_knn_cosine": (_I, [_P, _L, _I, _L, _I, _P, _P, _P, _L, _P]),
    "cm3p_pca2_workspace_floats": (_L, [_L, _I]),
    "cm3p_pca2": (_I, [_P, _L, _I, _P, _I, _P, _P, _P, _P, _L, _P]),
    "cm3p_kmeans_workspace_bytes": (_L, [_L, _I, _I]),
    "cm3p_kmeans": (_I, [_P, _L, _I, _I, _L, _I, _P, _P, _P, _P, _L, _P]),
    "cm3p_muon_momentum": (_I, [_P, _P, _P, _L, _F, _I, _P, _P]),
    "cm3p_bf16_normalize": (_I, [_P, _L, _P, _F, _P]),
    "cm3p_bf16_axpy": (_I, [_P, _L, _F, _P, _L, _P, _L, _L, _L, _P]),
    "cm3p_muon_apply": (_I, [_P, _P, _L, _F, _F, _P]),
    "cm3p_adamw_step": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m cm3p_b200.build` (nvcc, sm_100a). "
            "cm3p_b200 has no CPU or PyTorch fallback for its kernels.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().cm3p_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed with status {status}: {last_error()}")
