"""ctypes loader for libvitk.so (the C ABI in include/vitk.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VITK_LIB") or os.path.join(_HERE, "libvitk.so")

ERRORS = {
    -1: "VITK_ERR_ARG (bad shape / alignment / null pointer)",
    -2: "VITK_ERR_UNSUPPORTED (combination not compiled in)",
    -3: "VITK_ERR_CUDA (kernel launch failed)",
    -4: "VITK_ERR_TMAP (cuTensorMapEncodeTiled failed)",
}

# name -> argtypes; every function returns int
_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float
SIGNATURES = {
    "vitk_abi_version": [],
    "vitk_set_sm_limit": [_I],
    "vitk_gemm_bf16": [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I, _P],
    "vitk_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    "vitk_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "vitk_gemm_bf16_ex": [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I,
                          _P, _I, _I, _I, _I, _P, _P],
    "vitk_gemm_bf16_batched": [_P, _L, _L, _L, _I, _P, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _P],
    "vitk_layernorm_fwd_ex": [_P, _L, _P, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    "vitk_layernorm_bwd_ex": [_P, _I, _P, _L, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _L, _I, _P],
    "vitk_colsum_f32": [_P, _L, _L, _L, _P, _P],
    "vitk_cross_entropy": [_P, _L, _P, _I, _I, _P, _P, _L, _P],
    "vitk_colsum_prod": [_P, _L, _P, _L, _L, _I, _P, _P],
    "vitk_colsum_prod_ex": [_P, _L, _P, _L, _L, _I, _P, _P, _L, _P],
    "vitk_layerscale_bwd": [_P, _L, _P, _L, _P, _P, _L, _L, _I, _P, _P, _P, _P],
    "vitk_scale_cast": [_P, _L, _L, _L, _I, _P, _P, _L, _P, _P],
    "vitk_patchify": [_P, _P, _I, _I, _I, _I, _I, _P],
    "vitk_patch_embed_fwd": [_P, _I, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "vitk_patch_embed_wgrad": [_P, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "vitk_normalize_u8": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vitk_prefix_tokens": [_P, _P, _P, _I, _I, _L, _I, _P],
    "vitk_colsum_bf16": [_P, _L, _L, _I, _P, _P],
    "vitk_cast_f32_bf16": [_P, _P, _L, _P],
    "vitk_cast_bf16_f32": [_P, _P, _L, _P],
    "vitk_sgd_chunk_elems": [],
    "vitk_sgd_momentum_multi": [_P, _P, _I, _F, _F, _F, _I, _P],
    "vitk_sgd_momentum_multi_hp": [_P, _P, _I, _P, _P],
    "vitk_adam_multi": [_P, _P, _I, _P, _P],
    "vitk_attn_fwd": [_P, _P, _P, _I, _I, _I, _I, _F, _P],
    "vitk_th_mix_fwd": [_P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "vitk_th_mix_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vitk_th_mix_supports_bf16_dp": [_I],
    "vitk_th_mix_bwd_bf16": [_P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vitk_th_mix_fwd_s16": [_P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "vitk_th_mix_bwd_s16": [_P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vitk_th_gemm_supported": [_I, _I, _I],
    "vitk_th_scores": [_P, _L, _I, _I, _P, _L, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "vitk_th_apply": [_P, _P, _L, _I, _I, _P, _L, _I, _I, _P, _I, _I, _I, _I, _I, _P],
    "vitk_class_attn_fwd": [_P, _P, _P, _P, _P, _L, _L, _F, _P, _P, _I, _I, _I, _I, _P],
    "vitk_class_attn_bwd": [_P, _P, _P, _P, _P, _L, _L, _P, _P, _F, _P, _P, _P, _P, _P, _L, _L, _I, _I, _I, _I, _P],
    "vitk_attn_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "vitk_attn_bwd_ex": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "vitk_attn_bwd_head_supported": [_I, _I],
    "vitk_attn_bwd_head": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "vitk_attn_bwd_fused": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
}


class VitkError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load libvitk.so. Raises VitkError if it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build

            _build.build()
        else:
            raise VitkError(
                f"{LIB_PATH} not found: build it with `python -m vit_torch_b200.build` "
                "(the CUDA extension is mandatory; there is no CPU/PyTorch fallback)"
            )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise VitkError(f"{what} failed: {ERRORS.get(rc, rc)}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
