"""The C-ABI shared library loads and exports every symbol include/vitk.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vitk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(vitk_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vit_torch_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vitk.h but not exported"
    assert lib.vitk_abi_version() >= 1


def test_python_signatures_cover_the_header():
    from vit_torch_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from vit_torch_b200 import _lib, ops
    x = torch.zeros(8, 8)
    with pytest.raises(_lib.VitkError):
        ops.layernorm_fwd(x, torch.ones(8), torch.zeros(8))
    with pytest.raises(_lib.VitkError):
        ops.cast_bf16(x)


def test_bad_arguments_return_error_codes():
    from vit_torch_b200 import _lib
    lib = _lib.load()
    # null pointers / bad shapes are rejected before any launch (safe without a GPU)
    assert lib.vitk_layernorm_fwd(None, None, None, None, None, None, 4, 768, 1e-6, None) == -1
    assert lib.vitk_colsum_bf16(None, 8, 4, 8, None, None) == -1
    assert lib.vitk_attn_fwd(None, None, None, 1, 197, 6, 32, 0.125, None) == -1
    assert lib.vitk_patchify(None, None, 1, 3, 224, 224, 16, None) == -1


def test_talking_heads_products_reject_bad_shapes():
    """vitk_th_scores / vitk_th_apply validate before any launch: long sequences, odd head sizes, unaligned pitches."""
    from vit_torch_b200 import _lib
    lib = _lib.load()
    assert lib.vitk_th_gemm_supported(196, 48, 200) == 1 and lib.vitk_th_gemm_supported(197, 64, 200) == 1
    assert lib.vitk_th_gemm_supported(577, 48, 584) == 0        # CaiT at 384 px: generic batched GEMM instead
    assert lib.vitk_th_gemm_supported(196, 32, 200) == 0        # head size
    assert lib.vitk_th_gemm_supported(196, 48, 196) == 0        # plane pitch must be a multiple of 8
    assert lib.vitk_th_scores(None, 1152, 1152, 0, None, 1152, 1152, 384, None, 0, 1, 196, 8, 48, 200, None) == -1
    assert lib.vitk_th_apply(None, None, 1152, 1152, 768, None, 384, 0, 0, None, 1, 196, 8, 48, 200, None) == -1
    assert lib.vitk_th_mix_fwd_s16(None, None, None, None, None, 0.1, None, None, None, 1, 8, 196, 200, None) == -1
