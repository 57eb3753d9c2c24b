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
