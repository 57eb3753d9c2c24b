"""Tensor-level wrappers over the C ABI (include/vitk.h). PyTorch is only used for device memory and streams.

Every function launches hand-written sm_100a kernels through libvitk.so on the current torch CUDA stream and raises
if the library is missing or the tensors are not CUDA tensors -- there is no eager/CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr

EPI_STORE_BF16 = 0
EPI_BIAS_GELU = 1
EPI_RESID_F32 = 2
EPI_DGELU = 3
EPI_ATOMIC_F32 = 4
EPI_STORE_F32 = 5

# number of libvitk kernel launches issued through this module (bench.py reports it as gpu_launches)
launch_count = 0


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VitkError("vit_torch_b200 ops need CUDA tensors (no CPU fallback)")


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "matrix must be row-major with unit inner stride"
    return t.stride(0)


def gemm(a, b, *, a_mn=False, b_mn=False, M=None, N=None, K=None, epilogue=EPI_STORE_BF16, bias=None, gamma=None,
         resid=None, out=None, out2=None, aux=None, splits=0):
    """D[M,N] = A[M,K] @ B[N,K]^T with a fused epilogue. See vitk_gemm_bf16 in include/vitk.h.

    a: bf16 [M,K] (a_mn=False) or [K,M] (a_mn=True); b: bf16 [N,K] (b_mn=False) or [K,N] (b_mn=True).
    """
    global launch_count
    _need_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    lib = _lib.load()
    rc = lib.vitk_gemm_bf16(
        ptr(a), _ld(a), int(a_mn), ptr(b), _ld(b), int(b_mn), M, N, K, epilogue,
        ptr(bias), ptr(gamma), ptr(resid), _ld(resid) if resid is not None else 0,
        ptr(out), _ld(out), ptr(out2), _ld(out2) if out2 is not None else 0,
        ptr(aux), _ld(aux) if aux is not None else 0, splits, _stream())
    check(rc, "vitk_gemm_bf16")
    launch_count += 1
    return out


def layernorm_fwd(x, weight, bias, eps=1e-6):
    """x fp32 [rows, D] -> (y bf16 [rows, D], mean fp32 [rows], rstd fp32 [rows])."""
    global launch_count
    _need_cuda(x, weight, bias)
    assert x.dtype == torch.float32 and x.is_contiguous()
    rows, D = x.shape
    y = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    check(lib.vitk_layernorm_fwd(ptr(x), ptr(weight), ptr(bias), ptr(y), ptr(mean), ptr(rstd), rows, D, eps, _stream()),
          "vitk_layernorm_fwd")
    launch_count += 1
    return y, mean, rstd


def layernorm_bwd(dy, x, weight, mean, rstd, *, dres=None, dweight=None, dbias=None, want_f32=True, want_bf16=False,
                  colscale=None):
    """Returns (dx_f32 or None, dx_bf16 or None); accumulates into dweight/dbias (fp32 [D]) when given."""
    global launch_count
    _need_cuda(dy, x)
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and x.is_contiguous()
    rows, D = x.shape
    dx = torch.empty_like(x) if want_f32 else None
    dxb = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    lib = _lib.load()
    check(lib.vitk_layernorm_bwd(ptr(dy), ptr(x), ptr(weight), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dxb),
                                 ptr(colscale), ptr(dweight), ptr(dbias), rows, D, _stream()), "vitk_layernorm_bwd")
    launch_count += 1
    return dx, dxb


def colsum_accum(x, out):
    """out[N] (fp32) += column sums of bf16 x [rows, N]."""
    global launch_count
    _need_cuda(x, out)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32
    lib = _lib.load()
    check(lib.vitk_colsum_bf16(ptr(x), _ld(x), x.shape[0], x.shape[1], ptr(out), _stream()), "vitk_colsum_bf16")
    launch_count += 1
    return out


def cast_bf16(x, out=None):
    """fp32 -> bf16 copy (any shape, contiguous)."""
    global launch_count
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    check(lib.vitk_cast_f32_bf16(ptr(x), ptr(out), x.numel(), _stream()), "vitk_cast_f32_bf16")
    launch_count += 1
    return out


def attn_fwd(qkv, B, N, H, d, scale):
    """qkv bf16 [B*N, 3*H*d] -> (out bf16 [B*N, H*d], lse2 fp32 [B, H, N])."""
    global launch_count
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.numel() == B * N * 3 * H * d
    out = torch.empty((B * N, H * d), dtype=torch.bfloat16, device=qkv.device)
    lse2 = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    lib = _lib.load()
    check(lib.vitk_attn_fwd(ptr(qkv), ptr(out), ptr(lse2), B, N, H, d, scale, _stream()), "vitk_attn_fwd")
    launch_count += 1
    return out, lse2


def attn_bwd(qkv, out, dout, lse2, B, N, H, d, scale):
    """Returns dqkv bf16 [B*N, 3*H*d] (dQ | dK | dV in the qkv Linear output layout)."""
    global launch_count
    _need_cuda(qkv, out, dout, lse2)
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous() and out.is_contiguous() and qkv.is_contiguous()
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse2)
    lib = _lib.load()
    check(lib.vitk_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse2), ptr(delta), ptr(dqkv), B, N, H, d, scale,
                            _stream()), "vitk_attn_bwd")
    launch_count += 2
    return dqkv
