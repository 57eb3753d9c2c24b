"""Tensor-level wrappers over the C ABI (include/vitk.h). PyTorch is only used for device memory and streams.

Every function launches hand-written sm_100a kernels through libvitk.so on the current torch CUDA stream and raises
if the library is missing or the tensors are not CUDA tensors -- there is no eager/CPU fallback.
"""
from __future__ import annotations

import os

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


# optional per-launch CUDA-event timing of the GEMM kernel (bench.py roofline pass)
_gemm_events = None


def gemm_timing_begin():
    global _gemm_events
    _gemm_events = []


def gemm_timing_end():
    """Returns (total GEMM kernel milliseconds, number of GEMM launches) since gemm_timing_begin()."""
    global _gemm_events
    ev, _gemm_events = _gemm_events, None
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in ev), len(ev)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VitkError("vit_torch_b200 ops need CUDA tensors (no CPU fallback)")


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "matrix must be row-major with unit inner stride"
    return t.stride(0)


EPI_TOKENS_F32 = 6


def gemm(a, b, *, a_mn=False, b_mn=False, M=None, N=None, K=None, epilogue=EPI_STORE_BF16, bias=None, gamma=None,
         resid=None, out=None, out2=None, aux=None, splits=0, rowscale=None, rows_per_sample=0, tok=None, colsum=None):
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
    tok_n, tok_N, tok_T = tok if tok is not None else (0, 0, 0)
    if _gemm_events is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = lib.vitk_gemm_bf16_ex(
        ptr(a), _ld(a), int(a_mn), ptr(b), _ld(b), int(b_mn), M, N, K, epilogue,
        ptr(bias), ptr(gamma), ptr(resid), _ld(resid) if resid is not None else 0,
        ptr(out), _ld(out) if out is not None else 0, ptr(out2), _ld(out2) if out2 is not None else 0,
        ptr(aux), _ld(aux) if aux is not None else 0, splits, ptr(rowscale), rows_per_sample, tok_n, tok_N, tok_T,
        ptr(colsum), _stream())
    check(rc, "vitk_gemm_bf16_ex")
    if _gemm_events is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        _gemm_events.append((ev0, ev1))
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
                  colscale=None, dxsum=None):
    """Returns (dx_f32 or None, dx_bf16 or None); accumulates into dweight/dbias (fp32 [D]) when given, and the column
    sums of the bf16 copy into dxsum."""
    global launch_count
    _need_cuda(dy, x)
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and x.is_contiguous()
    rows, D = x.shape
    dx = torch.empty_like(x) if want_f32 else None
    dxb = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    lib = _lib.load()
    check(lib.vitk_layernorm_bwd_ex(ptr(dy), 0, ptr(x), D, ptr(weight), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), D,
                                    ptr(dxb), ptr(colscale), ptr(dweight), ptr(dbias), ptr(dxsum), rows, D, _stream()),
          "vitk_layernorm_bwd_ex")
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


def cast_f32_from_bf16(x, out):
    """bf16 -> fp32 copy into `out` (same number of elements)."""
    global launch_count
    _need_cuda(x, out)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and x.numel() == out.numel()
    check(_lib.load().vitk_cast_bf16_f32(ptr(x), ptr(out), x.numel(), _stream()), "vitk_cast_bf16_f32")
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


# The single-kernel backward (attn_bwd_fused.cu) is parity-tested but not the default: at B200 it measures 942 us vs
# 950 us (two kernels) on ViT-B/8 bs64 and 291 us vs 240 us on ViT-B/16 bs128 (profiles/r01_summary_v3.md).
_ATTN_BWD_FUSED = __import__("os").environ.get("VITK_ATTN_BWD_FUSED", "0") == "1"


def attn_bwd(qkv, out, dout, lse2, B, N, H, d, scale, fused=None, dbias=None):
    """Returns dqkv bf16 [B*N, 3*H*d] (dQ | dK | dV in the qkv Linear output layout). fused=True selects the
    single-kernel backward (d = 64), fused=None follows VITK_ATTN_BWD_FUSED. dbias (fp32 [3*H*d], optional): += column
    sums of dqkv = gradient of the qkv Linear bias, produced by the two-kernel backward itself."""
    global launch_count
    _need_cuda(qkv, out, dout, lse2)
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous() and out.is_contiguous() and qkv.is_contiguous()
    dqkv = torch.empty_like(qkv)
    lib = _lib.load()
    if not fused and lib.vitk_attn_bwd_head_supported(N, d):
        # opt-in (VITK_ATTN_BWD_HEAD=1): one block per (image, head) produces dQ, dK, dV in a single pass
        check(lib.vitk_attn_bwd_head(ptr(qkv), ptr(out), ptr(dout), ptr(lse2), ptr(dqkv), ptr(dbias), B, N, H, d, scale,
                                     _stream()), "vitk_attn_bwd_head")
        launch_count += 1
        return dqkv
    delta = torch.empty_like(lse2)
    if d == 64 and N <= 8192 and (_ATTN_BWD_FUSED if fused is None else fused):
        # single-kernel backward; dQ tiles of different key blocks meet in an fp32 workspace (zeroed by the library)
        dq32 = torch.empty((B * N, H * d), dtype=torch.float32, device=qkv.device)
        check(lib.vitk_attn_bwd_fused(ptr(qkv), ptr(out), ptr(dout), ptr(lse2), ptr(delta), ptr(dq32), ptr(dqkv), B, N,
                                      H, d, scale, _stream()), "vitk_attn_bwd_fused")
        launch_count += 3
        if dbias is not None:
            colsum_accum(dqkv, dbias)
        return dqkv
    check(lib.vitk_attn_bwd_ex(ptr(qkv), ptr(out), ptr(dout), ptr(lse2), ptr(delta), ptr(dqkv), ptr(dbias), B, N, H, d,
                               scale, _stream()), "vitk_attn_bwd_ex")
    launch_count += 2
    return dqkv


def layernorm_fwd_rows(x, x_stride, rows, D, weight, bias, eps=1e-6, out_f32=False):
    """LayerNorm over `rows` rows of width D spaced x_stride elements apart in fp32 tensor x.
    Returns (y [rows, D] bf16 or fp32, mean, rstd)."""
    global launch_count
    _need_cuda(x, weight, bias)
    assert x.dtype == torch.float32
    y = torch.empty((rows, D), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    check(lib.vitk_layernorm_fwd_ex(ptr(x), x_stride, ptr(weight), ptr(bias), None if out_f32 else ptr(y),
                                    ptr(y) if out_f32 else None, ptr(mean), ptr(rstd), rows, D, eps, _stream()),
          "vitk_layernorm_fwd_ex")
    launch_count += 1
    return y, mean, rstd


def layernorm_bwd_rows(dy, x, x_stride, rows, D, weight, mean, rstd, *, dres=None, dx=None, dx_stride=None,
                       dx_bf16=None, colscale=None, dweight=None, dbias=None):
    """Strided LayerNorm backward; dy dense [rows, D] fp32 or bf16; writes dx (fp32, rows dx_stride apart) in place."""
    global launch_count
    _need_cuda(dy, x)
    assert dy.is_contiguous() and dy.dtype in (torch.float32, torch.bfloat16)
    lib = _lib.load()
    check(lib.vitk_layernorm_bwd_ex(ptr(dy), int(dy.dtype == torch.float32), ptr(x), x_stride, ptr(weight), ptr(mean),
                                    ptr(rstd), ptr(dres), ptr(dx), dx_stride if dx_stride is not None else D,
                                    ptr(dx_bf16), ptr(colscale), ptr(dweight), ptr(dbias), None, rows, D, _stream()),
          "vitk_layernorm_bwd_ex")
    launch_count += 1


def colsum_f32_accum(x, ldx, rows, cols, out):
    """out[c] (fp32) += sum_r x[r*ldx + c]."""
    global launch_count
    _need_cuda(x, out)
    lib = _lib.load()
    check(lib.vitk_colsum_f32(ptr(x), ldx, rows, cols, ptr(out), _stream()), "vitk_colsum_f32")
    launch_count += 1


def cross_entropy(logits, labels, want_grad=True):
    """logits fp32 [B, C] (rows may be strided), labels int64 [B] -> (out2 fp32 [2] = (mean loss, #correct),
    dlogits fp32 [B, C] = d(mean loss)/d(logits) or None). One launch, no host sync."""
    global launch_count
    _need_cuda(logits, labels)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    assert labels.dtype == torch.int64 and labels.is_contiguous() and labels.numel() == logits.shape[0]
    B, C = logits.shape
    out2 = torch.empty((2,), dtype=torch.float32, device=logits.device)
    dl = torch.empty((B, C), dtype=torch.float32, device=logits.device) if want_grad else None
    lib = _lib.load()
    check(lib.vitk_cross_entropy(ptr(logits), logits.stride(0), ptr(labels), B, C, ptr(out2), ptr(dl), C, _stream()),
          "vitk_cross_entropy")
    launch_count += 1
    return out2, dl


def colsum_prod_accum(a, b, out, rowscale=None, rows_per_sample=1):
    """out[N] += sum_r rowscale[r // rows_per_sample] * a_f32[r, :] * b_bf16[r, :]  (rowscale optional)."""
    global launch_count
    _need_cuda(a, b, out, rowscale)
    assert a.dtype == torch.float32 and b.dtype == torch.bfloat16
    lib = _lib.load()
    check(lib.vitk_colsum_prod_ex(ptr(a), _ld(a), ptr(b), _ld(b), a.shape[0], a.shape[1], ptr(out), ptr(rowscale),
                                  rows_per_sample, _stream()), "vitk_colsum_prod_ex")
    launch_count += 1


def layerscale_bwd(dy, f, gamma, rowscale=None, rows_per_sample=1, dgamma=None, dbias=None):
    """One pass over a LayerScale branch's gradient: returns bf16(dy * gamma * rowscale) [rows, D]; dgamma += sum dy *
    rowscale * f; dbias += column sums of the returned matrix (see vitk_layerscale_bwd)."""
    global launch_count
    _need_cuda(dy, f, gamma)
    assert dy.dtype == torch.float32 and f.dtype == torch.bfloat16 and dy.shape == f.shape
    rows, D = dy.shape
    out = torch.empty((rows, D), dtype=torch.bfloat16, device=dy.device)
    check(_lib.load().vitk_layerscale_bwd(ptr(dy), _ld(dy), ptr(f), _ld(f), ptr(gamma), ptr(rowscale), rows_per_sample, rows,
                                          D, ptr(out), ptr(dgamma), ptr(dbias), _stream()), "vitk_layerscale_bwd")
    launch_count += 1
    return out


def scale_cast(x, rows, D, *, rows_per_group=None, group_stride=0, colscale=None, rowscale=None, rows_per_sample=0,
               offset_elems=0):
    """bf16 [rows, D] = x (fp32) * colscale * rowscale with optional token-row compaction (see vitk_scale_cast)."""
    global launch_count
    _need_cuda(x)
    assert x.dtype == torch.float32
    out = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    if rows_per_group is None:
        rows_per_group, group_stride = rows, 0
    lib = _lib.load()
    check(lib.vitk_scale_cast(x.data_ptr() + 4 * offset_elems, rows_per_group, group_stride, rows, D, ptr(colscale),
                              ptr(rowscale), rows_per_sample, ptr(out), _stream()), "vitk_scale_cast")
    launch_count += 1
    return out


def patchify(x, P):
    """x fp32 [B,C,H,W] -> bf16 [B*(H/P)*(W/P), C*P*P]."""
    global launch_count
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    B, C, H, W = x.shape
    out = torch.empty((B * (H // P) * (W // P), C * P * P), dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    check(lib.vitk_patchify(ptr(x), ptr(out), B, C, H, W, P, _stream()), "vitk_patchify")
    launch_count += 1
    return out


def patch_embed_fwd(img, weight, bias, pos2d, out, P, T):
    """out[b, T + p, :] = patch_p . weight^T + bias + pos2d[T + p]; img fp32 / bf16 [B,C,H,W] (TMA gather from NCHW),
    weight [D, C*P*P] of img's dtype, out fp32 [B, N, D]."""
    global launch_count
    _need_cuda(img, weight, pos2d, out)
    assert img.is_contiguous() and weight.is_contiguous() and img.dtype == weight.dtype
    assert img.dtype in (torch.float32, torch.bfloat16)
    B, C, H, W = img.shape
    D = weight.shape[0]
    check(_lib.load().vitk_patch_embed_fwd(ptr(img), int(img.dtype == torch.bfloat16), ptr(weight), ptr(bias), ptr(pos2d),
                                           pos2d.stride(0), ptr(out), B, C, H, W, P, D, out.shape[1], T, _stream()),
          "vitk_patch_embed_fwd")
    launch_count += 1
    return out


def patch_embed_wgrad(img, dy, dw, P, T):
    """dw fp32 [D, C*P*P] += sum_{b,p} dy[b, T + p, :]^T patch_p; dy [B, N, D] of img's dtype."""
    global launch_count
    _need_cuda(img, dy, dw)
    assert img.is_contiguous() and dy.is_contiguous() and img.dtype == dy.dtype and dw.dtype == torch.float32
    B, C, H, W = img.shape
    check(_lib.load().vitk_patch_embed_wgrad(ptr(img), int(img.dtype == torch.bfloat16), ptr(dy), dy.shape[1], T, ptr(dw),
                                             B, C, H, W, P, dy.shape[2], _stream()), "vitk_patch_embed_wgrad")
    launch_count += 1


def normalize_u8(x, mean, std):
    """uint8 [B,C,H,W] -> bf16 (x/255 - mean[c]) / std[c] (ToTensor + Normalize on the device)."""
    global launch_count
    _need_cuda(x, mean, std)
    assert x.dtype == torch.uint8 and x.is_contiguous() and mean.dtype == torch.float32 and std.dtype == torch.float32
    B, C, H, W = x.shape
    y = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().vitk_normalize_u8(ptr(x), ptr(y), ptr(mean), ptr(std), B, C, H, W, _stream()), "vitk_normalize_u8")
    launch_count += 1
    return y


def prefix_tokens(tok, pos, out, B, T, tokens_per_image, D):
    """out[b, t, :] = tok[t] + pos[t] for t < T."""
    global launch_count
    _need_cuda(tok, pos, out)
    lib = _lib.load()
    check(lib.vitk_prefix_tokens(ptr(tok), ptr(pos), ptr(out), B, T, tokens_per_image, D, _stream()),
          "vitk_prefix_tokens")
    launch_count += 1


def gemm_batched(a, lda, sa_h, sa_b, a_mn, b, ldb, sb_h, sb_b, b_mn, M, N, K, nh, nb, out, ldo, so_h, so_b,
                 a_off=0, b_off=0, out_off=0, out_f32=False):
    """nh*nb independent GEMMs addressed by element strides (see vitk_gemm_bf16_batched). a/b/out are the base tensors;
    *_off are element offsets into them."""
    global launch_count
    _need_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    lib = _lib.load()
    osz = 4 if out_f32 else 2
    rc = lib.vitk_gemm_bf16_batched(a.data_ptr() + 2 * a_off, lda, sa_h, sa_b, int(a_mn), b.data_ptr() + 2 * b_off, ldb,
                                    sb_h, sb_b, int(b_mn), M, N, K, nh, nb,
                                    EPI_STORE_F32 if out_f32 else EPI_STORE_BF16, out.data_ptr() + osz * out_off, ldo,
                                    so_h, so_b, _stream())
    check(rc, "vitk_gemm_bf16_batched")
    launch_count += 1
    return out


def th_gemm_ok(N, d, Np) -> bool:
    """True when the short-sequence talking-heads products (th_scores / th_apply) serve this shape.
    VITK_TH_GEMM=0 keeps the generic batched GEMM (A/B runs)."""
    if os.environ.get("VITK_TH_GEMM", "1") == "0":
        return False
    return bool(_lib.load().vitk_th_gemm_supported(int(N), int(d), int(Np)))


def th_scores(a, a_col0, b, b_col0, B, N, H, d, Np, out_f32=False):
    """out[b,h,i,j] = a_h[i] . b_h[j] over token-major bf16 matrices a, b ([B*N, cols], head h = columns col0 + h*d ..).
    -> plane [B,H,N,Np], bf16 (the dtype torch.autocast gives the reference's q @ k^T) or fp32."""
    global launch_count
    _need_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    out = torch.empty((B, H, N, Np), dtype=torch.float32 if out_f32 else torch.bfloat16, device=a.device)
    check(_lib.load().vitk_th_scores(ptr(a), a.stride(0), a.shape[1], a_col0, ptr(b), b.stride(0), b.shape[1], b_col0,
                                     ptr(out), int(out_f32), B, N, H, d, Np, _stream()), "vitk_th_scores")
    launch_count += 1
    return out


def th_apply(p, x, x_col0, out, o_col0, B, N, H, d, Np, transpose=False, colsum=None):
    """out[b*N+i, o_col0+h*d+e] = sum_j p[b,h,i,j] x[b*N+j, x_col0+h*d+e]  (transpose: sum over i of p[b,h,i,j] x[b*N+i, ..]
    into row j). p: bf16 plane [B,H,N,Np] with zero pad columns; x, out: token-major bf16. colsum (fp32 [H*d]) += the
    column sums of the written output."""
    global launch_count
    _need_cuda(p, x, out)
    assert p.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    if colsum is not None:
        _need_cuda(colsum)
        assert colsum.dtype == torch.float32 and colsum.is_contiguous() and colsum.numel() == H * d
    check(_lib.load().vitk_th_apply(ptr(p), ptr(x), x.stride(0), x.shape[1], x_col0, ptr(out), out.stride(0), o_col0,
                                    int(transpose), ptr(colsum), B, N, H, d, Np, _stream()), "vitk_th_apply")
    launch_count += 1
    return out


def th_apply_t(p, x, x_col0, out, o_col0, B, N, H, d, Np, colsum=None):
    """th_apply with the plane transposed: out[b*N+j] = sum_i p[b,h,i,j] x[b*N+i]."""
    return _th_apply_impl(p, x, x_col0, out, o_col0, B, N, H, d, Np, transpose=True, colsum=colsum)


def th_mix_fwd(S, wl, bl, ww, bw, scale, B, H, N, Np):
    """Talking-heads mixing forward. S fp32 or bf16 [B,H,N,Np] -> (Pm bf16 [B,H,N,Np], rowmax, rowsum fp32 [B,H,N])."""
    global launch_count
    _need_cuda(S, wl, bl, ww, bw)
    Pm = torch.empty((B, H, N, Np), dtype=torch.bfloat16, device=S.device)
    rmax = torch.empty((B, H, N), dtype=torch.float32, device=S.device)
    rsum = torch.empty((B, H, N), dtype=torch.float32, device=S.device)
    lib = _lib.load()
    fn, name = ((lib.vitk_th_mix_fwd_s16, "vitk_th_mix_fwd_s16") if S.dtype == torch.bfloat16
                else (lib.vitk_th_mix_fwd, "vitk_th_mix_fwd"))
    check(fn(ptr(S), ptr(wl), ptr(bl), ptr(ww), ptr(bw), scale, ptr(Pm), ptr(rmax), ptr(rsum), B, H, N, Np, _stream()),
          name)
    launch_count += 1
    return Pm, rmax, rsum


def th_mix_bf16_dp(Np) -> bool:
    """True when the talking-heads backward takes dP' in bf16 (version-2 kernels, rows of up to 208 keys)."""
    return bool(_lib.load().vitk_th_mix_supports_bf16_dp(int(Np)))


def th_mix_bwd(S, dPm, rmax, rsum, wl, bl, ww, bw, scale, dwl, dbl, dww, dbw, B, H, N, Np):
    """Talking-heads mixing backward -> dS bf16 [B,H,N,Np]; accumulates into dwl/dbl/dww/dbw. dPm fp32, or bf16 when
    th_mix_bf16_dp(Np)."""
    global launch_count
    _need_cuda(S, dPm)
    dS = torch.empty((B, H, N, Np), dtype=torch.bfloat16, device=S.device)
    lib = _lib.load()
    if S.dtype == torch.bfloat16:
        assert dPm.dtype == torch.bfloat16
        check(lib.vitk_th_mix_bwd_s16(ptr(S), ptr(dPm), ptr(rmax), ptr(rsum), ptr(wl), ptr(bl), ptr(ww), ptr(bw), scale,
                                      ptr(dS), ptr(dwl), ptr(dbl), ptr(dww), ptr(dbw), B, H, N, Np, _stream()),
              "vitk_th_mix_bwd_s16")
        launch_count += 1
        return dS
    if dPm.dtype == torch.bfloat16:
        check(lib.vitk_th_mix_bwd_bf16(ptr(S), ptr(dPm), ptr(rmax), ptr(rsum), ptr(wl), ptr(bl), ptr(ww), ptr(bw), scale,
                                       ptr(dS), ptr(dwl), ptr(dbl), ptr(dww), ptr(dbw), B, H, N, Np, _stream()),
              "vitk_th_mix_bwd_bf16")
        launch_count += 1
        return dS
    check(lib.vitk_th_mix_bwd(ptr(S), ptr(dPm), ptr(rmax), ptr(rsum), ptr(wl), ptr(bl), ptr(ww), ptr(bw), scale, ptr(dS),
                              ptr(dwl), ptr(dbl), ptr(dww), ptr(dbw), B, H, N, Np, _stream()), "vitk_th_mix_bwd")
    launch_count += 1
    return dS


def class_attn_fwd(q, kc, kx, vc, vx, ldkv, ldc, scale, B, H, n, d):
    """Class attention forward -> (out bf16 [B, H*d], p fp32 [B, H, n+1])."""
    global launch_count
    _need_cuda(q, kc, kx, vc, vx)
    out = torch.empty((B, H * d), dtype=torch.bfloat16, device=q.device)
    p = torch.empty((B, H, n + 1), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    check(lib.vitk_class_attn_fwd(ptr(q), ptr(kc), ptr(kx), ptr(vc), ptr(vx), ldkv, ldc, scale, ptr(out), ptr(p), B, H, n, d,
                                  _stream()), "vitk_class_attn_fwd")
    launch_count += 1
    return out, p


def class_attn_bwd(q, kc, kx, vc, vx, ldkv, ldc, p, dout, scale, dq, dkc, dkx, dvc, dvx, lddkv, lddc, B, H, n, d):
    global launch_count
    _need_cuda(q, dout)
    lib = _lib.load()
    check(lib.vitk_class_attn_bwd(ptr(q), ptr(kc), ptr(kx), ptr(vc), ptr(vx), ldkv, ldc, ptr(p), ptr(dout), scale,
                                  ptr(dq), ptr(dkc), ptr(dkx), ptr(dvc), ptr(dvx), lddkv, lddc, B, H, n, d, _stream()),
          "vitk_class_attn_bwd")
    launch_count += 1


# ------------------------------------------------------------------------------------------------------------------
# optional per-op CUDA-event timing of every wrapper in this module (scripts/step_breakdown.py)
# ------------------------------------------------------------------------------------------------------------------
_op_events = None


def op_timing_begin():
    global _op_events
    _op_events = []


def op_timing_end():
    """Returns {op name: (total ms, launches)} since op_timing_begin()."""
    global _op_events
    ev, _op_events = _op_events, None
    torch.cuda.synchronize()
    out = {}
    for name, s, e in ev:
        t, n = out.get(name, (0.0, 0))
        out[name] = (t + s.elapsed_time(e), n + 1)
    return out


def _timed(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if _op_events is None:
            return fn(*a, **k)
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn(*a, **k)
        e.record()
        _op_events.append((fn.__name__, s, e))
        return r

    return wrapper


for _name in ("gemm", "gemm_batched", "layernorm_fwd", "layernorm_bwd", "layernorm_fwd_rows", "layernorm_bwd_rows",
              "colsum_accum", "colsum_f32_accum", "colsum_prod_accum", "layerscale_bwd", "cast_bf16", "attn_fwd", "attn_bwd", "scale_cast",
              "patchify", "prefix_tokens", "patch_embed_fwd", "patch_embed_wgrad", "normalize_u8", "th_mix_fwd", "th_mix_bwd", "class_attn_fwd", "class_attn_bwd",
              "th_scores", "th_apply_t"):
    globals()[_name] = _timed(globals()[_name])
_th_apply_impl = th_apply
th_apply = _timed(th_apply)          # (th_apply_t calls the untimed implementation: one event pair per launch)


def sgd_momentum_multi(table, chunk_map, num_chunks, lr, momentum, grad_scale=1.0, first_step=False):
    """One launch of the fused multi-tensor SGD(momentum) + bf16 weight refresh. See vitk_sgd_momentum_multi."""
    global launch_count
    _need_cuda(table, chunk_map)
    lib = _lib.load()
    check(lib.vitk_sgd_momentum_multi(ptr(table), ptr(chunk_map), int(num_chunks), float(lr), float(momentum),
                                      float(grad_scale), int(bool(first_step)), _stream()), "vitk_sgd_momentum_multi")
    launch_count += 1


sgd_momentum_multi = _timed(sgd_momentum_multi) if "_timed" in globals() else sgd_momentum_multi


def sgd_momentum_multi_hp(table, chunk_map, num_chunks, hyper):
    """As sgd_momentum_multi with (lr, momentum, grad_scale) read from the device tensor `hyper` (fp32 [3])."""
    global launch_count
    _need_cuda(table, chunk_map, hyper)
    lib = _lib.load()
    check(lib.vitk_sgd_momentum_multi_hp(ptr(table), ptr(chunk_map), num_chunks, ptr(hyper), _stream()),
          "vitk_sgd_momentum_multi_hp")
    launch_count += 1


def adam_multi(table, chunk_map, num_chunks, hyper):
    """Multi-tensor Adam / AdamW step; `hyper` fp32 [11] on the device holds the hyper-parameters and the step counter
    (include/vitk.h). Two launches (tick + update)."""
    global launch_count
    _need_cuda(table, chunk_map, hyper)
    lib = _lib.load()
    check(lib.vitk_adam_multi(ptr(table), ptr(chunk_map), num_chunks, ptr(hyper), _stream()), "vitk_adam_multi")
    launch_count += 2


def set_sm_limit(n: int) -> None:
    """Size persistent-kernel grids for at most n SMs (0 = all). See vitk_set_sm_limit."""
    check(_lib.load().vitk_set_sm_limit(int(n)), "vitk_set_sm_limit")
