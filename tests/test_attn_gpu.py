"""Fused attention forward/backward vs the eager fp32 formula, bf16 tolerance 2e-2."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def ref_attn(qkv, B, N, H, d, scale):
    q, k, v = qkv.float().reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * scale
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B * N, H * d)
    lse2 = torch.logsumexp(s, -1) / math.log(2.0)
    return o, lse2


CASES = [(2, 197, 6, 64), (1, 128, 1, 64), (3, 37, 2, 64), (2, 785, 3, 64), (2, 196, 8, 48), (1, 145, 12, 64),
         (1, 1297, 2, 64), (2, 198, 3, 64), (4, 256, 2, 64), (2, 1, 2, 64)]


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("B,N,H,d", CASES + [(40, 197, 12, 64), (9, 300, 4, 48)])
def test_attn_fwd(B, N, H, d, variant, monkeypatch):
    """VITK_ATTN_WG2 = 1 (default): eight softmax warps per CTA, two online softmaxes merged per row; 0: the first
    kernel (four softmax warps)."""
    from vit_torch_b200 import ops
    monkeypatch.setenv("VITK_ATTN_WG2", str(variant))
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N)
    qkv = (torch.randn((B * N, 3 * H * d), device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    scale = d ** -0.5
    out, lse2 = ops.attn_fwd(qkv, B, N, H, d, scale)
    ro, rl = ref_attn(qkv, B, N, H, d, scale)
    e_o, e_l = nerr(out, ro), nerr(lse2, rl)
    print(f"attn_fwd B{B} N{N} H{H} d{d}: out nerr={e_o:.3e} lse nerr={e_l:.3e}")
    assert e_o <= 2e-2
    assert e_l <= 1e-3


@pytest.mark.parametrize("variant", ["head", "two_kernel_wg2", "two_kernel_wg1", "fused"])
@pytest.mark.parametrize("B,N,H,d", CASES + [(40, 197, 12, 64), (3, 129, 2, 64), (2, 64, 2, 64), (2, 200, 1, 64)])
def test_attn_bwd(B, N, H, d, variant, monkeypatch):
    """head: one block per (image, head) computes dQ, dK, dV in a single pass (default for N <= 256, d = 64);
    two_kernel_wg2: the two deterministic kernels with eight elementwise warps per CTA (default otherwise);
    two_kernel_wg1: the four-warp kernels (VITK_ATTN_WG2=0); fused: the single-kernel backward with fp32 dQ atomics."""
    from vit_torch_b200 import _lib, ops
    fused = variant == "fused"
    if fused and d != 64:
        pytest.skip("single-kernel backward is d = 64 only")
    monkeypatch.setenv("VITK_ATTN_BWD_HEAD", "1" if variant == "head" else "0")
    if variant == "head" and not _lib.load().vitk_attn_bwd_head_supported(N, d):
        pytest.skip("whole-head backward serves N <= 256, d = 64")
    monkeypatch.setenv("VITK_ATTN_WG2", "0" if variant == "two_kernel_wg1" else "1")
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N + 7)
    qkv = (torch.randn((B * N, 3 * H * d), device="cuda", generator=g) * 1.2).to(torch.bfloat16)
    dout = torch.randn((B * N, H * d), device="cuda", generator=g).to(torch.bfloat16)
    scale = d ** -0.5
    out, lse2 = ops.attn_fwd(qkv, B, N, H, d, scale)
    dbias = torch.ones((3 * H * d,), device="cuda")            # += column sums of dqkv (qkv Linear bias gradient)
    dqkv = ops.attn_bwd(qkv, out, dout, lse2, B, N, H, d, scale, fused=fused, dbias=dbias)
    assert nerr(dbias - 1, dqkv.float().sum(0)) <= 5e-3        # fp32 tile sums vs sums of the bf16-rounded outputs
    x = qkv.float().requires_grad_(True)
    ro, _ = ref_attn(x, B, N, H, d, scale)
    ro.backward(dout.float())
    ref = x.grad.reshape(B * N, 3, H * d)
    got = dqkv.float().reshape(B * N, 3, H * d)
    for i, name in enumerate("qkv"):
        e = nerr(got[:, i], ref[:, i])
        print(f"attn_bwd B{B} N{N} H{H} d{d}: d{name} nerr={e:.3e}")
        assert e <= 2e-2
