"""Short-sequence talking-heads products (csrc/th_gemm.cu: q k^T / P' v / P'^T dO of models/cait.py:116,125 and their
autograd) through the C ABI against fp32 einsums of the same bf16-rounded operands; bf16-plane mixing kernels against the
fp32-plane ones."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(3, 196, 8, 48), (2, 197, 4, 64), (2, 50, 6, 48), (1, 208, 8, 48), (2, 129, 2, 64), (2, 1, 4, 48), (2, 128, 16, 48)]


def _pad8(n):
    return (n + 7) // 8 * 8


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,N,H,d", SHAPES)
@pytest.mark.parametrize("out_f32", [False, True])
def test_th_scores_matches_einsum(B, N, H, d, out_f32):
    from vit_torch_b200 import ops
    torch.manual_seed(N + H)
    D, Np = H * d, _pad8(N)
    assert ops.th_gemm_ok(N, d, Np)
    qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
    S = ops.th_scores(qkv, 0, qkv, D, B, N, H, d, Np, out_f32=out_f32)
    q = qkv[:, :D].float().view(B, N, H, d)
    k = qkv[:, D:2 * D].float().view(B, N, H, d)
    ref = torch.einsum("bihe,bjhe->bhij", q, k)
    assert S.shape == (B, H, N, Np) and S.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    assert nerr(S[..., :N], ref) <= (1e-5 if out_f32 else 5e-3)
    assert S[..., N:].abs().max().item() == 0 if Np > N else True
    # second use: dP' = dO v^T (operands from two different matrices)
    do = torch.randn(B * N, D, device="cuda").bfloat16()
    dP = ops.th_scores(do, 0, qkv, 2 * D, B, N, H, d, Np, out_f32=out_f32)
    v = qkv[:, 2 * D:].float().view(B, N, H, d)
    ref2 = torch.einsum("bihe,bjhe->bhij", do.float().view(B, N, H, d), v)
    assert nerr(dP[..., :N], ref2) <= (1e-5 if out_f32 else 5e-3)


@pytest.mark.parametrize("B,N,H,d", SHAPES)
def test_th_apply_matches_einsum(B, N, H, d):
    from vit_torch_b200 import ops
    torch.manual_seed(N * 3 + H)
    D, Np = H * d, _pad8(N)
    qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
    P = torch.zeros(B, H, N, Np, device="cuda")
    P[..., :N] = torch.randn(B, H, N, N, device="cuda")
    P = P.bfloat16()
    Pf = P[..., :N].float()
    # O = P v
    o = torch.full((B * N, D), 7.0, device="cuda").bfloat16()
    ops.th_apply(P, qkv, 2 * D, o, 0, B, N, H, d, Np)
    v = qkv[:, 2 * D:].float().view(B, N, H, d)
    ref = torch.einsum("bhij,bjhe->bihe", Pf, v).reshape(B * N, D)
    assert nerr(o, ref) <= 5e-3
    # transposed products into column slices of one [B*N, 3D] buffer: dV = P^T dO (slice 2), dK = P^T q (slice 1)
    do = torch.randn(B * N, D, device="cuda").bfloat16()
    dqkv = torch.zeros(B * N, 3 * D, device="cuda").bfloat16()
    dbias = torch.full((3 * D,), 0.5, device="cuda")        # column sums are ACCUMULATED into the slices
    ops.th_apply(P, do, 0, dqkv, 2 * D, B, N, H, d, Np, transpose=True, colsum=dbias[2 * D:])
    ops.th_apply_t(P, qkv, 0, dqkv, D, B, N, H, d, Np, colsum=dbias[D:2 * D])
    ops.th_apply(P, qkv, D, dqkv, 0, B, N, H, d, Np, colsum=dbias[:D])
    q = qkv[:, :D].float().view(B, N, H, d)
    k = qkv[:, D:2 * D].float().view(B, N, H, d)
    ref_dv = torch.einsum("bhij,bihe->bjhe", Pf, do.float().view(B, N, H, d)).reshape(B * N, D)
    ref_dk = torch.einsum("bhij,bihe->bjhe", Pf, q).reshape(B * N, D)
    ref_dq = torch.einsum("bhij,bjhe->bihe", Pf, k).reshape(B * N, D)
    assert nerr(dqkv[:, 2 * D:], ref_dv) <= 5e-3
    assert nerr(dqkv[:, D:2 * D], ref_dk) <= 5e-3
    assert nerr(dqkv[:, :D], ref_dq) <= 5e-3
    ref_bias = torch.cat([ref_dq.sum(0), ref_dk.sum(0), ref_dv.sum(0)]) + 0.5
    assert nerr(dbias, ref_bias) <= 1e-3


def test_th_gemm_rejects_long_sequences():
    from vit_torch_b200 import ops
    assert not ops.th_gemm_ok(577, 48, 584)      # CaiT at 384 px keeps the generic batched GEMM
    assert not ops.th_gemm_ok(196, 32, 200)


@pytest.mark.parametrize("B,N,H", [(2, 196, 8), (2, 50, 4), (1, 197, 6), (1, 130, 16)])
def test_th_mix_bf16_planes_match_fp32_planes(B, N, H):
    """The mixing kernels on bf16 logit planes give what the fp32-plane kernels give on the same (bf16-rounded) logits."""
    from vit_torch_b200 import ops
    torch.manual_seed(H)
    Np = _pad8(N)
    S32 = torch.zeros(B, H, N, Np, device="cuda")
    S32[..., :N] = torch.randn(B, H, N, N, device="cuda") * 3
    S16 = S32.bfloat16()
    S32 = S16.float()
    wl = torch.randn(H, H, device="cuda") * 0.3 + torch.eye(H, device="cuda")
    ww = torch.randn(H, H, device="cuda") * 0.3 + torch.eye(H, device="cuda")
    bl = torch.randn(H, device="cuda") * 0.1
    bw = torch.randn(H, device="cuda") * 0.1
    P32, m32, s32 = ops.th_mix_fwd(S32, wl, bl, ww, bw, 0.144, B, H, N, Np)
    P16, m16, s16 = ops.th_mix_fwd(S16, wl, bl, ww, bw, 0.144, B, H, N, Np)
    assert nerr(P16, P32) <= 2e-3 and nerr(m16, m32) <= 1e-5 and nerr(s16, s32) <= 1e-4
    dPm = torch.zeros(B, H, N, Np, device="cuda")
    dPm[..., :N] = torch.randn(B, H, N, N, device="cuda")
    dPm = dPm.bfloat16()
    g32 = [torch.zeros_like(t) for t in (wl, bl, ww, bw)]
    g16 = [torch.zeros_like(t) for t in (wl, bl, ww, bw)]
    dS32 = ops.th_mix_bwd(S32, dPm, m32, s32, wl, bl, ww, bw, 0.144, *g32, B, H, N, Np)
    dS16 = ops.th_mix_bwd(S16, dPm, m16, s16, wl, bl, ww, bw, 0.144, *g16, B, H, N, Np)
    assert nerr(dS16, dS32) <= 5e-3
    for a, b in zip(g16, g32):
        assert nerr(a, b) <= 2e-3
