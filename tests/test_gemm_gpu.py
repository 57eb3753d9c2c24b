"""tcgen05 GEMM parity (through the C ABI) against torch fp32 matmul on bf16-rounded inputs.

Tolerances: <= 1e-3 normalised max error vs the bf16-rounded-input fp32 oracle for fp32 outputs,
<= 1e-2 for bf16 outputs (one bf16 rounding of the result: 2^-8 relative).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _mk(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g)


SHAPES = [
    (128, 256, 64), (128, 128, 64), (256, 256, 128), (197 * 8, 768, 768), (1576, 384, 384), (1576, 1152, 384),
    (25216, 2304, 768), (200, 264, 72), (64, 8, 8), (1000, 3072, 768), (1000, 768, 3072), (130, 48, 200),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_nt_store_f32(M, N, K):
    from vit_torch_b200 import ops
    a = _mk((M, K + (-K) % 8), 1)[:, :K].to(torch.bfloat16) if K % 8 else _mk((M, K), 1).to(torch.bfloat16)
    if K % 8:
        pytest.skip("K-major operands need K % 8 == 0 (16-byte TMA pitch)")
    b = _mk((N, K), 2).to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, b, epilogue=ops.EPI_STORE_F32, out=out)
    ref = a.float() @ b.float().t()
    e = nerr(out, ref)
    print(f"NT f32 {M}x{N}x{K}: nerr={e:.3e}")
    assert e <= 1e-3


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1576, 384, 1536), (25216, 768, 3072), (333, 768, 2304)])
def test_gemm_dgrad_layout(M, N, K):
    """A K-major, B MN-major: dX[M, N] = dY[M, K] @ W[K, N] with W stored row-major [K, N]."""
    from vit_torch_b200 import ops
    a = _mk((M, K), 3).to(torch.bfloat16)
    w = _mk((K, N), 4).to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, w, b_mn=True, epilogue=ops.EPI_STORE_F32, out=out)
    ref = a.float() @ w.float()
    e = nerr(out, ref)
    print(f"dgrad {M}x{N}x{K}: nerr={e:.3e}")
    assert e <= 1e-3


@pytest.mark.parametrize("rows,Nout,Kin,splits", [(256, 256, 128, 1), (1576, 384, 384, 0), (25216, 768, 768, 0),
                                                  (25216, 3072, 768, 0), (1000, 1152, 384, 3), (777, 128, 64, 0)])
def test_gemm_wgrad_layout(rows, Nout, Kin, splits):
    """A and B MN-major, split-K atomic accumulate: dW[Nout, Kin] += dY[rows, Nout]^T @ X[rows, Kin]."""
    from vit_torch_b200 import ops
    dy = _mk((rows, Nout), 5).to(torch.bfloat16)
    x = _mk((rows, Kin), 6).to(torch.bfloat16)
    init = _mk((Nout, Kin), 7)
    out = init.clone()
    ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=out, splits=splits)
    ref = init + dy.float().t() @ x.float()
    e = nerr(out, ref)
    print(f"wgrad {rows}: {Nout}x{Kin} splits={splits}: nerr={e:.3e}")
    assert e <= 1e-3


def test_gemm_epilogues():
    from vit_torch_b200 import ops
    M, N, K = 1576, 768, 384
    a = _mk((M, K), 8).to(torch.bfloat16)
    b = (_mk((N, K), 9) * 0.05).to(torch.bfloat16)
    bias = _mk((N,), 10)
    gamma = _mk((N,), 11)
    resid = _mk((M, N), 12)
    acc = a.float() @ b.float().t()

    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, epilogue=ops.EPI_STORE_BF16, bias=bias, out=out)
    assert nerr(out, acc + bias) <= 1e-2

    pre = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    act = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, epilogue=ops.EPI_BIAS_GELU, bias=bias, out=pre, out2=act)
    xp = (acc + bias).clone().requires_grad_(True)           # out = gelu'(pre): what EPI_DGELU multiplies by
    torch.nn.functional.gelu(xp).sum().backward()
    assert nerr(pre, xp.grad) <= 4e-3                         # bf16 rounding of values <= 1.13
    assert nerr(act, torch.nn.functional.gelu(acc + bias)) <= 4e-3
    # exact-erf GELU in fp32 before the bf16 store: error must be pure bf16 rounding, element by element
    ref = torch.nn.functional.gelu(acc + bias)
    assert ((act.float() - ref).abs() <= ref.abs() * 2 ** -8 + 5e-5).all()

    o32 = torch.empty((M, N), device="cuda")
    br = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, epilogue=ops.EPI_RESID_F32, bias=bias, gamma=gamma, resid=resid, out=o32, out2=br)
    assert nerr(o32, resid + gamma * (acc + bias)) <= 1e-3
    assert nerr(br, acc + bias) <= 1e-2
    o32b = torch.empty((M, N), device="cuda")
    ops.gemm(a, b, epilogue=ops.EPI_RESID_F32, bias=bias, resid=resid, out=o32b)
    assert nerr(o32b, resid + acc + bias) <= 1e-3

    # dgelu: out = (dY @ W) * aux, aux = gelu'(pre) saved by the forward epilogue
    w = (_mk((K, N), 13) * 0.05).to(torch.bfloat16)   # [K_red, N] -> B mn-major
    dg = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, w, b_mn=True, epilogue=ops.EPI_DGELU, aux=pre, out=dg)
    assert nerr(dg, (a.float() @ w.float()) * xp.grad) <= 1e-2


def test_gemm_strided_views():
    """Operands and outputs that are column slices of wider buffers (explicit leading dimensions)."""
    from vit_torch_b200 import ops
    M, D = 394, 384
    qkv = _mk((M, 3 * D), 15).to(torch.bfloat16)
    w = _mk((256, D), 16).to(torch.bfloat16)
    outbuf = torch.zeros((M, 512), device="cuda")
    ops.gemm(qkv[:, D:2 * D], w, epilogue=ops.EPI_STORE_F32, out=outbuf[:, 256:])
    assert nerr(outbuf[:, 256:], qkv[:, D:2 * D].float() @ w.float().t()) <= 1e-3
    assert outbuf[:, :256].abs().max().item() == 0.0


def test_gemm_bad_args_fail_loudly():
    from vit_torch_b200 import ops, _lib
    a = torch.zeros((16, 16), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros((10, 16), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.VitkError):   # N = 10 needs 12 output columns, the output pitch has only 10
        ops.gemm(a, b, epilogue=ops.EPI_STORE_F32, out=torch.zeros((16, 10), device="cuda"))
    with pytest.raises(_lib.VitkError):   # operand pitch must be a multiple of 8 elements (16-byte TMA stride)
        ops.gemm(torch.zeros((16, 12), dtype=torch.bfloat16, device="cuda"),
                 torch.zeros((16, 12), dtype=torch.bfloat16, device="cuda"), epilogue=ops.EPI_STORE_F32,
                 out=torch.zeros((16, 16), device="cuda"))


@pytest.mark.parametrize("B,N,H,d", [(2, 196, 8, 48), (3, 197, 2, 64), (1, 577, 4, 48)])
def test_gemm_batched_attention_products(B, N, H, d):
    """Batched mode over (head, image), operands read in place from the token-major qkv buffer: the three products the
    CaiT talking-heads attention needs (scores, P.V, P^T.dO)."""
    from vit_torch_b200 import ops
    D = H * d
    Np = (N + 7) // 8 * 8
    qkv = _mk((B * N, 3 * D), 21).to(torch.bfloat16)
    q, k, v = qkv.float().reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    # S[b,h,i,j] = q.k  (fp32 out, row pitch Np)
    S = torch.zeros((B, H, N, Np), device="cuda")
    ops.gemm_batched(qkv, 3 * D, d, N * 3 * D, False, qkv, 3 * D, d, N * 3 * D, False, N, N, d, H, B, S, Np, N * Np,
                     H * N * Np, a_off=0, b_off=D, out_f32=True)
    ref = q @ k.transpose(-2, -1)
    assert nerr(S[..., :N], ref) <= 1e-3
    assert S[..., N:].abs().max().item() == 0.0 if Np > N else True
    # O[b,i,h,:] = P[b,h,i,:] . v[b,:,h,:]
    P = torch.zeros((B, H, N, Np), device="cuda", dtype=torch.bfloat16)
    P[..., :N] = torch.softmax(ref * d ** -0.5, -1).to(torch.bfloat16)
    O = torch.zeros((B * N, D), device="cuda", dtype=torch.bfloat16)
    ops.gemm_batched(P, Np, N * Np, H * N * Np, False, qkv, 3 * D, d, N * 3 * D, True, N, d, N, H, B, O, D, d, N * D,
                     b_off=2 * D)
    ref_o = (P[..., :N].float() @ v).transpose(1, 2).reshape(B * N, D)
    assert nerr(O, ref_o) <= 1e-2
    # dV[b,j,h,:] = sum_i P[b,h,i,j] dO[b,i,h,:]   (both operands MN-major), written into the V third of a dqkv buffer
    dO = _mk((B * N, D), 22).to(torch.bfloat16)
    dqkv = torch.zeros((B * N, 3 * D), device="cuda", dtype=torch.bfloat16)
    ops.gemm_batched(P, Np, N * Np, H * N * Np, True, dO, D, d, N * D, True, N, d, N, H, B, dqkv, 3 * D, d, N * 3 * D,
                     out_off=2 * D)
    ref_dv = (P[..., :N].float().transpose(-2, -1) @ dO.float().reshape(B, N, H, d).permute(0, 2, 1, 3))
    got = dqkv.float().reshape(B, N, 3, H, d)[:, :, 2].permute(0, 2, 1, 3)
    assert nerr(got, ref_dv) <= 1e-2
    assert dqkv[:, :2 * D].abs().max().item() == 0.0
