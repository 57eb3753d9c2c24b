"""Memory-bound glue kernels (LayerNorm fwd/bwd, column sums, casts) vs torch fp32 (fp32-I/O tolerance 1e-4)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("rows,D", [(1576, 384), (25216, 768), (333, 192), (77, 288), (9, 1024), (1, 768),
                                     (4100, 384), (5001, 192), (4096, 1024), (12345, 288)])   # >= 4096 rows: TMA-staged ring kernel
def test_layernorm_fwd_bwd(rows, D):
    from vit_torch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((rows, D), device="cuda", generator=g) * 2 + 0.5
    w = torch.randn((D,), device="cuda", generator=g)
    b = torch.randn((D,), device="cuda", generator=g)
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-6)
    assert nerr(y, yr) <= 1e-2            # bf16 output rounding
    assert nerr(mean, x.mean(-1)) <= 1e-4
    assert nerr(rstd, (x.var(-1, unbiased=False) + 1e-6).rsqrt()) <= 1e-4

    dy = torch.randn((rows, D), device="cuda", generator=g).to(torch.bfloat16)
    dres = torch.randn((rows, D), device="cuda", generator=g)
    cs = torch.randn((D,), device="cuda", generator=g)
    dw = torch.zeros((D,), device="cuda")
    db = torch.zeros((D,), device="cuda")
    dx, dxb = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, dweight=dw, dbias=db, want_bf16=True, colscale=cs)
    yr.backward(dy.float())
    ref_dx = xr.grad + dres
    assert nerr(dx, ref_dx) <= 1e-4
    assert nerr(dxb, ref_dx * cs) <= 1e-2
    assert nerr(dw, wr.grad) <= 1e-4
    assert nerr(db, br.grad) <= 1e-4


@pytest.mark.parametrize("rows,D", [(6000, 768), (300, 768)])
def test_layernorm_bwd_variants(rows, D):
    """No residual gradient, fused column sums of the bf16 copy, fp32 dy with strided rows (both LN-backward kernels)."""
    from vit_torch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((rows, D), device="cuda", generator=g)
    w = torch.randn((D,), device="cuda", generator=g)
    b = torch.zeros((D,), device="cuda")
    _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), w, b, 1e-6)
    dy = torch.randn((rows, D), device="cuda", generator=g)
    yr.backward(dy.to(torch.bfloat16).float())
    dsum = torch.zeros((D,), device="cuda")
    dx, dxb = ops.layernorm_bwd(dy.to(torch.bfloat16), x, w, mean, rstd, want_bf16=True, dxsum=dsum)
    assert nerr(dx, xr.grad) <= 1e-4
    assert nerr(dxb, xr.grad) <= 1e-2
    assert nerr(dsum, dx.sum(0)) <= 1e-4                      # column sums of the values the bf16 copy is rounded from
    # fp32 dy, rows of x / dx embedded in a wider buffer (the final-norm-on-cls-token call pattern)
    wide = torch.zeros((rows, 2 * D), device="cuda")
    wide[:, :D] = x
    dxw = torch.zeros((rows, 2 * D), device="cuda")
    xr.grad = None
    yr2 = torch.nn.functional.layer_norm(xr, (D,), w, b, 1e-6)
    yr2.backward(dy)
    ops.layernorm_bwd_rows(dy, wide, 2 * D, rows, D, w, mean, rstd, dx=dxw, dx_stride=2 * D)
    assert nerr(dxw[:, :D], xr.grad) <= 1e-4
    assert dxw[:, D:].abs().max().item() == 0.0


@pytest.mark.parametrize("rows,N", [(25216, 3072), (1576, 384), (5, 8), (1000, 2304)])
def test_colsum(rows, N):
    from vit_torch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((rows, N), device="cuda", generator=g).to(torch.bfloat16)
    out = torch.ones((N,), device="cuda")
    ops.colsum_accum(x, out)
    assert nerr(out, 1 + x.float().sum(0)) <= 1e-4


def test_cast():
    from vit_torch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    for n in (1, 7, 8, 1000003):
        x = torch.randn((n,), device="cuda", generator=g)
        assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
