"""Memory-bound glue kernels (LayerNorm fwd/bwd, column sums, casts) vs torch fp32 (fp32-I/O tolerance 1e-4)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("rows,D", [(1576, 384), (25216, 768), (333, 192), (77, 288), (9, 1024), (1, 768)])
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
