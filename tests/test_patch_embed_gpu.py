"""im2col-free PatchEmbed (patch_embed.cu: TMA gather from NCHW, tf32 / bf16 tensor cores) against
Conv2d(C, D, P, P)(x).flatten(2).transpose(1, 2) + prefix tokens + pos_embed and its autograd, as the reference
computes them (models/swin.py:434-445 witness, models/cait.py:229-234, models/deit.py:35-43)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def _reference(img, w, b, pos, prefix, P):
    x = F.conv2d(img, w, b, stride=P).flatten(2).transpose(1, 2)
    if prefix is not None:
        x = torch.cat((prefix.expand(img.shape[0], -1, -1), x), dim=1)
    return x + pos


CASES = [  # B, C, size, P, D, T, dtype
    (3, 3, 224, 16, 384, 1, torch.float32),
    (2, 3, 224, 16, 768, 2, torch.float32),
    (2, 3, 96, 16, 192, 1, torch.float32),
    (2, 3, 224, 8, 384, 1, torch.float32),
    (2, 7, 64, 16, 96, 0, torch.float32),
    (3, 3, 224, 16, 384, 1, torch.bfloat16),
    (2, 3, 96, 16, 768, 0, torch.bfloat16),
    (2, 3, 224, 32, 256, 1, torch.bfloat16),
    (2, 3, 224, 8, 384, 1, torch.bfloat16),      # 16-byte patch rows: unswizzled core-matrix layout
    (2, 3, 160, 4, 64, 0, torch.float32),        # P = 4 (Swin-style patch size), 16-byte rows in fp32
]


@pytest.mark.parametrize("B,C,size,P,D,T,dtype", CASES)
def test_patch_embed_matches_conv(B, C, size, P, D, T, dtype):
    from vit_torch_b200 import functional as Fn
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + size + P + D)
    n = (size // P) ** 2
    img = torch.randn((B, C, size, size), device="cuda", generator=g)
    if dtype == torch.bfloat16:
        img = img.to(torch.bfloat16)
    w = (torch.randn((D, C, P, P), device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = torch.randn((D,), device="cuda", generator=g).requires_grad_(True)
    pos = (torch.randn((1, n + T, D), device="cuda", generator=g) * 0.1).requires_grad_(True)
    prefix = (torch.randn((1, T, D), device="cuda", generator=g) * 0.1).requires_grad_(True) if T else None
    out = Fn.TokensFn.apply(img, w, b, pos, prefix, P, None)
    gout = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(gout)
    got = [t.grad.clone() for t in (w, b, pos)] + ([prefix.grad.clone()] if T else [])
    for t in (w, b, pos, prefix):
        if t is not None:
            t.grad = None
    ref = _reference(img.float(), w, b, pos, prefix, P)
    ref.backward(gout)
    want = [t.grad for t in (w, b, pos)] + ([prefix.grad] if T else [])
    tol = 2e-3 if dtype == torch.float32 else 1e-2       # tf32 (10-bit mantissa) / bf16 weights, fp32 accumulation
    assert nerr(out, ref) <= tol, nerr(out, ref)
    tma = Fn.patch_embed_tma_ok(dtype, C, size, size, P)     # (the patch-matrix fallback sums a bf16 copy for db)
    for name, a, r in zip(("dW", "db", "dpos", "dprefix"), got, want):
        lim = 1e-2 if name == "dW" else (5e-3 if (name == "db" and not tma) else 1e-4)     # dW: bf16 operands
        assert nerr(a, r) <= lim, (name, nerr(a, r))


def test_uint8_input_pipeline_matches_totensor_normalize():
    """uint8 images + set_input_normalization == ToTensor() + Normalize(mean, std) on the host followed by the fp32
    model (utils_datasets.py:573-580), within bf16 rounding of the normalised pixels."""
    from vit_torch_b200 import models
    torch.manual_seed(0)
    mean, std = [0.4467, 0.4398, 0.4066], [0.2603, 0.2566, 0.2713]
    m = models.dino_vits16(pretrained=False).cuda()
    x8 = torch.randint(0, 256, (2, 3, 224, 224), dtype=torch.uint8, device="cuda")
    xf = (x8.float() / 255 - torch.tensor(mean, device="cuda").view(1, 3, 1, 1)) / torch.tensor(std, device="cuda").view(1, 3, 1, 1)
    with pytest.raises(ValueError):
        m(x8)
    m.patch_embed.set_input_normalization(mean, std)
    with torch.no_grad():
        a, b = m(x8), m(xf)
    assert nerr(a, b) <= 2e-2
    assert "input_norm" not in "".join(m.state_dict().keys())
