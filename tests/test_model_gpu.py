"""Whole-model parity of the fused sm_100a path against the fp32 oracle restatement (same weights, same inputs).

Tolerance (north_star): bf16 path, normalised max error <= 2e-2 per tensor on logits and every gradient, cosine
similarity >= 0.999.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def cosine(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _reset_like_zoo(m):
    """models/vision_all.py:322-329: recursive reset_parameters() (PyTorch default init) for non-pretrained DINO."""
    for c in m.children():
        _reset_like_zoo(c)
    if hasattr(m, "reset_parameters"):
        m.reset_parameters()


def _compare(ours, ref, x, labels, tol=2e-2, tol_grad=2e-2):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ours.load_state_dict(ref.state_dict())
    out_r = ref(x)
    loss_r = torch.nn.functional.cross_entropy(out_r, labels)
    loss_r.backward()
    out_o = ours(x)
    loss_o = torch.nn.functional.cross_entropy(out_o, labels)
    loss_o.backward()
    e = nerr(out_o, out_r)
    print(f"logits nerr={e:.3e} cos={cosine(out_o, out_r):.6f} loss {loss_o.item():.5f} vs {loss_r.item():.5f}")
    assert e <= tol
    assert abs(loss_o.item() - loss_r.item()) <= 2e-2 * abs(loss_r.item())
    worst = (0.0, None)
    gr = dict(ref.named_parameters())
    for name, p in ours.named_parameters():
        r = gr[name]
        if r.grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, name
            continue
        assert p.grad is not None, f"missing grad for {name}"
        ge, gc = nerr(p.grad, r.grad), cosine(p.grad, r.grad)
        if ge > worst[0]:
            worst = (ge, name)
        assert ge <= tol_grad, f"{name}: nerr {ge:.3e}"
        assert gc >= 0.999, f"{name}: cos {gc:.6f}"
    print(f"worst grad nerr={worst[0]:.3e} ({worst[1]})")


@pytest.mark.parametrize("arch,B,size", [("dino_vits16", 4, 224), ("dino_vits16", 3, 96), ("dino_vits8", 2, 96),
                                         ("dino_vitb16", 2, 224), ("dino_vitb8", 2, 224)])   # last: BASELINE config 3, N = 785
def test_dino_matches_oracle(arch, B, size):
    from oracle import vit as ovit
    from vit_torch_b200 import models
    torch.manual_seed(0)
    ref = getattr(ovit, arch)(pretrained=False).cuda()
    _reset_like_zoo(ref)
    ours = getattr(models, arch)(pretrained=False).cuda()
    x = torch.randn(B, 3, size, size, device="cuda")
    labels = torch.randint(0, 10, (B,), device="cuda")
    _compare(ours, ref, x, labels)


@pytest.mark.parametrize("distilled", [False, True])
def test_deit_matches_oracle(distilled):
    from oracle import vit as ovit
    from vit_torch_b200 import models
    torch.manual_seed(1)
    ref = ovit.TimmVisionTransformer(embed_dim=384, depth=12, num_heads=6, num_classes=10, distilled=distilled).cuda()
    _reset_like_zoo(ref)
    ours = models.VisionTransformer(embed_dim=384, depth=12, num_heads=6, num_classes=10, distilled=distilled).cuda()
    x = torch.randn(3, 3, 224, 224, device="cuda")
    labels = torch.randint(0, 10, (3,), device="cuda")
    _compare(ours, ref, x, labels)


def test_lineareval_no_grad_forward():
    """--lineareval: backbone forward under no_grad (utils_network.py:413-415) must match the training-mode forward."""
    from vit_torch_b200 import models
    torch.manual_seed(2)
    m = models.dino_vits16(pretrained=False).cuda()
    x = torch.randn(2, 3, 224, 224, device="cuda")
    with torch.no_grad():
        a = m(x)
    b = m(x)
    assert torch.equal(a, b.detach())
