"""Fused classifier head (models/vision_all.py:299-320) vs the torch fp32 Sequential with the same weights."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,fin,units", [(128, 384, [256, 128, 32, 10]), (8, 768, [10]), (512, 768, [256, 10]), (5, 384, [64, 100])])
def test_head_matches_torch(B, fin, units):
    from vit_torch_b200 import zoo
    torch.manual_seed(0)
    head = zoo.get_classifier_head(fin, units).cuda()
    ref = nn.Sequential(*[nn.Linear(m.in_features, m.out_features, bias=m.bias is not None) if isinstance(m, nn.Linear)
                          else nn.GELU() for m in head]).cuda()
    ref.load_state_dict(head.state_dict())
    assert list(head.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(B, fin, device="cuda")
    xo = x.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = torch.randint(0, units[-1], (B,), device="cuda")
    lo = torch.nn.functional.cross_entropy(head(xo), y)
    lr = torch.nn.functional.cross_entropy(ref(xr), y)
    assert abs(lo.item() - lr.item()) <= 2e-2 * abs(lr.item())
    lo.backward()
    lr.backward()
    assert nerr(xo.grad, xr.grad) <= 2e-2
    for (k, p), (_, q) in zip(head.named_parameters(), ref.named_parameters()):
        assert nerr(p.grad, q.grad) <= 2e-2, k


@pytest.mark.parametrize("B,C,pitch", [(128, 10, 16), (512, 10, 10), (7, 1000, 1000), (1, 3, 8), (33, 37, 40)])
def test_fused_cross_entropy_matches_torch(B, C, pitch):
    """vitk_cross_entropy (loss + gradient + argmax count in one launch) vs nn.CrossEntropyLoss + autograd and
    classification_count_correct (utils_network.py:85-95, 429-433). fp32 glue tolerance 1e-4."""
    from vit_torch_b200 import functional
    g = torch.Generator(device="cuda").manual_seed(B * 31 + C)
    buf = torch.randn((B, pitch), device="cuda", generator=g) * 3
    logits = buf[:, :C].detach().requires_grad_(True)           # strided rows, like the padded head output
    labels = torch.randint(0, C, (B,), device="cuda", generator=g)
    if B > 2:
        logits.data[1, :] = 0.5                                  # ties: argmax must pick the first maximum
    loss, ncorrect = functional.cross_entropy(logits, labels)
    (loss * 2.0).backward()
    ref_in = buf[:, :C].detach().clone().requires_grad_(True)
    if B > 2:
        ref_in.data[1, :] = 0.5
    ref = torch.nn.functional.cross_entropy(ref_in, labels)
    (ref * 2.0).backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item()))
    assert (logits.grad - ref_in.grad).abs().max().item() <= 1e-6 + 1e-4 * ref_in.grad.abs().max().item()
    assert int(ncorrect.item()) == int((ref_in.argmax(1) == labels).sum().item())
    assert not ncorrect.requires_grad
