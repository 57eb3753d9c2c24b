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
