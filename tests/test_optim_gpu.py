"""FusedSGD (one multi-tensor kernel + bf16 weight refresh) against torch.optim.SGD, the reference's 'sgd' optimiser
(utils_network.py:119-126): agreement to fp32 rounding (the kernel contracts mul+add into fma)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_sgd_matches_torch_sgd():
    from vit_torch_b200 import functional
    from vit_torch_b200.train import FusedSGD
    torch.manual_seed(0)
    shapes = [(768, 768), (3072,), (5,), (1, 1, 384), (2304, 768), (10, 32), (16385,)]
    ps_a = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    ws = [functional.bf16_weight(p) if p.dim() == 2 else None for p in ps_a]
    a = FusedSGD(ps_a, lr=1e-2, momentum=0.9)
    b = torch.optim.SGD(ps_b, lr=1e-2, momentum=0.9)
    for step in range(4):
        for pa, pb in zip(ps_a, ps_b):
            g = torch.randn_like(pa)
            pa.grad = g.clone() if not (step == 2 and pa.numel() == 5) else None   # a parameter without grad is skipped
            pb.grad = g.clone() if pa.grad is not None else None
        if step == 3:
            a.param_groups[0]["lr"] = b.param_groups[0]["lr"] = 5e-3                # LambdaLR turns this knob
        a.step(); b.step()
        for pa, pb in zip(ps_a, ps_b):
            assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-6)
            assert torch.allclose(a.state[pa]["momentum_buffer"], b.state[pb]["momentum_buffer"], rtol=1e-6, atol=2e-6)   # fma vs mul+add near cancellation
    for p, w in zip(ps_a, ws):
        if w is not None:
            assert functional.bf16_weight(p) is w                                   # cache entry still live ...
            assert torch.equal(w, p.detach().to(torch.bfloat16))                    # ... and refreshed by the step


def test_fused_sgd_rejects_other_configurations():
    from vit_torch_b200.train import FusedSGD
    p = [torch.nn.Parameter(torch.zeros(4, device="cuda"))]
    with pytest.raises(NotImplementedError):
        FusedSGD(p, lr=0.1, momentum=0.9, weight_decay=1e-4)
    with pytest.raises(NotImplementedError):
        FusedSGD(p, lr=0.1, momentum=0.9, nesterov=True)
