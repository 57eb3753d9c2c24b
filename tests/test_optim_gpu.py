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


@pytest.mark.parametrize("kind,wd", [("adam", 0.0), ("adam", 1e-2), ("adamw", 1e-2), ("adamw", 0.0)])
def test_fused_adam_matches_torch(kind, wd):
    """FusedAdam / FusedAdamW (vitk_adam_multi: device-resident step counter and bias corrections) against
    torch.optim.Adam / AdamW, the reference's 'adam' / 'adamw' entries (utils_network.py:119-126)."""
    from vit_torch_b200 import functional
    from vit_torch_b200.train import FusedAdam, FusedAdamW
    torch.manual_seed(1)
    shapes = [(768, 768), (3072,), (5,), (1, 1, 384), (10, 32), (16385,)]
    ps_a = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    ws = [functional.bf16_weight(p) if p.dim() == 2 else None for p in ps_a]
    a = (FusedAdam if kind == "adam" else FusedAdamW)(ps_a, lr=1e-2, weight_decay=wd)
    b = (torch.optim.Adam if kind == "adam" else torch.optim.AdamW)(ps_b, lr=1e-2, weight_decay=wd)
    for step in range(5):
        for pa, pb in zip(ps_a, ps_b):
            g = torch.randn_like(pa) * (0.1 + step)
            pa.grad, pb.grad = g.clone(), g.clone()
        if step == 3:
            a.param_groups[0]["lr"] = b.param_groups[0]["lr"] = 3e-3
        a.step(); b.step()
        for pa, pb in zip(ps_a, ps_b):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6), (kind, wd, step, (pa - pb).abs().max().item())
            assert torch.allclose(a.state[pa]["exp_avg_sq"], b.state[pb]["exp_avg_sq"], rtol=2e-5, atol=1e-7)
    assert int(a.state[ps_a[0]]["step"].item()) == 5
    for p, w in zip(ps_a, ws):
        if w is not None:
            assert functional.bf16_weight(p) is w and torch.equal(w, p.detach().to(torch.bfloat16))


def test_captured_step_follows_the_lr_schedule():
    """The graph Trainer keeps lr in device memory: changing param_groups[0]['lr'] (what the reference's LambdaLR does,
    utils_network.py:218-225) acts on the replayed step exactly as on eager launches."""
    from vit_torch_b200 import models, train
    torch.manual_seed(0)
    xs = [torch.randn(4, 3, 224, 224, device="cuda") for _ in range(7)]
    ys = [torch.randint(0, 10, (4,), device="cuda") for _ in range(7)]
    losses = {}
    for mode in (False, True):
        torch.manual_seed(1)
        m = models.dino_vits16(pretrained=False).cuda()
        train.reset_parameters_like_zoo(m)
        tr = train.Trainer(m, lr=1e-2, graph=mode)
        out = []
        for i, (x, y) in enumerate(zip(xs, ys)):
            if i == 4:
                tr.opt.param_groups[0]["lr"] = 0.2        # a large jump: a stale lr in the graph would show at once
            out.append(tr.step(x, y).item())
        losses[mode] = out
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (losses[False], losses[True])
    assert abs(losses[True][6] - losses[True][4]) > 1e-3      # the jump did change the trajectory
