"""Regression tests for defects found by review of round 1: the graph Trainer's pointer table under many
odd-shape eager steps, DropPath + LayerScale gradients with independent per-branch draws, optimiser state_dict reload,
ignore_index / bad labels in the fused cross-entropy."""
from functools import partial

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def test_graph_replay_survives_many_odd_shape_eager_steps():
    """The captured step uploads its pointer table from a capture-private pinned buffer: >= 9 eager steps (another
    batch shape: the reference loader has drop_last=False, utils_datasets.py:890) between replays must not change what
    the replays apply. Compared with the all-eager trainer on the same data."""
    from vit_torch_b200 import models, train
    torch.manual_seed(0)
    sizes = [4, 4, 4, 4] + [2] * 10 + [4, 4, 4]
    xs = [torch.randn(b, 3, 96, 96, device="cuda") for b in sizes]
    ys = [torch.randint(0, 10, (b,), device="cuda") for b in sizes]
    losses, finals = {}, {}
    for mode in (False, True):
        torch.manual_seed(1)
        m = models.dino_vits16(pretrained=False).cuda()
        train.reset_parameters_like_zoo(m)
        tr = train.Trainer(m, lr=5e-2, graph=mode, strict_graph=True)
        losses[mode] = [tr.step(x, y).item() for x, y in zip(xs, ys)]
        finals[mode] = {k: v.detach().clone() for k, v in m.state_dict().items()}
        if mode:
            assert tr.static_inputs() is not None
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 5e-3 * max(1.0, abs(a)), (losses[False], losses[True])
    # parameters outside the gradient arena (their gradient pointers travel through the table): cls/pos/patch/norm
    for k in ("cls_token", "pos_embed", "patch_embed.proj.weight", "norm.weight", "norm.bias"):
        assert nerr(finals[True][k], finals[False][k]) <= 2e-3, k


class _Masks:
    """Supplies the same DropPath masks to the fused modules and to the oracle, in call order."""

    def __init__(self, masks):
        self.masks, self.i = masks, 0

    def next(self):
        m = self.masks[self.i % len(self.masks)]
        self.i += 1
        return m


@pytest.mark.parametrize("layerscale", [False, True])
def test_droppath_block_two_independent_draws(monkeypatch, layerscale):
    """Block / LayerScale_Block with drop_path > 0 (timm Block; models/cait.py:140,148-149): one independent mask per
    residual branch, scale 1/keep; every gradient incl. gamma_1 / gamma_2 against an fp32 torch restatement fed the
    same masks."""
    from oracle import vit as ovit
    from vit_torch_b200 import modules
    torch.manual_seed(0)
    B, N, D, H, keep = 6, 37, 128, 2, 0.7
    masks = [torch.tensor([1, 0, 1, 1, 0, 1.], device="cuda") / keep, torch.tensor([0, 1, 1, 0, 1, 1.], device="cuda") / keep]
    ours = modules.Block(D, H, qkv_bias=True, drop_path=1 - keep, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                         init_values=0.5 if layerscale else None).cuda().train()
    ref = ovit.Block(D, H, qkv_bias=True, drop_path=1 - keep, norm_layer=partial(nn.LayerNorm, eps=1e-6)).cuda().train()
    with torch.no_grad():
        for p in ours.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.05)
    ref.load_state_dict({k: v for k, v in ours.state_dict().items() if not k.startswith("gamma")})
    if layerscale:
        with torch.no_grad():
            ours.gamma_1.uniform_(0.2, 1.0)
            ours.gamma_2.uniform_(0.2, 1.0)
    g1 = ours.gamma_1.detach().clone().requires_grad_(True) if layerscale else None
    g2 = ours.gamma_2.detach().clone().requires_grad_(True) if layerscale else None

    mo = _Masks(masks)
    monkeypatch.setattr(modules.DropPath, "rowscale", lambda self, batch, device, training: mo.next())
    x = torch.randn(B, N, D, device="cuda")
    xo = x.clone().requires_grad_(True)
    out = ours(xo)
    gout = torch.randn_like(out)
    out.backward(gout)
    assert mo.i == 2

    xr = x.clone().requires_grad_(True)
    a = ref.attn(ref.norm1(xr))
    x1 = xr + masks[0].view(B, 1, 1) * (a * g1 if layerscale else a)
    f = ref.mlp(ref.norm2(x1))
    x2 = x1 + masks[1].view(B, 1, 1) * (f * g2 if layerscale else f)
    x2.backward(gout)
    assert nerr(out, x2) <= 2e-2
    assert nerr(xo.grad, xr.grad) <= 2e-2
    for k, p in ours.named_parameters():
        if k.startswith("gamma"):
            r = (g1 if k == "gamma_1" else g2).grad
        else:
            r = dict(ref.named_parameters())[k].grad
        assert nerr(p.grad, r) <= 2e-2, (k, nerr(p.grad, r))


def test_fused_optimizer_reload_state_dict():
    """load_state_dict() swaps the state tensors: the cached launch plan must not keep the old pointers, and a resumed
    Adam continues its bias corrections from the loaded step count."""
    from vit_torch_b200.train import FusedAdam, FusedSGD
    torch.manual_seed(0)
    for cls, ref_cls, kw in ((FusedSGD, torch.optim.SGD, dict(lr=1e-2, momentum=0.9)),
                             (FusedAdam, torch.optim.Adam, dict(lr=1e-2))):
        ps_a = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in [(64, 32), (7,), (16385,)]]
        ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
        a, b = cls(ps_a, **kw), ref_cls(ps_b, **kw)
        grads = [[torch.randn_like(p) for p in ps_a] for _ in range(6)]
        for step in range(3):
            for pa, pb, g in zip(ps_a, ps_b, grads[step]):
                pa.grad, pb.grad = g.clone(), g.clone()
            a.step(); b.step()
        # "resume": fresh optimiser objects, state restored from the state_dicts
        a2, b2 = cls(ps_a, **kw), ref_cls(ps_b, **kw)
        a2.load_state_dict(a.state_dict())
        b2.load_state_dict(b.state_dict())
        a2.load_state_dict(a2.state_dict())      # reloading into an optimiser that already holds state is a no-op
        for step in range(3, 6):
            for pa, pb, g in zip(ps_a, ps_b, grads[step]):
                pa.grad, pb.grad = g.clone(), g.clone()
            a2.step(); b2.step()
            for pa, pb in zip(ps_a, ps_b):
                assert torch.allclose(pa, pb, rtol=3e-5, atol=3e-6), (cls.__name__, step, (pa - pb).abs().max().item())
        if cls is FusedAdam:
            assert int(a2.state[ps_a[0]]["step"].item()) == 6


def test_cross_entropy_ignore_index_and_bad_label():
    from vit_torch_b200 import functional
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = (torch.randn((9, 12), device="cuda", generator=g) * 2).requires_grad_(True)
    labels = torch.randint(0, 12, (9,), device="cuda", generator=g)
    labels[2] = -100
    labels[7] = -100
    loss, ncorrect = functional.cross_entropy(logits, labels)
    loss.backward()
    ref_in = logits.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, labels)       # default ignore_index = -100
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item()))
    assert (logits.grad - ref_in.grad).abs().max().item() <= 1e-6
    assert torch.all(logits.grad[2] == 0) and torch.all(logits.grad[7] == 0)
    assert int(ncorrect.item()) == int((ref_in.argmax(1) == labels).sum().item())
    bad = labels.clone()
    bad[0] = 12                                                   # out of range and not the ignore value: loud NaN
    loss_bad, _ = functional.cross_entropy(logits.detach(), bad)
    assert torch.isnan(loss_bad).item()
