"""CaiT on the fused path: against the committed golden vectors (generated from the reference's own models/cait.py)
and against the fp32 oracle at full model size. bf16 tolerance 2e-2 normalised max error, cosine >= 0.999."""
from functools import partial

import pytest
import torch
import torch.nn as nn

from golden_util import load, nerr, unpack

pytestmark = pytest.mark.gpu
TOL = 2e-2


def cosine(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def test_layerscale_block_vs_golden():
    from vit_torch_b200 import cait
    g = load("cait_blocks.pt")["layerscale_block"]
    mod = cait.LayerScale_Block(96, 2, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6)).cuda()
    mod.load_state_dict({k: v.float() for k, v in g["state_dict"].items()})
    x = g["x"].float().cuda().requires_grad_(True)
    out = mod(x)
    ref = unpack(g["out"]).cuda()
    print("block out nerr", nerr(out, ref))
    assert nerr(out, ref) <= TOL
    out.backward(g["gout"].float().cuda())
    assert nerr(x.grad, unpack(g["gx"]).cuda()) <= TOL
    for k, p in mod.named_parameters():
        r = unpack(g["grads"][k]).cuda()
        if k == "attn.proj_l.bias":
            assert (p.grad - r).abs().max().item() <= 1e-3 * max(1.0, unpack(g["grads"]["attn.proj_l.weight"]).abs().max().item())
            continue
        e = nerr(p.grad, r)
        print(f"  {k}: nerr {e:.3e}")
        assert e <= TOL, k


def test_layerscale_block_ca_vs_golden():
    from vit_torch_b200 import cait
    g = load("cait_blocks.pt")["layerscale_block_ca"]
    mod = cait.LayerScale_Block_CA(96, 2, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6)).cuda()
    mod.load_state_dict({k: v.float() for k, v in g["state_dict"].items()})
    x = g["x"].float().cuda().requires_grad_(True)
    cls = g["cls"].float().cuda().requires_grad_(True)
    out = mod(x, cls)
    ref = unpack(g["out"]).cuda()
    print("CA block out nerr", nerr(out, ref))
    assert nerr(out, ref) <= TOL
    out.backward(g["gout"].float().cuda())
    assert nerr(x.grad, unpack(g["gx"]).cuda()) <= TOL
    assert nerr(cls.grad, unpack(g["gcls"]).cuda()) <= TOL
    for k, p in mod.named_parameters():
        r = unpack(g["grads"][k]).cuda()
        if k == "attn.k.bias":  # softmax-invariant (adds the same q.b to every key): analytically zero gradient
            assert p.grad.abs().max().item() <= 1e-3 * unpack(g["grads"]["attn.k.weight"]).abs().max().item()
            continue
        e = nerr(p.grad, r)
        print(f"  {k}: nerr {e:.3e}")
        assert e <= TOL, k


def test_tiny_cait_model_vs_golden():
    from vit_torch_b200 import cait
    g = load("cait_xxs_tiny.pt")
    c = g["cfg"]
    m = cait.cait_models(img_size=c["img_size"], patch_size=c["patch_size"], embed_dim=c["embed_dim"], depth=c["depth"],
                         num_heads=c["num_heads"], mlp_ratio=4, qkv_bias=True,
                         norm_layer=partial(nn.LayerNorm, eps=1e-6), init_scale=c["init_scale"], depth_token_only=2,
                         num_classes=c["num_classes"]).cuda()
    m.load_state_dict({k: v.float() for k, v in g["state_dict"].items()})
    out = m(g["x"].float().cuda())
    e = nerr(out, g["out"].cuda())
    print("tiny cait logits nerr", e)
    assert e <= TOL
    loss = torch.nn.functional.cross_entropy(out, g["y"].cuda())
    assert abs(loss.item() - g["loss"].item()) <= 2e-2 * abs(g["loss"].item())
    loss.backward()
    worst = (0.0, None)
    for k, p in m.named_parameters():
        r = unpack(g["grads"][k]).cuda()
        if k.endswith("proj_l.bias") or k.endswith("attn.k.bias"):   # analytically zero gradients
            continue
        e = nerr(p.grad, r)
        worst = max(worst, (e, k))
        assert e <= TOL, f"{k}: {e:.3e}"
        assert cosine(p.grad, r) >= 0.999, k
    print("worst grad", worst)


@pytest.mark.parametrize("name,B", [("cait_XXS24_224", 2), ("cait_S24_224", 2)])
def test_cait_matches_oracle(name, B):
    from oracle import cait as ocait
    from vit_torch_b200 import cait
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    ref = ocait.create(name, num_classes=10, drop_rate=0.0, drop_path_rate=0.0).cuda()
    ours = getattr(cait, name)(pretrained=False, num_classes=10, drop_rate=0.0, drop_path_rate=0.0).cuda()
    # LayerScale init 1e-5 makes every block a near no-op; use O(1e-1) gammas so that block errors are visible
    with torch.no_grad():
        for n_, p in ref.named_parameters():
            if "gamma" in n_:
                p.fill_(0.1)
            if n_.endswith("proj_l.weight") or n_.endswith("proj_w.weight"):
                p.add_(torch.eye(p.shape[0], device=p.device))
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(B, 3, 224, 224, device="cuda")
    y = torch.randint(0, 10, (B,), device="cuda")
    o_r = ref(x)
    torch.nn.functional.cross_entropy(o_r, y).backward()
    o_o = ours(x)
    torch.nn.functional.cross_entropy(o_o, y).backward()
    e = nerr(o_o, o_r)
    print(f"{name} logits nerr {e:.3e} cos {cosine(o_o, o_r):.6f}")
    assert e <= TOL
    gr = dict(ref.named_parameters())
    worst = (0.0, None)
    for k, p in ours.named_parameters():
        r = gr[k].grad
        if k.endswith("proj_l.bias") or k.endswith("attn.k.bias"):   # analytically zero gradients
            assert p.grad.abs().max().item() <= 1e-3 * max(1e-6, gr[k.replace("bias", "weight")].grad.abs().max().item())
            continue
        ge = nerr(p.grad, r)
        worst = max(worst, (ge, k))
        # every tensor, the H x H talking-heads mixing parameters included, meets the 2e-2 north-star tolerance
        # (measured on B200: worst 1.9e-2 on blocks.3.attn.proj_w.bias, 1.8e-2 on a proj_l.weight)
        assert ge <= TOL, f"{k}: {ge:.3e}"
        assert cosine(p.grad, r) >= 0.999, k
    print("worst grad", worst)
