"""The oracle restatement against the committed golden vectors (generated from the reference's own models/cait.py by
tests/golden/make_golden.py). Runs anywhere (CPU); tolerance covers the fp16 packing of the stored results."""
from functools import partial

import pytest
import torch
import torch.nn as nn

from golden_util import load, nerr, unpack

TOL = 2e-3


def test_tiny_cait_model_vs_golden():
    from oracle import cait as ocait
    g = load("cait_xxs_tiny.pt")
    c = g["cfg"]
    m = ocait.CaiT(img_size=c["img_size"], patch_size=c["patch_size"], embed_dim=c["embed_dim"], depth=c["depth"],
                   num_heads=c["num_heads"], mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                   init_scale=c["init_scale"], depth_token_only=2, num_classes=c["num_classes"])
    m.load_state_dict({k: v.float() for k, v in g["state_dict"].items()})
    out = m(g["x"].float())
    assert nerr(out, g["out"]) <= 1e-5
    loss = torch.nn.functional.cross_entropy(out, g["y"])
    assert abs(loss.item() - g["loss"].item()) <= 1e-5
    loss.backward()
    for k, p in m.named_parameters():
        assert nerr(p.grad, unpack(g["grads"][k])) <= TOL, k


@pytest.mark.parametrize("name", ["talking_head", "class_attention", "layerscale_block", "layerscale_block_ca"])
def test_cait_modules_vs_golden(name):
    from oracle import cait as ocait
    g = load("cait_blocks.pt")[name]
    dim, heads = 96, 2
    norm = partial(nn.LayerNorm, eps=1e-6)
    mod = {"talking_head": lambda: ocait.TalkingHeadAttention(dim, num_heads=heads, qkv_bias=True),
           "class_attention": lambda: ocait.ClassAttention(dim, num_heads=heads, qkv_bias=True),
           "layerscale_block": lambda: ocait.LayerScaleBlock(dim, heads, qkv_bias=True, norm_layer=norm),
           "layerscale_block_ca": lambda: ocait.LayerScaleBlockCA(dim, heads, qkv_bias=True, norm_layer=norm)}[name]()
    mod.load_state_dict({k: v.float() for k, v in g["state_dict"].items()})
    x = g["x"].float().requires_grad_(True)
    if g["cls"] is not None:
        cls = g["cls"].float().requires_grad_(True)
        out = mod(x, cls)
    else:
        cls = None
        out = mod(x)
    assert nerr(out, unpack(g["out"])) <= TOL
    out.backward(g["gout"].float())
    assert nerr(x.grad, unpack(g["gx"])) <= TOL
    if cls is not None:
        assert nerr(cls.grad, unpack(g["gcls"])) <= TOL
    for k, p in mod.named_parameters():
        ref = unpack(g["grads"][k])
        if k == "attn.proj_l.bias" or k == "proj_l.bias":
            assert (p.grad - ref).abs().max().item() <= 1e-5   # analytically zero
        else:
            assert nerr(p.grad, ref) <= TOL, k
