"""CPU tests of the oracle restatement itself (shapes, parameter counts, first-principles attention)."""
import torch


def test_dino_param_counts_and_shapes():
    from oracle import vit
    m = vit.dino_vits16()
    n = sum(p.numel() for p in m.parameters())
    assert abs(n - 21.67e6) < 0.05e6
    x = torch.randn(2, 3, 224, 224)
    assert m(x).shape == (2, 384)
    assert m(torch.randn(1, 3, 96, 96)).shape == (1, 384)   # interpolate_pos_encoding path
    keys = list(m.state_dict().keys())
    assert "blocks.0.attn.qkv.weight" in keys and "patch_embed.proj.weight" in keys and "cls_token" in keys


def test_attention_matches_first_principles():
    from oracle import vit
    torch.manual_seed(0)
    a = vit.Attention(64, num_heads=2, qkv_bias=True)
    x = torch.randn(2, 5, 64)
    y = a(x)
    qkv = torch.nn.functional.linear(x, a.qkv.weight, a.qkv.bias).reshape(2, 5, 3, 2, 32)
    outs = []
    for h in range(2):
        q, k, v = qkv[:, :, 0, h], qkv[:, :, 1, h], qkv[:, :, 2, h]
        p = torch.softmax(torch.einsum("bid,bjd->bij", q, k) * 32 ** -0.5, -1)
        outs.append(torch.einsum("bij,bjd->bid", p, v))
    ref = torch.nn.functional.linear(torch.cat(outs, -1), a.proj.weight, a.proj.bias)
    assert torch.allclose(y, ref, atol=1e-5)


def test_timm_vit_distilled():
    from oracle import vit
    m = vit.TimmVisionTransformer(embed_dim=192, depth=2, num_heads=3, num_classes=10, distilled=True)
    assert m(torch.randn(2, 3, 224, 224)).shape == (2, 10)
    assert m.pos_embed.shape == (1, 198, 192)
