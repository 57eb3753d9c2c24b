"""Pins the CaiT oracle restatement against the reference's OWN models/cait.py (build container only: the reference
tree does not travel to the GPU box). Bit-level agreement in fp32 on CPU for identical weights and inputs."""
import importlib.util
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.reference

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def ref_cait():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden",
                                                                              "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.import_reference_cait()


@pytest.mark.parametrize("name", ["cait_XXS24_224", "cait_S24_224"])
def test_cait_model_bit_exact(ref_cait, name):
    from oracle import cait as ocait
    torch.manual_seed(0)
    ref = getattr(ref_cait, name)(pretrained=False, num_classes=10, drop_rate=0.0, drop_path_rate=0.0)
    ours = ocait.create(name, num_classes=10, drop_rate=0.0, drop_path_rate=0.0)
    assert [k for k, _ in ref.named_parameters()] == [k for k, _ in ours.named_parameters()]
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 224, 224)
    y = torch.tensor([3, 8])
    o_r, o_o = ref(x), ours(x)
    assert torch.equal(o_r, o_o)
    torch.nn.functional.cross_entropy(o_r, y).backward()
    torch.nn.functional.cross_entropy(o_o, y).backward()
    for (k, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
        assert torch.equal(pr.grad, po.grad), k


def test_cait_constructor_table_matches_reference(ref_cait):
    from oracle import cait as ocait
    for name, (img, dim, depth, heads, init) in ocait.CAIT_SIZES.items():
        if name in ("cait_M48", "cait_M36", "cait_S36", "cait_XXS36", "cait_XXS36_224"):
            continue  # large: checked structurally below without instantiating twice
        ref = getattr(ref_cait, name)(pretrained=False)
        assert ref.embed_dim == dim and len(ref.blocks) == depth and ref.blocks[0].attn.num_heads == heads
        assert ref.patch_embed.img_size == (img, img)
        assert abs(ref.blocks[0].gamma_1[0].item() - init) < 1e-12


# ----------------------------------------------------------------------------------------------------------------
# DeiT: the reference's own models/deit.py (token assembly with the distillation token :32-49, the two-head average
# :67-78 / :86-91, the constructors' hyper-parameters :95-211) executed on top of the oracle's timm base class.
# ----------------------------------------------------------------------------------------------------------------
def _import_reference(name, filename):
    """Import /root/reference/models/<filename> with timm resolved to oracle-backed stubs (models/deit.py:7-9,
    models/swin.py:11)."""
    import importlib.util
    import types

    from oracle import vit as ovit
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.VisionTransformer, vt.Mlp, vt.PatchEmbed = ovit.TimmVisionTransformer, ovit.Mlp, ovit.PatchEmbed
    vt._cfg = lambda **kw: dict(kw)
    reg = types.ModuleType("timm.models.registry")
    reg.register_model = lambda fn: fn
    lay = types.ModuleType("timm.models.layers")
    lay.trunc_normal_, lay.DropPath = ovit.trunc_normal_, ovit.DropPath
    lay.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    stubs = {"timm": types.ModuleType("timm"), "timm.models": types.ModuleType("timm.models"),
             "timm.models.vision_transformer": vt, "timm.models.registry": reg, "timm.models.layers": lay}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join("/root/reference/models", filename))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


@pytest.fixture(scope="module")
def ref_deit():
    return _import_reference("reference_deit", "deit.py")


def test_deit_distilled_bit_exact(ref_deit):
    """DeitCustomDistilled (models/deit.py:20-59,82-91): its own forward_features / forward vs the oracle's distilled
    path, same weights -> identical logits and gradients."""
    from oracle import vit as ovit
    from functools import partial
    torch.manual_seed(0)
    kw = dict(patch_size=16, embed_dim=192, depth=2, num_heads=3, mlp_ratio=4, qkv_bias=True, num_classes=10,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    ref = ref_deit.DeitCustomDistilled(**kw)
    ours = ovit.TimmVisionTransformer(distilled=True, **kw)
    assert sorted(k for k, _ in ref.named_parameters()) == sorted(k for k, _ in ours.named_parameters())
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 224, 224)
    y = torch.tensor([1, 7])
    f_r, f_o = ref.forward_features(x), ours.forward_features(x)
    assert torch.equal(f_r[0], f_o[0]) and torch.equal(f_r[1], f_o[1])
    o_r, o_o = ref(x), ours(x)
    assert torch.equal(o_r, o_o)
    torch.nn.functional.cross_entropy(o_r, y).backward()
    torch.nn.functional.cross_entropy(o_o, y).backward()
    gr = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        assert torch.equal(p.grad, gr[k].grad), k
    # the reference's training-mode DistilledVisionTransformer returns both heads (models/deit.py:51-59)
    base = ref_deit.DistilledVisionTransformer(**kw)
    base.load_state_dict(ref.state_dict())
    a, b = base.train()(x)
    assert torch.equal((a + b) / 2, o_r.detach())


def test_deit_custom_and_constructors(ref_deit):
    """DeitCustom.forward (models/deit.py:63-78) on the plain and the distilled base, and the hyper-parameters of the
    registered constructors (models/deit.py:95-211) against the oracle's / the product's tables."""
    from oracle import vit as ovit
    from functools import partial
    torch.manual_seed(1)
    kw = dict(patch_size=16, embed_dim=192, depth=1, num_heads=3, mlp_ratio=4, qkv_bias=True, num_classes=5,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    x = torch.randn(1, 3, 224, 224)
    for distilled in (False, True):
        ref = ref_deit.DeitCustom(distilled=distilled, **kw)
        ours = ovit.TimmVisionTransformer(distilled=distilled, **kw)
        ours.load_state_dict(ref.state_dict())
        assert torch.equal(ref(x), ours(x))
    from vit_torch_b200 import models as pm
    shapes = {"deit_tiny_patch16_224": (192, 3, False, 224), "deit_small_patch16_224": (384, 6, False, 224),
              "deit_base_patch16_224": (768, 12, False, 224), "deit_tiny_distilled_patch16_224": (192, 3, True, 224),
              "deit_small_distilled_patch16_224": (384, 6, True, 224),
              "deit_base_distilled_patch16_224": (768, 12, True, 224), "deit_base_patch16_384": (768, 12, False, 384),
              "deit_base_distilled_patch16_384": (768, 12, True, 384)}
    assert sorted(shapes) == sorted(ref_deit.__all__)
    for name in ("deit_tiny_patch16_224", "deit_tiny_distilled_patch16_224", "deit_small_distilled_patch16_224"):
        dim, heads, dist, img = shapes[name]
        ref = getattr(ref_deit, name)(pretrained=False)
        prod = getattr(pm, name)(pretrained=False)
        assert ref.embed_dim == prod.embed_dim == dim and len(ref.blocks) == len(prod.blocks) == 12
        assert ref.blocks[0].attn.num_heads == prod.blocks[0].attn.num_heads == heads
        assert (ref.dist_token is not None) == (prod.dist_token is not None) == dist
        assert ref.patch_embed.img_size == prod.patch_embed.img_size == (img, img)
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
               {k: tuple(v.shape) for k, v in prod.state_dict().items()}


# ----------------------------------------------------------------------------------------------------------------
# In-tree witnesses of the un-vendored timm / DINO pieces: models/swin.py carries its own copies of Mlp (:14-30),
# the attention operation order (:119-144) and the conv PatchEmbed (:410-448).
# ----------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_swin():
    return _import_reference("reference_swin", "swin.py")


def test_mlp_and_patch_embed_witnesses_bit_exact(ref_swin):
    from oracle import vit as ovit
    torch.manual_seed(2)
    ref, ours = ref_swin.Mlp(96, 384), ovit.Mlp(96, 384)
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(3, 17, 96)
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert torch.equal(yr, yo)
    g = torch.randn_like(yr)
    yr.backward(g); yo.backward(g)
    assert torch.equal(xr.grad, xo.grad)
    for (k, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        assert torch.equal(p.grad, q.grad), k
    pr = ref_swin.PatchEmbed(img_size=64, patch_size=16, in_chans=3, embed_dim=48)
    po = ovit.PatchEmbed(img_size=64, patch_size=16, in_chans=3, embed_dim=48, strict_size=True)
    po.load_state_dict(pr.state_dict())
    img = torch.randn(2, 3, 64, 64)
    assert torch.equal(pr(img), po(img))
    with pytest.raises(AssertionError):
        po(torch.randn(1, 3, 32, 32))       # the size assert of models/swin.py:441-443 / timm PatchEmbed


def test_attention_witness(ref_swin):
    """models/swin.py:119-144 with a zero relative-position table and no mask is softmax((q*scale) k^T) v + proj: the
    same function as the oracle's timm/DINO Attention, which scales AFTER q k^T -- equal to fp32
    rounding, not bit-exact, because of that operation order."""
    from oracle import vit as ovit
    torch.manual_seed(3)
    dim, heads, ws = 96, 2, 7
    ref = ref_swin.WindowAttention(dim, window_size=(ws, ws), num_heads=heads, qkv_bias=True)
    with torch.no_grad():
        ref.relative_position_bias_table.zero_()
    ours = ovit.Attention(dim, num_heads=heads, qkv_bias=True)
    ours.load_state_dict({k: v for k, v in ref.state_dict().items() if k.startswith(("qkv", "proj"))})
    assert abs(ours.scale - ref.scale) < 1e-12
    x = torch.randn(4, ws * ws, dim)
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert torch.allclose(yr, yo, rtol=1e-5, atol=1e-6)
    g = torch.randn_like(yr)
    yr.backward(g); yo.backward(g)
    assert torch.allclose(xr.grad, xo.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(ref.qkv.weight.grad, ours.qkv.weight.grad, rtol=1e-4, atol=1e-6)
