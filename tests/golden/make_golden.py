"""Generate golden vectors from the REFERENCE's own models/cait.py (run in the build container only).

    python tests/golden/make_golden.py

The reference tree is not shipped to the GPU box, so the vectors are committed: tests/golden/cait_xxs_tiny.pt holds
seeded inputs, a state_dict and the reference's outputs / gradients for a shrunken CaiT built from the reference's
classes (cait_models with embed_dim=96, depth=2, heads=2 -> head_dim 48, 2 class-attention blocks, 64x64 images), and
tests/golden/cait_blocks.pt holds per-module vectors for Attention_talking_head, Class_Attention, LayerScale_Block and
LayerScale_Block_CA at head_dim 48 (dim 96, 2 heads, N=196). Weights and inputs are rounded to bf16 before the
reference runs (stored exactly in 2 bytes); results are stored as per-tensor max-abs scale + fp16.
"""
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference_cait():
    """Import /root/reference/models/cait.py with a stub of the five timm symbols it uses (models/cait.py:8-10)."""
    from oracle import vit as ovit
    tm = types.ModuleType("timm")
    tmm = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.Mlp, vt.PatchEmbed = ovit.Mlp, ovit.PatchEmbed
    vt._cfg = lambda **kw: dict(kw)
    reg = types.ModuleType("timm.models.registry")
    reg.register_model = lambda fn: fn
    lay = types.ModuleType("timm.models.layers")
    lay.trunc_normal_ = ovit.trunc_normal_
    lay.DropPath = ovit.DropPath
    stubs = {"timm": tm, "timm.models": tmm, "timm.models.vision_transformer": vt, "timm.models.registry": reg,
             "timm.models.layers": lay}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)          # force the oracle-backed stubs even if another timm (or shim) is loaded
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("reference_cait", os.path.join(REF, "models", "cait.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def pack(t):
    """Per-tensor max-abs scale + fp16 mantissas: 2 bytes/element, ~5e-4 relative-to-max quantisation."""
    t = t.detach().float()
    scale = t.abs().max().clamp_min(1e-30)
    return {"scale": scale.clone(), "q": (t / scale).to(torch.float16)}


def bf16_exact_(module):
    """Round every parameter to a bf16-representable value so that the fixture can store weights in 2 bytes exactly."""
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(p.to(torch.bfloat16).float())


def main():
    from functools import partial
    ref = import_reference_cait()
    torch.manual_seed(1234)
    norm = partial(nn.LayerNorm, eps=1e-6)

    # ---- whole (shrunken) model
    m = ref.cait_models(img_size=64, patch_size=16, embed_dim=96, depth=2, num_heads=2, mlp_ratio=4, qkv_bias=True,
                        norm_layer=norm, init_scale=0.1, depth_token_only=2, num_classes=10)
    with torch.no_grad():  # make every parameter non-trivial (biases, proj_l / proj_w, gammas differ per channel)
        for n_, p in m.named_parameters():
            if p.dim() == 1 and "norm" not in n_:
                p.add_(torch.randn_like(p) * 0.05)
    bf16_exact_(m)
    x = torch.randn(3, 3, 64, 64).to(torch.bfloat16).float()
    y = torch.tensor([1, 7, 3])
    out = m(x)
    loss = nn.functional.cross_entropy(out, y)
    loss.backward()
    torch.save({"state_dict": {k: v.to(torch.bfloat16) for k, v in m.state_dict().items()}, "x": x.to(torch.bfloat16),
                "y": y, "out": out.detach(), "loss": loss.detach(),
                "grads": {k: pack(p.grad) for k, p in m.named_parameters()},
                "cfg": dict(img_size=64, patch_size=16, embed_dim=96, depth=2, num_heads=2, init_scale=0.1,
                            num_classes=10)},
               os.path.join(HERE, "cait_xxs_tiny.pt"))

    # ---- per-module vectors at head_dim 48 (dim 96, 2 heads), N = 196 tokens
    blobs = {}
    dim, heads, N, B = 96, 2, 196, 2
    for name, ctor, is_ca in [
        ("talking_head", lambda: ref.Attention_talking_head(dim, num_heads=heads, qkv_bias=True), False),
        ("class_attention", lambda: ref.Class_Attention(dim, num_heads=heads, qkv_bias=True), False),
        ("layerscale_block", lambda: ref.LayerScale_Block(dim, heads, qkv_bias=True, norm_layer=norm, init_values=0.1), False),
        ("layerscale_block_ca", lambda: ref.LayerScale_Block_CA(dim, heads, qkv_bias=True, norm_layer=norm, init_values=0.1), True),
    ]:
        mod = ctor()
        with torch.no_grad():  # non-trivial values everywhere (biases / proj_l / proj_w / gammas included)
            for n_, p in mod.named_parameters():
                if p.dim() == 1 and "norm" not in n_:
                    p.add_(torch.randn_like(p) * 0.1)
        bf16_exact_(mod)
        xin = torch.randn(B, N + (1 if name == "class_attention" else 0), dim).to(torch.bfloat16).float().requires_grad_(True)
        if is_ca:
            cls = torch.randn(B, 1, dim).to(torch.bfloat16).float().requires_grad_(True)
            o = mod(xin, cls)
        else:
            cls = None
            o = mod(xin)
        go = torch.randn_like(o).to(torch.bfloat16).float()
        o.backward(go)
        blobs[name] = {"state_dict": {k: v.to(torch.bfloat16) for k, v in mod.state_dict().items()},
                       "x": xin.detach().to(torch.bfloat16),
                       "cls": None if cls is None else cls.detach().to(torch.bfloat16), "out": pack(o),
                       "gout": go.to(torch.bfloat16), "gx": pack(xin.grad),
                       "gcls": None if cls is None else pack(cls.grad),
                       "grads": {k: pack(p.grad) for k, p in mod.named_parameters()}}
    # weights / inputs are bf16-exact (stored in bf16), results are stored as per-tensor-scaled fp16
    torch.save(blobs, os.path.join(HERE, "cait_blocks.pt"))
    for f in ("cait_xxs_tiny.pt", "cait_blocks.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
