"""Model constructors behind the reference's plugin boundary.

  * dino_vits16 / dino_vits8 / dino_vitb16 / dino_vitb8  -- torch.hub entrypoints of facebookresearch/dino
    (reference call site models/vision_all.py:156); forward returns norm(x)[:, 0] and never applies `head`
     unless `apply_head=True` is passed.
  * VisionTransformer / DistilledVisionTransformer + deit_* -- timm ~0.4.12 ViT as subclassed by models/deit.py:20-91.
  * cait_* -- see cait.py.
Parameter names match upstream so that pretrained state_dicts load (`blocks.i.attn.qkv.weight`, `patch_embed.proj.*`,
`cls_token`, `pos_embed`, `norm.*`, `head.*`).
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .modules import Block, PatchEmbed


def load_pretrained(model, url, key=None, strip_prefix="", check_hash=False):
    """The reference's / upstream's `pretrained=True` path: torch.hub.load_state_dict_from_url(url, map_location="cpu")
    (models/cait.py:264-273, models/deit.py:100-105, DINO hubconf) and a strict load_state_dict. Works offline when the
    file is already in $TORCH_HOME/hub/checkpoints/ (torch.hub only downloads what is not cached; TORCH_HOME is the
    reference's --root_path, main.py:111); otherwise torch.hub raises its own download error, as the reference would.
    `key`: sub-dict of the checkpoint ("model"); `strip_prefix`: e.g. "module." (CaiT checkpoints were saved from
    DistributedDataParallel, models/cait.py:270-271)."""
    ckpt = torch.hub.load_state_dict_from_url(url=url, map_location="cpu", check_hash=check_hash)
    sd = ckpt[key] if key is not None else ckpt
    if strip_prefix:
        sd = {k: sd[strip_prefix + k] for k in model.state_dict().keys()}
    model.load_state_dict(sd, strict=True)
    return model


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Truncated-normal init used by DINO / timm / CaiT (`trunc_normal_(w, std=.02)`, models/cait.py:209-216)."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _init_weights(m):
    if isinstance(m, nn.Linear):
        trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class DinoVisionTransformer(nn.Module):
    """facebookresearch/dino VisionTransformer on fused sm_100a kernels."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 norm_layer=nn.LayerNorm, apply_head=False, **_unused):
        super().__init__()
        if isinstance(img_size, (list, tuple)):
            img_size = img_size[0]
        self.num_features = self.embed_dim = embed_dim
        self.apply_head = apply_head
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [r.item() for r in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(_init_weights)

    def interpolate_pos_encoding(self, npatch, w, h):
        """Bicubic resize of the patch position grid for inputs that are not the training resolution (DINO)."""
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        class_pos, patch_pos = self.pos_embed[:, 0], self.pos_embed[:, 1:]
        dim = self.pos_embed.shape[-1]
        P = self.patch_embed.patch_size[0]
        w0, h0 = w // P + 0.1, h // P + 0.1
        side = int(math.sqrt(N))
        patch_pos = F.interpolate(patch_pos.reshape(1, side, side, dim).permute(0, 3, 1, 2),
                                  scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
        assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
        patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
        return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)

    def prepare_tokens(self, x):
        self.patch_embed.check(x)
        _, _, w, h = x.shape
        P = self.patch_embed.patch_size[0]
        pos = self.interpolate_pos_encoding((w // P) * (h // P), w, h)
        x = Fn.TokensFn.apply(x, self.patch_embed.proj.weight, self.patch_embed.proj.bias, pos, self.cls_token, P,
                              self.patch_embed.norm_for(x))
        return self.pos_drop(x)

    def forward(self, x):
        x = self.prepare_tokens(x)
        for blk in self.blocks:
            x = blk(x)
        x = Fn.TokenNormFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps, 0)
        return self.head(x) if self.apply_head else x


_DINO = {"vits": dict(embed_dim=384, depth=12, num_heads=6), "vitb": dict(embed_dim=768, depth=12, num_heads=12)}


# upstream hubconf.py checkpoint files (facebookresearch/dino@main)
_DINO_URLS = {
    ("vits", 16): "https://dl.fbaipublicfiles.com/dino/dino_deitsmall16_pretrain/dino_deitsmall16_pretrain.pth",
    ("vits", 8): "https://dl.fbaipublicfiles.com/dino/dino_deitsmall8_pretrain/dino_deitsmall8_pretrain.pth",
    ("vitb", 16): "https://dl.fbaipublicfiles.com/dino/dino_vitbase16_pretrain/dino_vitbase16_pretrain.pth",
    ("vitb", 8): "https://dl.fbaipublicfiles.com/dino/dino_vitbase8_pretrain/dino_vitbase8_pretrain.pth",
}


def _dino(kind, patch, pretrained, **kw):
    model = DinoVisionTransformer(patch_size=patch, num_classes=0, mlp_ratio=4, qkv_bias=True,
                                  norm_layer=partial(nn.LayerNorm, eps=1e-6), **_DINO[kind], **kw)
    if pretrained:
        load_pretrained(model, _DINO_URLS[(kind, patch)])
    return model


def dino_vits16(pretrained=True, **kwargs):
    return _dino("vits", 16, pretrained, **kwargs)


def dino_vits8(pretrained=True, **kwargs):
    return _dino("vits", 8, pretrained, **kwargs)


def dino_vitb16(pretrained=True, **kwargs):
    return _dino("vitb", 16, pretrained, **kwargs)


def dino_vitb8(pretrained=True, **kwargs):
    return _dino("vitb", 8, pretrained, **kwargs)


class VisionTransformer(nn.Module):
    """timm ~0.4.12 VisionTransformer (the base class models/deit.py:20,63 subclasses): forward applies `head`;
    `distilled=True` adds dist_token / head_dist and returns the average of both heads (models/deit.py:67-91)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=None, **_unused):
        super().__init__()
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 2 if distilled else 1
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      strict_size=True)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if distilled else None
        self.pos_embed = nn.Parameter(torch.zeros(1, n + self.num_tokens, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [r.item() for r in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.pre_logits = nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.head_dist = None
        if distilled:
            self.head_dist = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        if distilled:
            trunc_normal_(self.dist_token, std=0.02)
        self.apply(_init_weights)

    def forward_features(self, x):
        self.patch_embed.check(x)
        prefix = self.cls_token if self.dist_token is None else torch.cat((self.cls_token, self.dist_token), dim=1)
        x = Fn.TokensFn.apply(x, self.patch_embed.proj.weight, self.patch_embed.proj.bias, self.pos_embed, prefix,
                              self.patch_embed.patch_size[0], self.patch_embed.norm_for(x))
        x = self.pos_drop(x)
        x = self.blocks(x)
        cls = Fn.TokenNormFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps, 0)
        if self.dist_token is None:
            return self.pre_logits(cls)
        return cls, Fn.TokenNormFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps, 1)

    def forward(self, x):
        x = self.forward_features(x)
        if self.head_dist is not None:
            return (self.head(x[0]) + self.head_dist(x[1])) / 2
        return self.head(x)


# checkpoint files of models/deit.py:100-209
_DEIT_URLS = {
    (192, False, 224): "deit_tiny_patch16_224-a1311bcf.pth", (384, False, 224): "deit_small_patch16_224-cd65a155.pth",
    (768, False, 224): "deit_base_patch16_224-b5f2ef4d.pth",
    (192, True, 224): "deit_tiny_distilled_patch16_224-b40b3cf7.pth",
    (384, True, 224): "deit_small_distilled_patch16_224-649709d9.pth",
    (768, True, 224): "deit_base_distilled_patch16_224-df68dfff.pth",
    (768, False, 384): "deit_base_patch16_384-8de9b5d1.pth",
    (768, True, 384): "deit_base_distilled_patch16_384-d0272ac0.pth",
}


def _deit(embed_dim, num_heads, distilled, img_size=224, pretrained=False, **kw):
    model = VisionTransformer(img_size=img_size, patch_size=16, embed_dim=embed_dim, depth=12, num_heads=num_heads,
                              mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                              distilled=distilled, **kw)
    if pretrained:
        load_pretrained(model, "https://dl.fbaipublicfiles.com/deit/" + _DEIT_URLS[(embed_dim, distilled, img_size)],
                        key="model", check_hash=True)
    return model


def deit_tiny_patch16_224(pretrained=False, **kw):
    return _deit(192, 3, False, pretrained=pretrained, **kw)


def deit_small_patch16_224(pretrained=False, **kw):
    return _deit(384, 6, False, pretrained=pretrained, **kw)


def deit_base_patch16_224(pretrained=False, **kw):
    return _deit(768, 12, False, pretrained=pretrained, **kw)


def deit_tiny_distilled_patch16_224(pretrained=False, **kw):
    return _deit(192, 3, True, pretrained=pretrained, **kw)


def deit_small_distilled_patch16_224(pretrained=False, **kw):
    return _deit(384, 6, True, pretrained=pretrained, **kw)


def deit_base_distilled_patch16_224(pretrained=False, **kw):
    return _deit(768, 12, True, pretrained=pretrained, **kw)


def deit_base_patch16_384(pretrained=False, **kw):
    return _deit(768, 12, False, img_size=384, pretrained=pretrained, **kw)


def deit_base_distilled_patch16_384(pretrained=False, **kw):
    return _deit(768, 12, True, img_size=384, pretrained=pretrained, **kw)
