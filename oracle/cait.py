"""fp32 PyTorch restatement of the reference's CaiT (TEST INFRASTRUCTURE; see oracle/__init__.py).

Follows /root/reference models/cait.py: Class_Attention :21-55, LayerScale_Block_CA :57-84, Attention_talking_head
:87-128, LayerScale_Block :130-150, cait_models :155-253, constructors :255-480. Parameter names and the order of
floating-point operations are kept so that tests/test_oracle_vs_reference.py can demand bit-level agreement with the
reference file (imported there through a stub of the five timm symbols it needs). Written independently: the sizes
live in one table and the two attention variants share helpers.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .vit import DropPath, Mlp, PatchEmbed, trunc_normal_


def _split_heads(t, B, N, H):
    return t.reshape(B, N, H, t.shape[-1] // H).permute(0, 2, 1, 3)


class ClassAttention(nn.Module):
    """One query (the class token) attends over all N tokens (models/cait.py:38-55): q is scaled BEFORE q k^T."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        H = self.num_heads
        q = _split_heads(self.q(x[:, 0]).unsqueeze(1), B, 1, H)
        k = _split_heads(self.k(x), B, N, H)
        q = q * self.scale
        v = _split_heads(self.v(x), B, N, H)
        attn = self.attn_drop((q @ k.transpose(-2, -1)).softmax(dim=-1))
        cls = (attn @ v).transpose(1, 2).reshape(B, 1, C)
        return self.proj_drop(self.proj(cls))


class TalkingHeadAttention(nn.Module):
    """Talking-heads attention (models/cait.py:111-128): logits mixed across heads by proj_l before the softmax and
    probabilities mixed by proj_w after it; q is pre-scaled."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_l = nn.Linear(num_heads, num_heads)
        self.proj_w = nn.Linear(num_heads, num_heads)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        H = self.num_heads
        qkv = self.qkv(x).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = q @ k.transpose(-2, -1)
        attn = self.proj_l(attn.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        attn = attn.softmax(dim=-1)
        attn = self.proj_w(attn.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        attn = self.attn_drop(attn)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class _LayerScaleBase(nn.Module):
    def __init__(self, dim, num_heads, attn_cls, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, act_layer,
                 norm_layer, init_values):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = attn_cls(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                             proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)


class LayerScaleBlock(_LayerScaleBase):
    """x += g1 * TalkingHeads(LN1 x); x += g2 * Mlp(LN2 x)   (models/cait.py:147-150)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, init_values=1e-4):
        super().__init__(dim, num_heads, TalkingHeadAttention, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path,
                         act_layer, norm_layer, init_values)

    def forward(self, x):
        x = x + self.drop_path(self.gamma_1 * self.attn(self.norm1(x)))
        x = x + self.drop_path(self.gamma_2 * self.mlp(self.norm2(x)))
        return x


class LayerScaleBlockCA(_LayerScaleBase):
    """cls += g1 * ClassAttn(LN1 cat(cls, x)); cls += g2 * Mlp(LN2 cls)   (models/cait.py:75-84)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, init_values=1e-4):
        super().__init__(dim, num_heads, ClassAttention, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path,
                         act_layer, norm_layer, init_values)

    def forward(self, x, x_cls):
        u = torch.cat((x_cls, x), dim=1)
        x_cls = x_cls + self.drop_path(self.gamma_1 * self.attn(self.norm1(u)))
        x_cls = x_cls + self.drop_path(self.gamma_2 * self.mlp(self.norm2(x_cls)))
        return x_cls


class CaiT(nn.Module):
    """cait_models (models/cait.py:155-253): patches + pos (no cls yet) -> depth x LayerScaleBlock ->
    depth_token_only x LayerScaleBlockCA on the class token -> LN(cat(cls, x))[:, 0] -> head."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=nn.LayerNorm, init_scale=1e-4, depth_token_only=2, mlp_ratio_clstk=4.0,
                 act_layer=nn.GELU, **_unused):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      strict_size=True)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([
            LayerScaleBlock(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                            qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=drop_path_rate,
                            norm_layer=norm_layer, act_layer=act_layer, init_values=init_scale)
            for _ in range(depth)])
        self.blocks_token_only = nn.ModuleList([
            LayerScaleBlockCA(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio_clstk, qkv_bias=qkv_bias,
                              qk_scale=qk_scale, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=norm_layer,
                              act_layer=act_layer, init_values=init_scale)
            for _ in range(depth_token_only)])
        self.norm = norm_layer(embed_dim)
        self.feature_info = [dict(num_chs=embed_dim, reduction=0, module="head")]
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_features(self, x):
        B = x.shape[0]
        x = self.patch_embed(x)
        cls_tokens = self.cls_token.expand(B, -1, -1)
        x = self.pos_drop(x + self.pos_embed)
        for blk in self.blocks:
            x = blk(x)
        for blk in self.blocks_token_only:
            cls_tokens = blk(x, cls_tokens)
        x = torch.cat((cls_tokens, x), dim=1)
        return self.norm(x)[:, 0]

    def forward(self, x):
        return self.head(self.forward_features(x))


# name -> (img_size, embed_dim, depth, num_heads, init_scale)      (models/cait.py:255-480; head_dim is 48 everywhere)
CAIT_SIZES = {
    "cait_XXS24_224": (224, 192, 24, 4, 1e-5), "cait_XXS24": (384, 192, 24, 4, 1e-5),
    "cait_XXS36_224": (224, 192, 36, 4, 1e-5), "cait_XXS36": (384, 192, 36, 4, 1e-5),
    "cait_XS24": (384, 288, 24, 6, 1e-5), "cait_S24_224": (224, 384, 24, 8, 1e-5), "cait_S24": (384, 384, 24, 8, 1e-5),
    "cait_S36": (384, 384, 36, 8, 1e-6), "cait_M36": (384, 768, 36, 16, 1e-6), "cait_M48": (448, 768, 48, 16, 1e-6),
}


def create(name, pretrained=False, **kwargs):
    assert not pretrained, "no network: pretrained CaiT weights are unavailable"
    img, dim, depth, heads, init = CAIT_SIZES[name]
    return CaiT(img_size=img, patch_size=16, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                norm_layer=partial(nn.LayerNorm, eps=1e-6), init_scale=init, depth_token_only=2, **kwargs)
