"""CaiT on the fused sm_100a kernels: same constructors, kwargs, parameter names and forward as the reference's
models/cait.py (cait_models :155-253, LayerScale_Block :130-150, Attention_talking_head :87-128,
LayerScale_Block_CA :57-84, Class_Attention :21-55, constructors :255-480)."""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import functional as Fn
from .models import load_pretrained, trunc_normal_
from .modules import Block, DropPath, Mlp, PatchEmbed


class Attention_talking_head(nn.Module):
    """Parameter holder for talking-heads attention (qkv, proj, proj_l, proj_w); computed inside BlockFn."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("attention dropout > 0 is not on the reference path (zoo passes 0)")
        if dim // num_heads not in (48, 64):
            raise NotImplementedError("fused talking-heads attention supports head_dim 48 / 64")
        if num_heads not in (2, 4, 6, 8, 16):
            raise NotImplementedError("talking-heads mixing kernel is instantiated for 2/4/6/8/16 heads")
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_l = nn.Linear(num_heads, num_heads)
        self.proj_w = nn.Linear(num_heads, num_heads)
        self.proj_drop = nn.Dropout(proj_drop)


class Class_Attention(nn.Module):
    """Parameter holder for class attention (q, k, v, proj); computed inside ClassAttnBlockFn."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("attention dropout > 0 is not on the reference path (zoo passes 0)")
        if dim // num_heads not in (48, 64):
            raise NotImplementedError("fused class attention supports head_dim 48 / 64")
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class LayerScale_Block(nn.Module):
    """x += g1 * TalkingHeads(LN1 x); x += g2 * Mlp(LN2 x)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, Attention_block=Attention_talking_head,
                 Mlp_block=Mlp, init_values=1e-4):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention_block(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                    attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp_block(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)

    def forward(self, x):
        rs = None
        if isinstance(self.drop_path, DropPath) and self.training and self.drop_path.drop_prob > 0.0:
            # one independent draw per residual branch, as the reference's two drop_path() calls
            rs = (self.drop_path.rowscale(x.shape[0], x.device, True),
                  self.drop_path.rowscale(x.shape[0], x.device, True))
        a, m = self.attn, self.mlp
        th = isinstance(a, Attention_talking_head)
        return Fn.BlockFn.apply(x, a.num_heads, self.norm1.eps, a.scale, rs, self.norm1.weight, self.norm1.bias,
                                a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, self.norm2.weight,
                                self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, self.gamma_1,
                                self.gamma_2, a.proj_l.weight if th else None, a.proj_l.bias if th else None,
                                a.proj_w.weight if th else None, a.proj_w.bias if th else None)


class LayerScale_Block_CA(nn.Module):
    """cls += g1 * ClassAttn(LN1 cat(cls, x)); cls += g2 * Mlp(LN2 cls)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, Attention_block=Class_Attention,
                 Mlp_block=Mlp, init_values=1e-4):
        super().__init__()
        if drop_path:
            raise NotImplementedError("DropPath on the class-attention blocks is never used (models/cait.py:195)")
        self.norm1 = norm_layer(dim)
        self.attn = Attention_block(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                    attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp_block(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)

    def forward(self, x, x_cls):
        a, m = self.attn, self.mlp
        return Fn.ClassAttnBlockFn.apply(x, x_cls, a.num_heads, self.norm1.eps, a.scale, self.norm1.weight,
                                         self.norm1.bias, a.q.weight, a.q.bias, a.k.weight, a.k.bias, a.v.weight,
                                         a.v.bias, a.proj.weight, a.proj.bias, self.norm2.weight, self.norm2.bias,
                                         m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, self.gamma_1, self.gamma_2)


class cait_models(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=nn.LayerNorm, global_pool=None, block_layers=LayerScale_Block,
                 block_layers_token=LayerScale_Block_CA, Patch_layer=PatchEmbed, act_layer=nn.GELU,
                 Attention_block=Attention_talking_head, Mlp_block=Mlp, init_scale=1e-4,
                 Attention_block_token_only=Class_Attention, Mlp_block_token_only=Mlp, depth_token_only=2,
                 mlp_ratio_clstk=4.0):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = Patch_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        self.patch_embed.strict_size = True  # timm PatchEmbed asserts the input size
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [drop_path_rate for _ in range(depth)]
        self.blocks = nn.ModuleList([
            block_layers(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                         drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                         act_layer=act_layer, Attention_block=Attention_block, Mlp_block=Mlp_block,
                         init_values=init_scale)
            for i in range(depth)])
        self.blocks_token_only = nn.ModuleList([
            block_layers_token(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio_clstk, qkv_bias=qkv_bias,
                               qk_scale=qk_scale, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=norm_layer,
                               act_layer=act_layer, Attention_block=Attention_block_token_only,
                               Mlp_block=Mlp_block_token_only, init_values=init_scale)
            for _ in range(depth_token_only)])
        self.norm = norm_layer(embed_dim)
        self.feature_info = [dict(num_chs=embed_dim, reduction=0, module="head")]
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def forward_features(self, x):
        B = x.shape[0]
        self.patch_embed.check(x)
        x = Fn.TokensFn.apply(x, self.patch_embed.proj.weight, self.patch_embed.proj.bias, self.pos_embed, None,
                              self.patch_embed.patch_size[0], self.patch_embed.norm_for(x))
        x = self.pos_drop(x)
        for blk in self.blocks:
            x = blk(x)
        cls_tokens = self.cls_token.expand(B, -1, -1)
        for blk in self.blocks_token_only:
            cls_tokens = blk(x, cls_tokens)
        # norm(cat(cls, x))[:, 0] == norm(cls)[:, 0]: LayerNorm is per token
        return Fn.TokenNormFn.apply(cls_tokens, self.norm.weight, self.norm.bias, self.norm.eps, 0)

    def forward(self, x):
        return self.head(self.forward_features(x))


# name -> (img_size, embed_dim, depth, num_heads, init_scale, checkpoint file)   (models/cait.py:255-480)
_SIZES = {
    "cait_XXS24_224": (224, 192, 24, 4, 1e-5, "XXS24_224.pth"), "cait_XXS24": (384, 192, 24, 4, 1e-5, "XXS24_384.pth"),
    "cait_XXS36_224": (224, 192, 36, 4, 1e-5, "XXS36_224.pth"), "cait_XXS36": (384, 192, 36, 4, 1e-5, "XXS36_384.pth"),
    "cait_XS24": (384, 288, 24, 6, 1e-5, "XS24_384.pth"), "cait_S24_224": (224, 384, 24, 8, 1e-5, "S24_224.pth"),
    "cait_S24": (384, 384, 24, 8, 1e-5, "S24_384.pth"), "cait_S36": (384, 384, 36, 8, 1e-6, "S36_384.pth"),
    "cait_M36": (384, 768, 36, 16, 1e-6, "M36_384.pth"), "cait_M48": (448, 768, 48, 16, 1e-6, "M48_448.pth"),
}


def _make(name):
    img, dim, depth, heads, init, ckpt = _SIZES[name]

    def ctor(pretrained=False, **kwargs):
        model = cait_models(img_size=img, patch_size=16, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=4,
                            qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), init_scale=init,
                            depth_token_only=2, **kwargs)
        if pretrained:      # models/cait.py:264-273: checkpoint["model"]["module." + key]
            load_pretrained(model, "https://dl.fbaipublicfiles.com/deit/" + ckpt, key="model", strip_prefix="module.",
                            check_hash=True)
        return model

    ctor.__name__ = name
    ctor.__doc__ = f"{name}: img {img}, dim {dim}, depth {depth}+2, heads {heads}, LayerScale init {init}"
    return ctor


for _n in _SIZES:
    globals()[_n] = _make(_n)
__all__ = ["cait_models", "LayerScale_Block", "LayerScale_Block_CA", "Attention_talking_head", "Class_Attention",
           *_SIZES]
