"""nn.Modules with reference-identical parameter names whose forward runs the fused sm_100a kernels.

Sub-modules (nn.Linear / nn.LayerNorm / nn.Conv2d) are kept as plain parameter containers so that the reference's
model zoo can walk `.children()`, call `reset_parameters()`, read `patch_embed.proj.{out_channels,stride,...}` and swap
`patch_embed.proj` / `head` (models/vision_all.py:157-174,194-213,322-329), and so that upstream DINO / timm / CaiT
state_dicts load unchanged. Their own forward() is never used on the hot path: Block.forward hands their parameters
to functional.BlockFn.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn


class DropPath(nn.Module):
    """Stochastic depth. The per-sample mask/keep_prob vector is drawn on the host side (torch RNG, like the
    reference: floor(keep + U[0,1))) and folded into the GEMM epilogue by the owning Block."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob or 0.0)

    def rowscale(self, batch: int, device, training: bool):
        if self.drop_prob == 0.0 or not training:
            return None
        keep = 1.0 - self.drop_prob
        return (keep + torch.rand((batch,), dtype=torch.float32, device=device)).floor_().div_(keep)

    def forward(self, x):  # standalone use (not on the fused path)
        rs = self.rowscale(x.shape[0], x.device, self.training)
        return x if rs is None else x * rs.view(-1, *([1] * (x.dim() - 1)))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("fused Mlp implements the exact-erf GELU used by every model on the path")
        if drop:
            raise NotImplementedError("Mlp dropout > 0 is not on the reference path (zoo passes 0)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("attention dropout > 0 is not on the reference path (zoo passes 0)")
        head_dim = dim // num_heads
        if head_dim not in (64, 48):
            raise NotImplementedError(f"fused attention supports head_dim 64 / 48, got {head_dim}")
        self.num_heads = num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(nn.Module):
    """Pre-LN transformer block (timm / DINO `Block`); optional LayerScale gammas (`init_values`)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, init_values=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if init_values is not None:
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))
        else:
            self.gamma_1 = self.gamma_2 = None

    def forward(self, x):
        rs = None
        if isinstance(self.drop_path, DropPath) and self.training and self.drop_path.drop_prob > 0.0:
            # one independent draw per residual branch, as the reference's two drop_path() calls
            rs = (self.drop_path.rowscale(x.shape[0], x.device, True),
                  self.drop_path.rowscale(x.shape[0], x.device, True))
        a, m = self.attn, self.mlp
        return Fn.BlockFn.apply(x, a.num_heads, self.norm1.eps, a.scale, rs, self.norm1.weight, self.norm1.bias,
                                a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, self.norm2.weight,
                                self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, self.gamma_1,
                                self.gamma_2, None, None, None, None)


class PatchEmbed(nn.Module):
    """Holds the Conv2d(C, D, kernel=P, stride=P) parameters; the conv itself runs as a patch GEMM in TokensFn."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, strict_size=False):
        super().__init__()
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.num_patches = (self.img_size[0] // self.patch_size[0]) * (self.img_size[1] // self.patch_size[1])
        self.strict_size = strict_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.input_norm = None   # (mean[C], std[C]) device vectors: uint8 images are normalised on the device

    def set_input_normalization(self, mean, std):
        """Device input pipeline: feed raw uint8 [B,C,H,W] images; ToTensor + Normalize(mean, std)
        (utils_datasets.py:573-580) run on the GPU in front of the patch GEMM. Not part of the state_dict."""
        dev = self.proj.weight.device
        self.input_norm = (torch.as_tensor(mean, dtype=torch.float32, device=dev).contiguous(),
                           torch.as_tensor(std, dtype=torch.float32, device=dev).contiguous())
        return self

    def norm_for(self, x):
        if x.dtype != torch.uint8:
            return None
        if self.input_norm is None:
            raise ValueError("uint8 input: call patch_embed.set_input_normalization(mean, std) first")
        if self.input_norm[0].device != x.device:
            self.input_norm = tuple(t.to(x.device) for t in self.input_norm)
        return self.input_norm

    def check(self, x):
        P = self.patch_size[0]
        proj = self.proj
        if tuple(proj.kernel_size) != (P, P) or tuple(proj.stride) != (P, P) or tuple(proj.padding) != (0, 0):
            raise NotImplementedError("PatchEmbed kernel expects Conv2d(kernel=P, stride=P, padding=0)")
        if x.shape[1] != proj.in_channels:
            raise ValueError(f"input has {x.shape[1]} channels, patch_embed.proj expects {proj.in_channels}")
        if self.strict_size and tuple(x.shape[-2:]) != self.img_size:
            raise AssertionError(f"Input image size {tuple(x.shape[-2:])} doesn't match model {self.img_size}")
        if x.shape[-2] % P or x.shape[-1] % P:
            raise ValueError("image size must be a multiple of the patch size")

    def forward(self, x):  # standalone: tokens without prefix / pos
        self.check(x)
        D = self.proj.out_channels
        n = (x.shape[-2] // self.patch_size[0]) * (x.shape[-1] // self.patch_size[0])
        zero_pos = torch.zeros((1, n, D), dtype=torch.float32, device=x.device)
        return Fn.TokensFn.apply(x, self.proj.weight, self.proj.bias, zero_pos, None, self.patch_size[0],
                                 self.norm_for(x))
