"""timm.models.vision_transformer symbols the reference imports: VisionTransformer, _cfg, Mlp, PatchEmbed."""
from vit_torch_b200.models import VisionTransformer  # noqa: F401
from vit_torch_b200.modules import Attention, Block, Mlp, PatchEmbed  # noqa: F401


def _cfg(url="", **kwargs):
    cfg = {"url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None, "crop_pct": 0.9,
           "interpolation": "bicubic", "mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225),
           "first_conv": "patch_embed.proj", "classifier": "head"}
    cfg.update(kwargs)
    return cfg
