"""Minimal `timm` stand-in: only the symbols khuongnd6/ViT_torch imports (models/cait.py:8-10, models/deit.py:7-9,
models/swin.py:11, models/xcit.py:16-18, models/vision_all.py:3,13). create_model() returns the fused sm_100a models."""
__vit_torch_b200_shim__ = True
__version__ = "0.4.12+vit_torch_b200.shim"

from . import models  # noqa: F401
from .models.factory import create_model  # noqa: F401
