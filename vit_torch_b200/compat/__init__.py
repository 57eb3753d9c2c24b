"""Compatibility shims that let the reference's own files import and dispatch to the fused models offline:

  * install_hub_shim(torch_home): makes `torch.hub.load('facebookresearch/dino:main', ...)` resolve to
    vit_torch_b200.models (models/vision_all.py:156).
  * install_timm_shim(): puts a minimal `timm` package (only the symbols the reference imports: SURVEY 8b) on sys.path
    when the real timm is not installed; its create_model() prefers the fused constructors.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))


def install_hub_shim(torch_home: str) -> str:
    """Create $TORCH_HOME/hub/facebookresearch_dino_main/hubconf.py pointing at this checkout. Returns the directory."""
    src = os.path.join(os.path.dirname(_HERE), "hub", "facebookresearch_dino_main", "hubconf.py")
    dst_dir = os.path.join(torch_home, "hub", "facebookresearch_dino_main")
    os.makedirs(dst_dir, exist_ok=True)
    with open(src) as f:
        text = f.read().replace("@VIT_TORCH_B200_ROOT@", _ROOT)
    with open(os.path.join(dst_dir, "hubconf.py"), "w") as f:
        f.write(text)
    return dst_dir


def install_timm_shim(force: bool = False) -> bool:
    """Make `import timm` resolve to the bundled minimal shim unless a real timm is importable. Returns True if the
    shim is (now) the active timm."""
    if not force and "timm" not in sys.modules and importlib.util.find_spec("timm") is not None:
        return False
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import timm  # noqa: F401

    return getattr(sys.modules["timm"], "__vit_torch_b200_shim__", False)


__all__ = ["install_hub_shim", "install_timm_shim", "shutil"]
