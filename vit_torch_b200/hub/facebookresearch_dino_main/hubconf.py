"""Offline stand-in for the hubconf.py of facebookresearch/dino (reference call site models/vision_all.py:156:
`torch.hub.load('facebookresearch/dino:main', arch, pretrained=bool(pretrained))`).

Place (symlink or copy via vit_torch_b200.compat.install_hub_shim) this directory at
`$TORCH_HOME/hub/facebookresearch_dino_main/`; torch.hub then resolves the entrypoints below without network access
and the reference's model zoo receives the fused sm_100a models. The reference sets TORCH_HOME from --root_path
(main.py:111, models/vision_all.py:88-89).
"""
import os
import sys

dependencies = ["torch"]

_ROOT = os.environ.get("VIT_TORCH_B200_ROOT") or "@VIT_TORCH_B200_ROOT@"
if not os.path.isdir(_ROOT):
    _ROOT = os.path.realpath(os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from vit_torch_b200.models import dino_vitb8, dino_vitb16, dino_vits8, dino_vits16  # noqa: E402,F401
