"""Drop-in boundary: the reference's own models/vision_all.py, imported unchanged, must receive the fused
models through the torch.hub / timm shims. Construction only (no GPU needed); needs /root/reference."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def zoo(tmp_path_factory):
    from vit_torch_b200 import compat
    home = str(tmp_path_factory.mktemp("torch_home"))
    compat.install_hub_shim(home)
    assert compat.install_timm_shim()
    sys.path.insert(0, "/root/reference")
    try:
        from models.vision_all import VisionModelZoo
    finally:
        sys.path.remove("/root/reference")
    return VisionModelZoo, home


def test_dino_through_reference_zoo(zoo):
    from vit_torch_b200 import models, modules
    Zoo, home = zoo
    m = Zoo.get_model("dino_vits16", pretrained=False, classifier=[256, 128, 32, 10], root_path=home)
    assert isinstance(m, models.DinoVisionTransformer)
    assert isinstance(m.blocks[0], modules.Block)
    # head attached by models/vision_all.py:168-174: Sequential(Linear+GELU x3, Linear(bias=False))
    assert isinstance(m.head, torch.nn.Sequential) and m.head[-1].bias is None and m.head[0].in_features == 384
    assert os.environ["TORCH_HOME"] == home
    keys = m.state_dict().keys()
    assert "blocks.11.mlp.fc2.weight" in keys and "pos_embed" in keys


def test_dino_channel_swap_and_lineareval(zoo):
    Zoo, home = zoo
    m = Zoo.get_model("dino_vitb8", pretrained=False, image_channels=7, classifier=None, root_path=home)
    assert m.patch_embed.proj.in_channels == 7 and tuple(m.patch_embed.proj.kernel_size) == (8, 8)
    assert isinstance(m.head, torch.nn.Identity)
    head = Zoo.get_model(arch=None, image_channels=768, classifier=[256, 10])
    assert isinstance(head, torch.nn.Sequential)
