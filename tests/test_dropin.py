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


def test_cait_through_reference_zoo(zoo):
    """models/vision_all.py:184-221 with the timm shim: create_model('cait_S24_224', num_classes=1000, drop_rate=0.,
    drop_path_rate=0., drop_block_rate=None) must dispatch to the fused CaiT (case-sensitive name), accept the conv
    swap for image_channels != 3 (:194-202) and the head / head_dist assignment (:204-213)."""
    from vit_torch_b200 import cait, zoo as pzoo
    Zoo, home = zoo
    m = Zoo.get_model("cait_S24_224", pretrained=False, classifier=[64, 10], root_path=home)
    assert isinstance(m, cait.cait_models) and len(m.blocks) == 24 and len(m.blocks_token_only) == 2
    assert m.blocks[0].attn.num_heads == 8 and m.embed_dim == 384
    assert isinstance(m.head, torch.nn.Sequential) and m.head[-1].bias is None and m.head[0].in_features == 384
    assert m.head_dist is m.head
    assert abs(m.blocks[0].gamma_1[0].item() - 1e-5) < 1e-12
    m7 = Zoo.get_model("cait_XXS24_224", pretrained=False, image_channels=7, classifier=False, root_path=home)
    assert m7.patch_embed.proj.in_channels == 7 and isinstance(m7.head, torch.nn.Identity)
    # with the zoo patched the attached head is the fused one
    original = Zoo.__dict__["get_classifier_head"]
    pzoo.patch_reference_zoo(Zoo)
    try:
        m = Zoo.get_model("cait_XXS24_224", pretrained=False, classifier=[64, 10], root_path=home)
        assert isinstance(m.head, pzoo.ClassifierHead)
        head = Zoo.get_classifier_head(768, [256, 128, 32, 10])       # main.py:196-201 (lineareval head)
        assert isinstance(head, pzoo.ClassifierHead) and head[-1].bias is None
        assert list(head.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias",
                                                  "6.weight"]
    finally:
        Zoo.get_classifier_head = original      # back to the reference's own classmethod for the other tests
