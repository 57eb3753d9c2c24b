"""Checkpoint interop: `pretrained=True` follows the reference -- torch.hub.load_state_dict_from_url +
strict load_state_dict, CaiT keys carrying the DistributedDataParallel 'module.' prefix (models/cait.py:264-273), DeiT
under "model" (models/deit.py:100-105), DINO a bare state_dict (upstream hubconf). No network: the files are placed in
$TORCH_HOME/hub/checkpoints/, which torch.hub consults first; they are written from the ORACLE's modules, so the key
names, shapes and the prefix handling are checked against the reference-pinned restatement."""
import os

import pytest
import torch


@pytest.fixture()
def hub_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("TORCH_HOME", str(tmp_path))
    torch.hub.set_dir(str(tmp_path / "hub"))
    d = tmp_path / "hub" / "checkpoints"
    d.mkdir(parents=True)
    yield d
    torch.hub._hub_dir = None


def _same_state(model, sd):
    msd = model.state_dict()
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(msd[k].cpu(), sd[k]), k


def test_cait_checkpoint_with_module_prefix(hub_dir):
    from oracle import cait as ocait
    from vit_torch_b200 import cait
    torch.manual_seed(0)
    src = ocait.create("cait_XXS24_224", num_classes=1000)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    torch.save({"model": {"module." + k: v for k, v in sd.items()}}, hub_dir / "XXS24_224.pth")
    m = cait.cait_XXS24_224(pretrained=True)
    _same_state(m, sd)
    with pytest.raises(Exception):          # another size: no cached file and no network -> torch.hub's own error
        cait.cait_XXS36_224(pretrained=True)


def test_dino_and_deit_checkpoints(hub_dir):
    from oracle import vit as ovit
    from vit_torch_b200 import models
    torch.manual_seed(1)
    src = ovit.dino_vits16()
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    torch.save(sd, hub_dir / "dino_deitsmall16_pretrain.pth")
    _same_state(models.dino_vits16(pretrained=True), sd)
    src = ovit.TimmVisionTransformer(embed_dim=192, depth=12, num_heads=3, distilled=True)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    torch.save({"model": sd}, hub_dir / "deit_tiny_distilled_patch16_224-b40b3cf7.pth")
    _same_state(models.deit_tiny_distilled_patch16_224(pretrained=True), sd)
    # a checkpoint with a missing / renamed key must fail loudly (strict load), as in the reference
    bad = {("blocks.0.attn.qkv.weight_x" if k == "blocks.0.attn.qkv.weight" else k): v for k, v in sd.items()}
    torch.save({"model": bad}, hub_dir / "deit_tiny_distilled_patch16_224-b40b3cf7.pth")
    with pytest.raises(RuntimeError):
        models.deit_tiny_distilled_patch16_224(pretrained=True)


@pytest.mark.gpu
def test_pretrained_checkpoint_gives_oracle_logits(hub_dir):
    """Save an oracle CaiT with 'module.' prefixes, load it through cait_*(pretrained=True), compare logits on the
    GPU with the oracle (bf16 tolerance); same for a DINO checkpoint at 96 px (interpolated position table)."""
    from oracle import cait as ocait
    from oracle import vit as ovit
    from vit_torch_b200 import cait, models
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    src = ocait.create("cait_XXS24_224", num_classes=1000).cuda()
    with torch.no_grad():
        for b in list(src.blocks) + list(src.blocks_token_only):
            b.gamma_1.fill_(0.5); b.gamma_2.fill_(0.5)      # "trained" LayerScale: the branches matter
    torch.save({"model": {"module." + k: v.cpu() for k, v in src.state_dict().items()}}, hub_dir / "XXS24_224.pth")
    m = cait.cait_XXS24_224(pretrained=True).cuda()
    x = torch.randn(2, 3, 224, 224, device="cuda")
    with torch.no_grad():
        a, b = m(x), src(x)
    assert ((a - b).abs().max() / b.abs().max()).item() <= 2e-2
    src = ovit.dino_vits16().cuda()
    torch.save({k: v.cpu() for k, v in src.state_dict().items()}, hub_dir / "dino_deitsmall16_pretrain.pth")
    m = models.dino_vits16(pretrained=True).cuda()
    x = torch.randn(2, 3, 96, 96, device="cuda")
    with torch.no_grad():
        a, b = m(x), src(x)
    assert ((a - b).abs().max() / b.abs().max()).item() <= 2e-2
