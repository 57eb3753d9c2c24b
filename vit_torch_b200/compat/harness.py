"""Run the reference's own `main.py` (fine-tune and --lineareval, main.py:104-282) UNCHANGED on the fused models, in
an image without network access, datasets or the three packages its imports expect.

    python -m vit_torch_b200.compat.harness /path/to/ViT_torch -- --arch dino_vits16 --fc 256 128 32 --bs 128 \
        --image_size 224 --epoch 1 --root_path /tmp/vitk_home

What `prepare()` installs, all of it outside the reference tree (nothing there is edited):
  * import stubs for modules the reference imports but never needs on this path: `numpy.lib.arraysetops`
    (utils_datasets.py:3, removed in numpy 2), `skimage.feature` (utils_datasets.py:17, tire dataset only),
    `adabelief_pytorch` (utils_network.py:21, `--opt adabelief` only) -- each raises when actually used;
  * the `timm` shim and the torch.hub shim (vit_torch_b200.compat): `torch.hub.load('facebookresearch/dino:main', arch)`
    (models/vision_all.py:156) and `timm.models.factory.create_model` (:186) hand the zoo the fused sm_100a models.
    torch.hub's directory is pinned with torch.hub.set_dir(), because VisionModelZoo.get_model resets TORCH_HOME to its
    default '/host/ubuntu/torch' when main.py's fine-tune branch omits root_path (main.py:203-208 vs vision_all.py:88-89);
  * synthetic stand-ins for torchvision.datasets.{STL10,CIFAR10,CIFAR100} (class-conditional images of the right size,
    generated from the sample index): the reference's own `Datasets` class (utils_datasets.py:758-907) -- transforms,
    Subset limits, DataLoaders, `loaders` / `info` / `num_labels` -- runs as written on top of them;
  * VisionModelZoo.get_classifier_head -> the fused head (zoo.patch_reference_zoo), so that CaiT / DeiT forwards and
    the lineareval head train on tcgen05 GEMMs instead of a stock nn.Sequential.
"""
from __future__ import annotations

import os
import runpy
import sys
import types

import numpy as np

from . import install_hub_shim, install_timm_shim


# ------------------------------------------------------------------------------------------------------------------
# import stubs
# ------------------------------------------------------------------------------------------------------------------
def _unavailable(what):
    def fn(*a, **k):
        raise ImportError(f"{what} is not installed in this image (vit_torch_b200.compat.harness provides an import "
                          "stub only; it is not needed on the ViT fine-tune / lineareval path)")
    return fn


def install_import_stubs() -> list[str]:
    """Stubs for `numpy.lib.arraysetops`, `skimage.feature`, `adabelief_pytorch` when they cannot be imported.
    Returns the names that were stubbed."""
    import importlib.util
    made = []

    def missing(name):
        if name in sys.modules:
            return False
        try:
            return importlib.util.find_spec(name) is None
        except (ImportError, ValueError, AttributeError):
            return True

    if missing("numpy.lib.arraysetops"):
        m = types.ModuleType("numpy.lib.arraysetops")
        for nm in ("isin", "unique", "in1d", "intersect1d", "union1d", "setdiff1d", "setxor1d"):
            if hasattr(np, nm):
                setattr(m, nm, getattr(np, nm))
        sys.modules["numpy.lib.arraysetops"] = m
        made.append("numpy.lib.arraysetops")
    if missing("skimage"):
        sk = types.ModuleType("skimage")
        ft = types.ModuleType("skimage.feature")
        ft.local_binary_pattern = _unavailable("scikit-image (skimage.feature.local_binary_pattern)")
        sk.feature = ft
        sys.modules["skimage"], sys.modules["skimage.feature"] = sk, ft
        made.append("skimage")
    if missing("adabelief_pytorch"):
        ab = types.ModuleType("adabelief_pytorch")

        class AdaBelief:  # noqa: D401 - stand-in
            def __init__(self, *a, **k):
                _unavailable("adabelief_pytorch.AdaBelief")()

        ab.AdaBelief = AdaBelief
        sys.modules["adabelief_pytorch"] = ab
        made.append("adabelief_pytorch")
    return made


# ------------------------------------------------------------------------------------------------------------------
# synthetic torchvision datasets
# ------------------------------------------------------------------------------------------------------------------
class _SyntheticImages:
    """Class-conditional synthetic images: a fixed low-frequency template per class plus per-sample noise, generated
    from (seed, index) on demand -- learnable (the loss falls), deterministic, no files. Returns (PIL image, label) and
    applies `transform` / `target_transform` like a torchvision VisionDataset."""

    num_classes = 10
    size = 32
    counts = {True: 50000, False: 10000}

    def __init__(self, train: bool, transform=None, target_transform=None, seed: int = 0):
        self.train, self.transform, self.target_transform = train, transform, target_transform
        n = int(os.environ.get("VITK_SYNTH_SAMPLES", "0")) or self.counts[train]
        self._n = n
        self._seed = seed + (0 if train else 7919)
        rng = np.random.default_rng(1234)           # templates are shared by the train and test splits
        low = rng.uniform(40, 215, size=(self.num_classes, 4, 4, 3))
        self._templates = np.stack([np.kron(t, np.ones((self.size // 4, self.size // 4, 1))) for t in low])
        self.classes = [str(i) for i in range(self.num_classes)]

    def __len__(self):
        return self._n

    def __getitem__(self, index):
        from PIL import Image
        index = int(index)
        if index < 0 or index >= self._n:
            raise IndexError(index)
        label = index % self.num_classes
        rng = np.random.default_rng((self._seed, index))
        img = self._templates[label] + rng.normal(0.0, 25.0, size=(self.size, self.size, 3))
        img = Image.fromarray(np.clip(img, 0, 255).astype(np.uint8))
        if self.transform is not None:
            img = self.transform(img)
        if self.target_transform is not None:
            label = self.target_transform(label)
        return img, label


class SyntheticSTL10(_SyntheticImages):
    """torchvision.datasets.STL10(root, split, folds, transform, target_transform, download): 96x96, 10 classes."""
    size = 96
    counts = {True: 5000, False: 8000}

    def __init__(self, root=None, split="train", folds=None, transform=None, target_transform=None, download=False):
        super().__init__(split == "train", transform, target_transform)
        self.root, self.split = root, split


class SyntheticCIFAR10(_SyntheticImages):
    """torchvision.datasets.CIFAR10(root, train, transform, target_transform, download): 32x32, 10 classes."""

    def __init__(self, root=None, train=True, transform=None, target_transform=None, download=False):
        super().__init__(bool(train), transform, target_transform)
        self.root = root


class SyntheticCIFAR100(SyntheticCIFAR10):
    num_classes = 100


def install_synthetic_datasets():
    """torchvision.datasets.{STL10,CIFAR10,CIFAR100} -> synthetic stand-ins (the reference constructs them with
    download=True, utils_datasets.py:625-636,683-694,740-752). Returns a function that restores the originals."""
    import torchvision
    saved = {k: getattr(torchvision.datasets, k) for k in ("STL10", "CIFAR10", "CIFAR100")}
    torchvision.datasets.STL10 = SyntheticSTL10
    torchvision.datasets.CIFAR10 = SyntheticCIFAR10
    torchvision.datasets.CIFAR100 = SyntheticCIFAR100

    def restore():
        for k, v in saved.items():
            setattr(torchvision.datasets, k, v)
    return restore


# ------------------------------------------------------------------------------------------------------------------
# putting it together
# ------------------------------------------------------------------------------------------------------------------
def prepare(reference_root: str, torch_home: str, synthetic_data: bool = True, fused_head: bool = True,
            hub_shim: bool = True):
    """Make `import main` / `runpy.run_path(main.py)` of the reference work here. Returns the reference's
    VisionModelZoo class (already patched when fused_head)."""
    import torch
    install_import_stubs()
    install_timm_shim()
    os.makedirs(torch_home, exist_ok=True)
    if hub_shim:
        install_hub_shim(torch_home)
    torch.hub.set_dir(os.path.join(torch_home, "hub"))
    if synthetic_data:
        install_synthetic_datasets()
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from models.vision_all import VisionModelZoo       # the reference's own zoo
    if fused_head:
        from .. import zoo
        zoo.patch_reference_zoo(VisionModelZoo)
    return VisionModelZoo


def run_main(reference_root: str, argv: list[str], torch_home: str | None = None, **prepare_kw) -> None:
    """Execute the reference's main.py as __main__ with `argv` (its own command line, main.py:73-101)."""
    reference_root = os.path.abspath(reference_root)
    if torch_home is None:
        torch_home = "/tmp/vitk_home"
        if "--root_path" in argv:
            torch_home = argv[argv.index("--root_path") + 1]
    prepare(reference_root, torch_home, **prepare_kw)
    saved_argv, saved_cwd = sys.argv, os.getcwd()
    sys.argv = [os.path.join(reference_root, "main.py")] + list(argv)
    try:
        runpy.run_path(os.path.join(reference_root, "main.py"), run_name="__main__")
    finally:
        sys.argv = saved_argv
        os.chdir(saved_cwd)


def _cli():
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help"):
        print(__doc__)
        raise SystemExit(0)
    ref = sys.argv[1]
    rest = sys.argv[2:]
    if rest and rest[0] == "--":
        rest = rest[1:]
    run_main(ref, rest)


if __name__ == "__main__":
    _cli()
