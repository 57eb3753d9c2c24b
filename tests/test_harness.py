"""The reference's main.py (fine-tune and --lineareval, main.py:104-282) executed UNCHANGED through
vit_torch_b200.compat.harness. Needs /root/reference (build container only).

CPU: the harness itself (import stubs, synthetic datasets under the reference's own Datasets class, hub directory
pinning, Stats / LambdaLR / early-stop plumbing) with a test-local hubconf that serves the ORACLE's DINO on the CPU --
the product models have no CPU path. GPU: the real shims, i.e. the fused sm_100a models and head."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.reference
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

_ORACLE_HUBCONF = '''
import sys
sys.path.insert(0, {root!r})
dependencies = ["torch"]
from oracle.vit import dino_vits16, dino_vits8, dino_vitb16, dino_vitb8  # noqa: E402,F401
'''

_DRIVER = '''
import sys
sys.path.insert(0, {root!r})
from vit_torch_b200.compat import harness
harness.run_main({ref!r}, {argv!r}, torch_home={home!r}, fused_head={fused!r}, hub_shim={shim!r})
'''


def _run(tmp_path, argv, fused, shim, timeout=900):
    home = str(tmp_path / "home")
    if not shim:     # serve the oracle's DINO from a test-local hub directory (CPU)
        d = os.path.join(home, "hub", "facebookresearch_dino_main")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "hubconf.py"), "w") as f:
            f.write(_ORACLE_HUBCONF.format(root=ROOT))
    stats = str(tmp_path / "stats.json")
    argv = argv + ["--root_path", home, "--stats_fp", stats]
    env = dict(os.environ, VITK_SYNTH_SAMPLES="64", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-c", _DRIVER.format(root=ROOT, ref=REF, argv=argv, home=home, fused=fused,
                                                             shim=shim)],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    with open(stats) as f:
        return json.load(f), p.stdout


def _check_stats(st, epochs):
    assert st["info"]["arch"].startswith("dino_")
    assert len(st["train"]) == epochs and len(st["val"]) == epochs
    for r in st["train"] + st["val"]:
        assert r["loss"] == r["loss"] and r["loss"] > 0          # finite
        assert 0.0 <= r["acc"] <= 1.0


def test_import_stubs_and_synthetic_datasets(tmp_path):
    """The reference's Datasets class (utils_datasets.py:758-907) on the synthetic STL10: attributes main.py and
    Network read (loaders, info, num_labels; utils_network.py:173-178)."""
    code = f'''
import sys
sys.path.insert(0, {ROOT!r})
from vit_torch_b200.compat import harness
stubbed = harness.install_import_stubs()
harness.install_synthetic_datasets()
sys.path.insert(0, {REF!r})
from utils_datasets import Datasets
ds = Datasets(dataset="stl10", image_size=32, root_path="/tmp", batchsize=8, splits=["train", "test"], num_workers=0,
              limit_train=24, limit_test=16)
x, y = next(iter(ds.loaders["train"]))
assert tuple(x.shape) == (8, 3, 32, 32) and x.dtype.is_floating_point and y.shape == (8,)
assert ds.num_labels == 10 and ds.info["sample_count"] == {{"train": 24, "test": 16}}
assert ds.info["batch_count"] == {{"train": 3, "test": 2}}
assert abs(float(x.mean())) < 3 and 0.2 < float(x.std()) < 3       # Normalize (utils_datasets.py:578-580) applied
print("STUBBED", stubbed)
'''
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, VITK_SYNTH_SAMPLES="64"))
    assert p.returncode == 0, p.stderr[-3000:]
    assert "numpy.lib.arraysetops" in p.stdout


def test_reference_main_finetune_and_lineareval_cpu(tmp_path):
    st, _ = _run(tmp_path / "ft", ["--arch", "dino_vits16", "--fc", "32", "16", "--device", "cpu", "--bs", "8",
                                   "--image_size", "32", "--epoch", "2", "--limit_train", "16", "--limit_test", "8",
                                   "--lr", "0.01"], fused=False, shim=False)
    _check_stats(st, 2)
    assert st["telem"]["mode"] == "finetune"
    st, _ = _run(tmp_path / "le", ["--arch", "dino_vits16", "--fc", "32", "16", "--device", "cpu", "--bs", "8",
                                   "--image_size", "32", "--epoch", "1", "--limit_train", "16", "--limit_test", "8",
                                   "--lineareval"], fused=False, shim=False)
    _check_stats(st, 1)
    assert st["telem"]["mode"] == "lineareval"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["finetune", "lineareval", "cait"])
def test_reference_main_on_fused_models_gpu(tmp_path, mode):
    """main.py --arch dino_vits16 --fc 256 128 32 (BASELINE config 1's command line) and --lineareval, plus a CaiT
    fine-tune, on cuda through the real hub / timm shims and the fused head."""
    argv = ["--fc", "256", "128", "32", "--device", "cuda", "--bs", "16", "--epoch", "2", "--limit_train", "48",
            "--limit_test", "16", "--lr", "0.01"]
    if mode == "cait":
        argv = ["--arch", "cait_XXS24_224", "--image_size", "224"] + argv
    else:
        argv = ["--arch", "dino_vits16", "--image_size", "96"] + argv + (["--lineareval"] if mode == "lineareval" else [])
    st, out = _run(tmp_path, argv, fused=True, shim=True, timeout=1500)
    _check_stats(st, 2) if mode != "cait" else None
    assert st["telem"]["mode"] == ("lineareval" if mode == "lineareval" else "finetune")
    assert len(st["train"]) == 2
    assert st["train"][-1]["loss"] < st["train"][0]["loss"] * 1.05      # it trains (no divergence / NaN)
