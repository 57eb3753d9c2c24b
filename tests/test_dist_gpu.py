"""NCCL data parallelism on real GPUs (needs >= 2): the gradients every rank holds after Trainer(graph=True) with the
deferred arena all-reduce equal the single-GPU gradients of the concatenated batch; same for the overlapped (eager)
mode, the split-graph mode, the bf16 wire format (looser tolerance) and an arena registered with the communicator."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("mode", ["deferred", "overlap", "bf16", "split", "registered"])
def test_nccl_gradients_match_single_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = 29600 + {"deferred": 1, "overlap": 2, "bf16": 3, "split": 4, "registered": 5}[mode]
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(HERE, "dist_parity_worker.py"), mode], capture_output=True, text=True, timeout=900)
    print(p.stdout[-2000:])
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    assert "dist parity" in p.stdout
