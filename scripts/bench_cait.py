"""BASELINE config 4: cait_S24_224 (class attention + talking heads + LayerScale) bf16 fine-tune step, bs 128 on one GPU
(the 8-GPU line needs only the data-parallel wrapper of bench.py). CUDA events, 3 warm-up (2 eager + capture) + 10 timed
steps through train.Trainer(graph=True); writes gpurun_out/bench_cait.json.   python scripts/bench_cait.py [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import cait, train

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
torch.manual_seed(0)
model = cait.cait_S24_224(pretrained=False, num_classes=10).to(dev)
rows = []
for graph in (False, True):
    tr = train.Trainer(model, lr=1e-3, momentum=0.9, graph=graph)
    x = torch.randn(bs, 3, 224, 224, device=dev)
    y = torch.randint(0, 10, (bs,), device=dev)
    for _ in range(3):
        tr.step(x, y)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        loss = tr.step(x, y)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    # 18.654 GFLOP forward per image, train step = 3x
    rows.append({"workload": "cait_S24_224 fine-tune (fwd+CE+bwd+SGD momentum 0.9) 224x224 synthetic, random init",
                 "batch": bs, "launch": "CUDA graph" if graph else "eager", "ms_per_step": ms,
                 "images_per_s": bs / (ms * 1e-3), "step_tflops": bs * 3 * 18.654e9 / (ms * 1e-3) / 1e12,
                 "loss": loss.item()})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_cait.json", "w"), indent=1)
