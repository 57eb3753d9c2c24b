"""BASELINE config 5: dino_vitb16 --lineareval (main.py:184-201, utils_network.py:413-418) -- frozen backbone forward under
no_grad + fused fc head (--fc 256 128 32) trained on its features, bs 512 per GPU. Prints images/s for a batch sweep
(CUDA events, 3 warm-up + 10 timed steps, batch resident in HBM) and writes gpurun_out/bench_lineareval.json.
    python scripts/bench_lineareval.py [batch ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import functional, models, train, zoo

batches = [int(v) for v in sys.argv[1:]] or [128, 256, 512]
dev = torch.device("cuda")
torch.manual_seed(0)
backbone = models.dino_vitb16(pretrained=False).to(dev)
train.reset_parameters_like_zoo(backbone)
head = zoo.get_classifier_head(768, [256, 128, 32, 10]).to(dev)
opt = train.FusedSGD(head.parameters(), lr=1e-3, momentum=0.9)
N, D, L, P = 197, 768, 12, 16
fwd_flops = 2 * (N - 1) * 3 * P * P * D + L * (2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 16 * N * D * D)
rows = []
for bs in batches:
    x = torch.randn(bs, 3, 224, 224, device=dev)
    y = torch.randint(0, 10, (bs,), device=dev)

    def step():
        with torch.no_grad():
            f = backbone(x)
        loss, _ = functional.cross_entropy(head(f), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    rows.append({"workload": "dino_vitb16 lineareval (frozen backbone fwd + fc [256,128,32,10] head train)", "batch": bs,
                 "ms_per_step": ms, "images_per_s": bs / (ms * 1e-3), "backbone_tflops": bs * fwd_flops / (ms * 1e-3) / 1e12,
                 "loss": loss.item(), "launch": "eager"})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_lineareval.json", "w"), indent=1)
