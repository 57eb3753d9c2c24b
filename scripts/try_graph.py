"""Whole-step CUDA graph experiment: capture Trainer.step on static input buffers and compare step time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import models, train
torch.manual_seed(0)
m = models.dino_vitb16(pretrained=False).cuda()
train.reset_parameters_like_zoo(m)
tr = train.Trainer(m)
bs = 128
x = torch.randn(bs, 3, 224, 224, device="cuda"); y = torch.randint(0, 10, (bs,), device="cuda")
for _ in range(3): tr.step(x, y)
torch.cuda.synchronize()
def timeit(fn, n=10):
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
print(f"eager step {timeit(lambda: tr.step(x, y)):.2f} ms", flush=True)
g = torch.cuda.CUDAGraph()
sx, sy = x.clone(), y.clone()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): tr.step(sx, sy)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    loss = tr.step(sx, sy)
torch.cuda.synchronize()
l0 = loss.item()
print(f"graph step {timeit(g.replay):.2f} ms  loss after capture {l0:.4f} -> {loss.item():.4f}", flush=True)
