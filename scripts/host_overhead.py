"""How long does the host need to enqueue one train step (no syncs), versus the GPU time of the step?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import models, train
torch.manual_seed(0)
m = models.dino_vitb16(pretrained=False).cuda()
train.reset_parameters_like_zoo(m)
tr = train.Trainer(m)
x = torch.randn(128, 3, 224, 224, device="cuda"); y = torch.randint(0, 10, (128,), device="cuda")
for _ in range(3): tr.step(x, y)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10): tr.step(x, y)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step", flush=True)
# per-phase GPU time
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
import torch.nn.functional as F
for rep in range(2):
    e0 = ev(); out = m(x); e1 = ev(); loss = F.cross_entropy(out, y); tr.opt.zero_grad(set_to_none=True); loss.backward(); e2 = ev(); tr.opt.step(); e3 = ev()
    torch.cuda.synchronize()
    print(f"fwd {e0.elapsed_time(e1):.2f} bwd {e1.elapsed_time(e2):.2f} opt {e2.elapsed_time(e3):.2f} ms")
