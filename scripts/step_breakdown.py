"""In-step per-op GPU time (CUDA events around every libvitk launch) vs the whole step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import models, train, ops
name = sys.argv[1] if len(sys.argv) > 1 else "dino_vitb16"
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
if name.startswith("cait"):
    from vit_torch_b200 import cait
    m = getattr(cait, name)(pretrained=False, num_classes=10).cuda()
else:
    m = getattr(models, name)(pretrained=False).cuda()
    train.reset_parameters_like_zoo(m)
tr = train.Trainer(m)
x = torch.randn(bs, 3, 224, 224, device="cuda"); y = torch.randint(0, 10, (bs,), device="cuda")
for _ in range(3): tr.step(x, y)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): tr.step(x, y)
e.record(); torch.cuda.synchronize()
print(f"plain step {s.elapsed_time(e)/5:.2f} ms")
ops.op_timing_begin()
s.record()
for _ in range(5): tr.step(x, y)
e.record()
res = ops.op_timing_end()
tot = s.elapsed_time(e) / 5
acc = 0
for k, (t, n) in sorted(res.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} {t/5:8.3f} ms/step  x{n//5:4d}  avg {1e3*t/n:7.1f} us")
    acc += t / 5
print(f"sum of ops {acc:.2f} ms, instrumented step {tot:.2f} ms")
