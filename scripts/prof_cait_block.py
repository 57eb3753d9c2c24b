"""One CaiT-S24 LayerScale block (talking-heads attention) forward + backward at bs128 for ncu capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import cait
torch.manual_seed(0)
blk = cait.LayerScale_Block(384, 8, qkv_bias=True, init_values=1e-5).cuda()
x = torch.randn(128, 196, 384, device="cuda", requires_grad=True)
for _ in range(2):
    y = blk(x)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok")
