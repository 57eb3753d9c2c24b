"""Launch attention fwd/bwd a few times for ncu capture: python scripts/prof_attn.py [B N H d]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops
B, N, H, d = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (128, 197, 12, 64)
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * H * d, device="cuda").bfloat16()
do = torch.randn(B * N, H * d, device="cuda").bfloat16()
for _ in range(3):
    out, lse2 = ops.attn_fwd(qkv, B, N, H, d, d ** -0.5)
    dqkv = ops.attn_bwd(qkv, out, do, lse2, B, N, H, d, d ** -0.5)
torch.cuda.synchronize()
print("ok")
