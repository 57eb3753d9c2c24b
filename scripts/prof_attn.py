"""Launch attention fwd/bwd (or LayerNorm fwd/bwd with `ln`) a few times for ncu capture:
python scripts/prof_attn.py [B N H d] | python scripts/prof_attn.py ln"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops
torch.manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "ln":
    M, D = 25216, 768
    x = torch.randn(M, D, device="cuda"); w = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
    dy = torch.randn(M, D, device="cuda").bfloat16(); dres = torch.randn(M, D, device="cuda")
    dw = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
    for _ in range(3):
        h, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
        dx, dxb = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, dweight=dw, dbias=db, want_bf16=True)
else:
    B, N, H, d = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (128, 197, 12, 64)
    qkv = torch.randn(B * N, 3 * H * d, device="cuda").bfloat16()
    do = torch.randn(B * N, H * d, device="cuda").bfloat16()
    for _ in range(3):
        out, lse2 = ops.attn_fwd(qkv, B, N, H, d, d ** -0.5)
        dqkv = ops.attn_bwd(qkv, out, do, lse2, B, N, H, d, d ** -0.5)
torch.cuda.synchronize()
print("ok")
