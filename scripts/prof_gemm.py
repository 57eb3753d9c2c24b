"""Launch a few GEMMs of one shape for ncu (--set full) capture:
python scripts/prof_gemm.py fwd|gelu|resid|dgrad|dgelu|wgrad N K."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops
kind = sys.argv[1] if len(sys.argv) > 1 else "fwd"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2304
K = int(sys.argv[3]) if len(sys.argv) > 3 else 768
M = 25216
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").bfloat16()
w = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
o32 = torch.empty(M, N, device="cuda")
res = torch.randn(M, N, device="cuda")
dy = torch.randn(M, N, device="cuda").bfloat16()
dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
aux = torch.randn(M, K, device="cuda").bfloat16()
dw = torch.zeros(N, K, device="cuda")
cs = torch.zeros(K, device="cuda")
for _ in range(4):
    if kind == "fwd":
        ops.gemm(a, w, epilogue=ops.EPI_STORE_BF16, bias=bias, out=out)
    elif kind == "gelu":
        ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, out=out, out2=out2)
    elif kind == "resid":
        ops.gemm(a, w, epilogue=ops.EPI_RESID_F32, bias=bias, resid=res, out=o32)
    elif kind == "dgrad":
        ops.gemm(dy, w, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=dx)
    elif kind == "dgelu":
        ops.gemm(dy, w, b_mn=True, epilogue=ops.EPI_DGELU, aux=aux, out=dx, colsum=cs)
    else:
        ops.gemm(dy, a, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dw)
torch.cuda.synchronize()
print("ok")
