import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops
dev = "cuda"; M = 25216; N = 768; NB = 4
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e6
out = []
for K in (64, 768):
    As = [torch.randn(M, K, device=dev).bfloat16() for _ in range(NB)]
    W = torch.randn(N, K, device=dev).bfloat16()
    o32 = [torch.empty(M, N, device=dev) for _ in range(NB)]
    res = [torch.randn(M, N, device=dev) for _ in range(NB)]
    bias = torch.randn(N, device=dev)
    i = [0]
    def resid():
        j = i[0] % NB; i[0] += 1
        ops.gemm(As[j], W, epilogue=ops.EPI_RESID_F32, bias=bias, resid=res[j], out=o32[j])
    def f32():
        j = i[0] % NB; i[0] += 1
        ops.gemm(As[j], W, epilogue=ops.EPI_STORE_F32, bias=bias, out=o32[j])
    out.append(f"K={K}: resid {timeit(resid):.1f} store_f32 {timeit(f32):.1f}")
print("dbg", os.environ.get("VITK_GEMM_DBG", "0"), "|", " | ".join(out), flush=True)
