"""RESID_F32 epilogue streaming rate vs K (K=64: almost pure epilogue). Achieved bytes = A + resid + out."""
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
    return s.elapsed_time(e) / iters * 1e-3
for K in (64, 256, 768, 3072):
    As = [torch.randn(M, K, device=dev).bfloat16() for _ in range(NB)]
    W = torch.randn(N, K, device=dev).bfloat16()
    o32 = [torch.empty(M, N, device=dev) for _ in range(NB)]
    ob = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    res = [torch.randn(M, N, device=dev) for _ in range(NB)]
    bias = torch.randn(N, device=dev)
    i = [0]
    def nxt():
        j = i[0] % NB; i[0] += 1; return j
    def resid():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_RESID_F32, bias=bias, resid=res[j], out=o32[j])
    def plain():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_STORE_BF16, bias=bias, out=ob[j])
    def copy():
        j = nxt(); torch.add(res[j], 1.0, out=o32[j])
    tr, tp, tc = timeit(resid), timeit(plain), timeit(copy)
    by = M * K * 2 + 2 * M * N * 4
    print(f"K={K}: resid {tr*1e6:.1f} us ({by/tr/1e12:.2f} TB/s) plain {tp*1e6:.1f} us; torch add (same fp32 in/out bytes) {tc*1e6:.1f} us", flush=True)
