"""Epilogue experiments on proj-resid / fc2-dgelu / fc1-gelu (env VITK_GEMM_DBG, VITK_GEMM_2CTA select variants)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops
dev = "cuda"; M = 25216; NB = 4
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
out = {}
for name, N, K in [("proj", 768, 768), ("fc2", 768, 3072)]:
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
    def resid_nores():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_RESID_F32, bias=bias, out=o32[j])
    def store_bf16():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_STORE_BF16, bias=bias, out=ob[j])
    out[name + "_resid"] = timeit(resid); out[name + "_resid_noload"] = timeit(resid_nores); out[name + "_bf16"] = timeit(store_bf16)
As = [torch.randn(M, 768, device=dev).bfloat16() for _ in range(NB)]
W = torch.randn(3072, 768, device=dev).bfloat16()
o1 = [torch.empty(M, 3072, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
o2 = [torch.empty(M, 3072, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
dx = [torch.empty(M, 768, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
bias = torch.randn(3072, device=dev)
i = [0]
def nxt():
    j = i[0] % NB; i[0] += 1; return j
def gelu():
    j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_BIAS_GELU, bias=bias, out=o1[j], out2=o2[j])
def gelu_nograd():
    j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_BIAS_GELU, bias=bias, out=None, out2=o2[j])
def plain():
    j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_STORE_BF16, bias=bias, out=o2[j])
W2 = torch.randn(768, 3072, device=dev).bfloat16()
dy = [torch.randn(M, 768, device=dev).bfloat16() for _ in range(NB)]
def dgelu():
    j = nxt(); ops.gemm(dy[j], W2, b_mn=True, epilogue=ops.EPI_DGELU, aux=o1[j], out=o2[j])
def dplain():
    j = nxt(); ops.gemm(dy[j], W2, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=o2[j])
out["fc1_gelu"] = timeit(gelu); out["fc1_gelu_nograd"] = timeit(gelu_nograd); out["fc1_plain"] = timeit(plain)
out["fc2_dgelu"] = timeit(dgelu); out["fc2_dgrad_plain"] = timeit(dplain)
print(os.environ.get("VITK_GEMM_DBG", "0"), os.environ.get("VITK_GEMM_2CTA", "1"), {k: round(v, 1) for k, v in out.items()}, flush=True)
