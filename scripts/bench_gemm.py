"""Time the tcgen05 GEMM on the ViT-B/16 bs128 shapes with every epilogue the Block uses (CUDA events, rotating
buffers > L2) and compare with cuBLAS.  python scripts/bench_gemm.py [--quick]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops

torch.manual_seed(0)
dev = "cuda"
PEAK = 1621.6e12
quick = "--quick" in sys.argv

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3

M = int(os.environ.get("M", 25216))
rows = []
NB = 4  # rotate operand sets so that inputs do not stay L2-resident
for name, N, K in [("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]:
    As = [torch.randn(M, K, device=dev).bfloat16() for _ in range(NB)]
    W = torch.randn(N, K, device=dev).bfloat16()
    outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    outs2 = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    o32 = [torch.empty(M, N, device=dev) for _ in range(NB)] if N == 768 else None
    res = [torch.randn(M, N, device=dev) for _ in range(NB)] if N == 768 else None
    dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(NB)]
    dxs = [torch.empty(M, K, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    auxs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(NB)] if name == "fc2" else None
    dW = torch.zeros(N, K, device=dev)
    bias = torch.randn(N, device=dev)
    csum = torch.zeros(K, device=dev)
    i = [0]
    def nxt():
        j = i[0] % NB; i[0] += 1
        return j
    def fwd():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_STORE_BF16, bias=bias, out=outs[j])
    def fwd_gelu():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_BIAS_GELU, bias=bias, out=outs[j], out2=outs2[j])
    def fwd_resid():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_RESID_F32, bias=bias, resid=res[j], out=o32[j])
    def dgrad():
        j = nxt(); ops.gemm(dys[j], W, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=dxs[j])
    def dgrad_dgelu():
        j = nxt(); ops.gemm(dys[j], W, b_mn=True, epilogue=ops.EPI_DGELU, aux=auxs[j], out=dxs[j], colsum=csum)
    def wgrad():
        j = nxt(); ops.gemm(dys[j], As[j], a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dW)
    def cublas():
        j = nxt(); torch.addmm(bias.bfloat16(), As[j], W.t(), out=outs[j])
    def cublas_wgrad():
        j = nxt(); torch.mm(dys[j].t(), As[j])
    fl = 2.0 * M * N * K
    cases = [("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)]
    if name == "fc1": cases.append(("fwd_gelu", fwd_gelu))
    if N == 768: cases.append(("fwd_resid", fwd_resid))
    if name == "fc2": cases.append(("dgrad_dgelu", dgrad_dgelu))
    if not quick: cases += [("cublas_fwd", cublas), ("cublas_wgrad", cublas_wgrad)]
    for tag, fn in cases:
        t = timeit(fn)
        rows.append(dict(op=name, kind=tag, M=M, N=N, K=K, us=round(t * 1e6, 1), tflops=round(fl / t / 1e12, 1), frac=round(fl / t / PEAK, 3)))
        print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_gemm.json", "w"), indent=1)
