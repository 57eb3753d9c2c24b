"""Time fused attention fwd/bwd (CUDA events) on the BASELINE shapes."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
import torch
from vit_torch_b200 import ops
print("import", round(time.time() - t0, 2), flush=True)
dev = "cuda"

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3

rows = []
tag = "v" + os.environ.get("VITK_ATTN_WG2", "1")
for (B, N, H, d) in [(2, 197, 6, 64), (128, 197, 12, 64), (64, 785, 12, 64), (128, 196, 8, 48)]:
    qkv = torch.randn(B * N, 3 * H * d, device=dev).bfloat16()
    do = torch.randn(B * N, H * d, device=dev).bfloat16()
    t1 = time.time()
    out, lse2 = ops.attn_fwd(qkv, B, N, H, d, d ** -0.5)
    torch.cuda.synchronize()
    print("first fwd call wall", round(time.time() - t1, 3), flush=True)
    t = timeit(lambda: ops.attn_fwd(qkv, B, N, H, d, d ** -0.5))
    fl = 4.0 * B * H * N * N * d
    rows.append(dict(op="attn_fwd", B=B, N=N, H=H, d=d, us=round(t * 1e6, 1), tflops=round(fl / t / 1e12, 1)))
    print(rows[-1], flush=True)
    if hasattr(ops, "attn_bwd"):
        t1 = time.time()
        dqkv = ops.attn_bwd(qkv, out, do, lse2, B, N, H, d, d ** -0.5)
        torch.cuda.synchronize()
        print("first bwd call wall", round(time.time() - t1, 3), flush=True)
        t = timeit(lambda: ops.attn_bwd(qkv, out, do, lse2, B, N, H, d, d ** -0.5))
        rows.append(dict(op="attn_bwd", B=B, N=N, H=H, d=d, us=round(t * 1e6, 1), tflops=round(2.5 * fl / t / 1e12, 1)))
        print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/bench_attn_{tag}.json", "w"), indent=1)
