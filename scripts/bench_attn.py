"""Event-timed attention kernels (forward, backward variants) at the BASELINE shapes; L2 flushed between iterations.
    python scripts/bench_attn.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops

def timeit(fn, iters=20):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

rows = []
for name, B, N, H, d in (("vitb16_bs128", 128, 197, 12, 64), ("vitb8_bs64", 64, 785, 12, 64), ("vits16_bs128", 128, 197, 6, 64)):
    qkv = (torch.randn((B * N, 3 * H * d), device="cuda") * 1.2).to(torch.bfloat16)
    dout = torch.randn((B * N, H * d), device="cuda").to(torch.bfloat16)
    scale = d ** -0.5
    out, lse2 = ops.attn_fwd(qkv, B, N, H, d, scale)
    dbias = torch.zeros((3 * H * d,), device="cuda")
    fl = 4.0 * N * N * H * d * B
    r = {"shape": name, "B": B, "N": N, "H": H, "d": d}
    r["fwd_us"] = timeit(lambda: ops.attn_fwd(qkv, B, N, H, d, scale))
    r["fwd_tflops"] = fl / r["fwd_us"] / 1e6
    if os.environ.get("BENCH_ATTN_FWD_ONLY") == "1":
        print(r, flush=True)
        rows.append(r)
        continue
    for variant, env in (("head", "1"), ("two_kernel", "0")):
        os.environ["VITK_ATTN_BWD_HEAD"] = env
        if variant == "head" and N > 256:
            continue
        r[f"bwd_{variant}_us"] = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse2, B, N, H, d, scale, dbias=dbias))
        r[f"bwd_{variant}_tflops"] = 2.5 * fl / r[f"bwd_{variant}_us"] / 1e6
    os.environ.pop("VITK_ATTN_BWD_HEAD", None)     # default: two-kernel
    if d == 64:     # single-kernel backward (S / dP computed once, dQ through fp32 atomics + conversion pass)
        r["bwd_fused_us"] = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse2, B, N, H, d, scale, fused=True, dbias=dbias))
        r["bwd_fused_tflops"] = 2.5 * fl / r["bwd_fused_us"] / 1e6
    print(r, flush=True)
    rows.append(r)
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_attn.json"
os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
json.dump(rows, open(path, "w"), indent=1)
