"""Sustained (power-capped) GEMM throughput: run each kind back to back for ~1.5 s while sampling SM clock and power.
Reports TFLOP/s, median SM MHz / W, and the clock-normalised tensor-pipe efficiency flops / (148 SM * 8192 flop/clk * f)."""
import sys, os, json, subprocess, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops

dev = "cuda"
M = 25216
rows = []
samples = []
stop = False
def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        try:
            c, w = [float(x) for x in line.split(",")]
            samples.append((time.time(), c, w))
        except Exception:
            pass
        if stop: break
    p.terminate()
threading.Thread(target=sampler, daemon=True).start()
time.sleep(1.0)

def run(name, fn, fl, secs=1.5):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.time(); n = 0
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < secs:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    t1 = time.time()
    dt = s.elapsed_time(e) * 1e-3 / n
    win = [(c, w) for t, c, w in samples if t0 + 0.5 <= t <= t1]
    mhz = statistics.median([c for c, _ in win]) if win else float("nan")
    watt = statistics.median([w for _, w in win]) if win else float("nan")
    tf = fl / dt / 1e12
    eff = fl / dt / (148 * 8192 * mhz * 1e6) if win else float("nan")
    r = dict(kind=name, us=round(dt * 1e6, 1), tflops=round(tf, 1), sm_mhz=mhz, watts=watt, pipe_eff=round(eff, 3))
    rows.append(r); print(r, flush=True)

NB = 3
for name, N, K in [("fc1", 3072, 768), ("fc2", 768, 3072), ("proj", 768, 768)]:
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
    i = [0]
    def nxt():
        j = i[0] % NB; i[0] += 1
        return j
    fl = 2.0 * M * N * K
    def fwd():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_STORE_BF16, bias=bias, out=outs[j])
    def gelu():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_BIAS_GELU, bias=bias, out=outs[j], out2=outs2[j])
    def resid():
        j = nxt(); ops.gemm(As[j], W, epilogue=ops.EPI_RESID_F32, bias=bias, resid=res[j], out=o32[j])
    def dgelu():
        j = nxt(); ops.gemm(dys[j], W, b_mn=True, epilogue=ops.EPI_DGELU, aux=auxs[j], out=dxs[j])
    def wgrad():
        j = nxt(); ops.gemm(dys[j], As[j], a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dW)
    def cublas():
        j = nxt(); torch.mm(As[j], W.t(), out=outs[j])
    run(name + "_fwd", fwd, fl)
    run(name + "_cublas", cublas, fl)
    if name == "fc1": run(name + "_gelu", gelu, fl)
    if N == 768: run(name + "_resid", resid, fl)
    if name == "fc2": run(name + "_dgelu", dgelu, fl)
    run(name + "_wgrad", wgrad, fl)
    del As, outs, outs2, o32, res, dys, dxs, auxs
stop = True
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_gemm_sustained.json", "w"), indent=1)
