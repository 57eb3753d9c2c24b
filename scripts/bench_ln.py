"""LayerNorm fwd/bwd achieved HBM bandwidth at the ViT-B/16 bs128 shape, or `python scripts/bench_ln.py M D`
(env VITK_LN_RING=0: register kernel; VITK_LN_RING_ROWS=r: rows per ring stage)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops
M, D, NB = 25216, 768, 4
if len(sys.argv) > 2:
    M, D = int(sys.argv[1]), int(sys.argv[2])
xs = [torch.randn(M, D, device="cuda") for _ in range(NB)]
dys = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(NB)]
drs = [torch.randn(M, D, device="cuda") for _ in range(NB)]
w = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
dw = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda"); ds = torch.zeros(D, device="cuda")
_, mean, rstd = ops.layernorm_fwd(xs[0], w, b, 1e-6)
def timeit(fn, iters=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3
i = [0]
def fwd():
    j = i[0] % NB; i[0] += 1; ops.layernorm_fwd(xs[j], w, b, 1e-6)
def bwd():
    j = i[0] % NB; i[0] += 1
    ops.layernorm_bwd(dys[j], xs[j], w, mean, rstd, dres=drs[j], dweight=dw, dbias=db, want_bf16=True, dxsum=ds)
tf, tb = timeit(fwd), timeit(bwd)
print(f"M={M} D={D} ring={os.environ.get('VITK_LN_RING', '1')} rows/stage={os.environ.get('VITK_LN_RING_ROWS', 'auto')} ln_fwd {tf*1e6:.1f} us {M*D*6/tf/1e12:.2f} TB/s | ln_bwd {tb*1e6:.1f} us {M*D*16/tb/1e12:.2f} TB/s")
