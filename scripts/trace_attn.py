"""Per-CTA event timeline of the attention kernels (instrumented build: VITK_NVCC_EXTRA=-DVITK_TRACE python -m
vit_torch_b200.build, run with VITK_LIB=vit_torch_b200/libvitk_dbg.so). Prints mean / median SM-clock intervals between
the events the kernels record (see VITK_TRACE_EV in csrc/attn_fwd.cu, attn_bwd.cu):
  0 CTA start (TMA warp)   1 MMA warp past the TMEM-alloc sync   2 first S issued   3 (bwd dq) delta computed
  8+j  accumulate-MMA of tile j issued       24+2j tile j scores visible to elementwise warp 0
  25+2j elementwise warp 0 finished tile j   60 final accumulator visible   61 CTA write-out done
usage: python scripts/trace_attn.py [fwd|dq|dkv] [B N H d]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_torch_b200 import ops, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, N, H, d = (int(v) for v in sys.argv[2:6]) if len(sys.argv) > 5 else (128, 197, 12, 64)
lib = _lib.load()
lib.vitk_debug_set_trace.argtypes = [ctypes.c_void_p]
ncta = ((N + 127) // 128) * H * B
trace = torch.zeros(ncta * 64, dtype=torch.int64, device="cuda")
qkv = torch.randn(B * N, 3 * H * d, device="cuda").bfloat16()
do = torch.randn(B * N, H * d, device="cuda").bfloat16()
out, lse2 = ops.attn_fwd(qkv, B, N, H, d, d ** -0.5)
if which != "fwd":
    os.environ["VITK_ATTN_DBG_SKIP"] = "2" if which == "dq" else "1"
    if which == "dkv":   # delta comes from the dq kernel: produce it once
        os.environ["VITK_ATTN_DBG_SKIP"] = "0"
        ops.attn_bwd(qkv, out, do, lse2, B, N, H, d, d ** -0.5)
        os.environ["VITK_ATTN_DBG_SKIP"] = "1"
run = (lambda: ops.attn_fwd(qkv, B, N, H, d, d ** -0.5)) if which == "fwd" else \
      (lambda: ops.attn_bwd(qkv, out, do, lse2, B, N, H, d, d ** -0.5))
for _ in range(3): run()
torch.cuda.synchronize()
lib.vitk_debug_set_trace(trace.data_ptr())
trace.zero_()
run()
torch.cuda.synchronize()
lib.vitk_debug_set_trace(None)
t = trace.view(ncta, 64).cpu().double()
nt = min(8, (N + 63) // 64)
t0 = t[:, 0]
span = (t[:, 61].max() - t0.min()).item()   # (different SMs: clocks are only roughly aligned)
def stat(name, a, b):
    m = (t[:, a] > 0) & (t[:, b] > 0)
    if int(m.sum()) == 0:
        return
    dlt = (t[m, b] - t[m, a])
    print(f"  {name:44s} mean {dlt.mean().item():8.0f}  median {dlt.median().item():8.0f}  p90 {dlt.quantile(0.9).item():8.0f}  (n={int(m.sum())})")
print(f"{which} B{B} N{N} H{H} d{d}: {ncta} CTAs, clocks (SM cycles)")
stat("CTA lifetime 0 -> 61", 0, 61)
stat("start -> MMA warp past alloc sync (0->1)", 0, 1)
stat("-> first S issued (1->2)", 1, 2)
stat("start -> first scores visible (0->24)", 0, 24)
for j in range(nt):
    stat(f"tile {j}: elementwise (24+2j -> 25+2j)", 24 + 2 * j, 25 + 2 * j)
    if which == "fwd" and j == 1:     # skew between the eight softmax warps at the end of tile 1 (warp 0 = reference)
        for w in range(1, 8):
            stat(f"tile 1:   warp {w} done - warp 0 done", 27, 48 + w)
    if which == "fwd" and j < 2:      # finer split of the forward's elementwise pass (warp 0) and the other group's span
        stat(f"tile {j}:   scores visible -> loaded, S released", 24 + 2 * j, 40 + 4 * j)
        stat(f"tile {j}:   -> max / rescale / P buffer free", 40 + 4 * j, 41 + 4 * j)
        stat(f"tile {j}:   -> exp, sums, pack, st.shared", 41 + 4 * j, 42 + 4 * j)
        stat(f"tile {j}:   -> fence.proxy.async", 42 + 4 * j, 43 + 4 * j)
        stat(f"tile {j}:   group 1 (warp 4) whole tile", 16 + 2 * j, 17 + 2 * j)
        stat(f"tile {j}:   group 1 start - group 0 start", 24 + 2 * j, 16 + 2 * j)
        stat(f"tile {j}:   warp 0 done -> MMA warp sees P of group 0", 25 + 2 * j, 56 + j)
        stat(f"tile {j}:   warp 0 done -> MMA warp sees P of group 1", 25 + 2 * j, 3 + j)
    if j + 1 < nt: stat(f"tile {j}: done -> next scores visible", 25 + 2 * j, 26 + 2 * j)
    stat(f"tile {j}: elementwise done -> acc MMA issued", 25 + 2 * j, 8 + j)
stat("last tile done -> accumulator visible (->60)", 25 + 2 * (nt - 1), 60)
stat("write-out (60->61)", 60, 61)
