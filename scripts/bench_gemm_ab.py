"""A/B two builds of libvitk on the same box: VITK_LIB selects the library; only vitk_gemm_bf16 (old ABI) is used."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
libs = {}
for tag, path in [("new", "vit_torch_b200/libvitk.so"), ("old", "vit_torch_b200/libvitk_old.so")]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.vitk_gemm_bf16.argtypes = [P, L, I, P, L, I, I, I, I, I, P, P, P, L, P, L, P, L, P, L, I, P]
    lib.vitk_gemm_bf16.restype = I
    libs[tag] = lib
M = 25216
def run(lib, a, lda, amn, b, ldb, bmn, Mm, N, K, epi, bias, out, ldo):
    rc = lib.vitk_gemm_bf16(a.data_ptr(), lda, amn, b.data_ptr(), ldb, bmn, Mm, N, K, epi, bias.data_ptr() if bias is not None else None, None, None, 0, out.data_ptr(), ldo, None, 0, None, 0, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
for name, N, K in [("qkv", 2304, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); dy = torch.randn(M, N, device="cuda").bfloat16()
    dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16); dW = torch.zeros(N, K, device="cuda"); bias = torch.randn(N, device="cuda")
    for rep in range(2):
        for tag in ("old", "new"):
            lib = libs[tag]
            tf = timeit(lambda: run(lib, A, K, 0, W, K, 0, M, N, K, 0, bias, out, N))
            td = timeit(lambda: run(lib, dy, N, 0, W, K, 1, M, K, N, 0, None, dx, K))
            tw = timeit(lambda: run(lib, dy, N, 1, A, K, 1, N, K, M, 4, None, dW, K))
            print(f"{name} {tag}: fwd {tf:.1f} dgrad {td:.1f} wgrad {tw:.1f}", flush=True)
