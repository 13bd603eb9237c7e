"""Dev: time the plain NT GEMM at the linear_fuse shapes for every tile width (mdv_gemm_tune force_bn)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib(); dev = "cuda"
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    t.record(); torch.cuda.synchronize()
    return s.elapsed_time(t) / n * 1e3
for (M, N, K, f32) in ((131072, 2112, 512, 0), (131072, 512, 2112, 1)):
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    e = L.GemmEpi(); e.out, e.ldc, e.out_bf16 = L.ptr(out), N, 0 if f32 else 1
    st = L.stream()
    fn = lambda: lib.mdv_gemm_nt(L.ptr(A), K, L.ptr(W), K, M, N, K, ctypes.byref(e), st)
    for bn in (64, 96, 128, 160, 192, 224, 256):
        lib.mdv_gemm_tune(bn, 0, 0)
        us = bench(fn)
        print(f"M={M} N={N} K={K} BN={bn}: {us:.1f} us  {2*M*N*K/us/1e6:.0f} TFLOP/s", flush=True)
    lib.mdv_gemm_tune(0, 0, 0)
