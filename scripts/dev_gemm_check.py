"""Dev check (GPU): tcgen05 GEMM vs torch.matmul.  Not a pytest file; the pytest coverage is tests/test_gpu_ops.py::test_gemm_*."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L

lib = L.lib()
torch.manual_seed(0)
dev = "cuda"

def run_nt(M, N, K, bias=True, res=False, out_bf16=True, act=0, lda=None):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    R = torch.randn(M, N, device=dev) if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    e = L.GemmEpi()
    e.bias = L.ptr(b); e.residual = L.ptr(R); e.out = L.ptr(out); e.ldc = N; e.ld_res = N
    e.out_bf16 = int(out_bf16); e.act = act
    rc = lib.mdv_gemm_nt(L.ptr(A), K, L.ptr(W), K, M, N, K, ctypes.byref(e), L.stream())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias: ref = ref + b
    if act == 1: ref = torch.nn.functional.gelu(ref)
    if res: ref = ref + R
    err = (out.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    return rc, err

def run_tn(R, P, Q):
    A = torch.randn(R, P, device=dev).bfloat16()
    B = torch.randn(R, Q, device=dev).bfloat16()
    C = torch.zeros(P, Q, device=dev)
    rc = lib.mdv_gemm_tn(L.ptr(A), P, L.ptr(B), Q, R, P, Q, L.ptr(C), Q, L.stream())
    torch.cuda.synchronize()
    ref = A.float().t() @ B.float()
    err = (C - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    return rc, err

ok = True
for (M, N, K) in [(128, 64, 64), (256, 128, 64), (1000, 192, 64), (4096, 384, 128), (300, 320, 320), (2048, 512, 2112),
                  (8, 512, 512), (8192, 1280, 320), (2048, 1024, 4608), (131072, 192, 64), (131072, 64, 512), (5000, 32, 64),
                  (3000, 288, 64), (70000, 960, 320)]:
    for kw in [dict(), dict(res=True, out_bf16=False), dict(act=1, bias=True)]:
        try:
            rc, err = run_nt(M, N, K, **kw)
        except Exception as ex:
            rc, err = -99, float("nan"); print("EXC", ex)
        flag = "OK " if (rc == 0 and err < 2e-2) else "BAD"
        ok &= flag == "OK "
        print(f"NT {flag} M={M} N={N} K={K} {kw} rc={rc} relerr={err:.3e}", flush=True)
for (R, P, Q) in [(64, 128, 64), (256, 128, 64), (4096, 192, 64), (1000, 64, 64), (8192, 512, 64), (8192, 320, 1280),
                  (131072, 192, 64), (16384, 512, 2112), (2048, 1024, 4608), (9000, 32, 64), (9000, 64, 288), (8, 512, 4608)]:
    try:
        rc, err = run_tn(R, P, Q)
    except Exception as ex:
        rc, err = -99, float("nan"); print("EXC", ex)
    flag = "OK " if (rc == 0 and err < 2e-2) else "BAD"
    ok &= flag == "OK "
    print(f"TN {flag} R={R} P={P} Q={Q} rc={rc} relerr={err:.3e}", flush=True)

# quick timing
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

for (M, N, K) in [(131072, 192, 64), (131072, 512, 64), (131072, 64, 512), (32768, 1024, 128), (8192, 1280, 320),
                  (131072, 512, 2112), (8192, 8192, 8192)]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    e = L.GemmEpi(); e.out = L.ptr(out); e.ldc = N; e.out_bf16 = 1
    st = L.stream()
    for bn, stg in [(0, 0), (64, 0), (128, 0), (256, 0), (256, 2), (128, 3)]:
        lib.mdv_gemm_tune(bn, stg, 0)
        t = bench(lambda: lib.mdv_gemm_nt(L.ptr(A), K, L.ptr(W), K, M, N, K, ctypes.byref(e), st))
        fl = 2.0 * M * N * K / t / 1e9
        by = (M * K + N * K + M * N) * 2 / t / 1e6
        print(f"time NT M={M} N={N} K={K} bn={bn} st={stg}: {t*1e3:.1f} us  {fl:.0f} TFLOP/s  {by:.0f} GB/s", flush=True)
    lib.mdv_gemm_tune(0, 0, 0)
    t = bench(lambda: torch.matmul(A, W.t()))
    print(f"   torch.matmul: {t*1e3:.1f} us {2.0*M*N*K/t/1e9:.0f} TFLOP/s")
print("ALL OK" if ok else "SOME BAD")
