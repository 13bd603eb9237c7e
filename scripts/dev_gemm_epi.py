"""Dev: time / profile the GEMM with the MLP epilogues (fc1 forward, fc2 dgrad) at the stage-0/1 shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib()
dev = "cuda"
torch.manual_seed(0)
rng = torch.tensor([1, 2], dtype=torch.int64, device=dev)

def make(M, N, K, kind):
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    e = L.GemmEpi()
    keep = [A, W, bias]
    if kind == "plain":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); e.out, e.ldc, e.out_bf16 = L.ptr(out), N, 1
    elif kind == "fc1":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(out)
        e.out, e.ldc, e.out_bf16, e.bias, e.act, e.out_preact, e.ld_preact = L.ptr(out), N, 1, L.ptr(bias), 1, L.ptr(pre), N
        e.dropout_p, e.rng, e.drop_stream = 0.1, L.ptr(rng), 3
        keep += [pre]
    elif kind == "fc1_nodrop":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(out)
        e.out, e.ldc, e.out_bf16, e.bias, e.act, e.out_preact, e.ld_preact = L.ptr(out), N, 1, L.ptr(bias), 1, L.ptr(pre), N
        keep += [pre]
    elif kind == "fc2d":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); u = torch.randn(M, N, device=dev).bfloat16(); cs = torch.zeros(N, device=dev)
        e.out, e.ldc, e.out_bf16, e.mul_gelu_grad, e.ld_mul, e.colsum = L.ptr(out), N, 1, L.ptr(u), N, L.ptr(cs)
        e.dropout_p, e.rng, e.drop_stream = 0.1, L.ptr(rng), 3
        keep += [u, cs]
    elif kind in ("fc2d_new", "fc2d_new_nocs"):
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); u = torch.randn(M, N, device=dev).bfloat16(); cs = torch.zeros(N, device=dev)
        e.out, e.ldc, e.out_bf16, e.mul_gelu_grad, e.ld_mul, e.mul_mode = L.ptr(out), N, 1, L.ptr(u), N, 1
        if kind == "fc2d_new": e.colsum = L.ptr(cs)
        keep += [u, cs]
    elif kind == "fc1_new":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(out)
        e.out, e.ldc, e.out_bf16, e.bias, e.act, e.out_preact, e.ld_preact, e.preact_mode = L.ptr(out), N, 1, L.ptr(bias), 1, L.ptr(pre), N, 1
        e.dropout_p, e.rng, e.drop_stream = 0.1, L.ptr(rng), 3
        keep += [pre]
    elif kind == "lf":     # MLPDecoderFM.linear_fuse forward: bias, fp32 out
        out = torch.empty(M, N, device=dev); e.out, e.ldc, e.out_bf16, e.bias = L.ptr(out), N, 0, L.ptr(bias)
    elif kind == "res":
        out = torch.empty(M, N, device=dev); res = torch.randn(M, N, device=dev)
        e.out, e.ldc, e.out_bf16, e.bias, e.residual, e.ld_res = L.ptr(out), N, 0, L.ptr(bias), L.ptr(res), N
        e.dropout_p, e.rng, e.drop_stream = 0.1, L.ptr(rng), 3
        keep += [res]
    keep.append(out)
    st = L.stream()
    return (lambda: lib.mdv_gemm_nt(L.ptr(A), K, L.ptr(W), K, M, N, K, ctypes.byref(e), st)), keep

def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    t.record(); torch.cuda.synchronize()
    return s.elapsed_time(t) / n * 1e3

cases = [(131072, 512, 64, k) for k in ("plain", "fc1_nodrop", "fc1", "fc2d", "fc1_new", "fc2d_new", "fc2d_new_nocs")] + [(131072, 64, 512, "res"), (32768, 1024, 128, "fc1"), (32768, 1024, 128, "fc2d"), (131072, 512, 2112, "lf"), (131072, 2112, 512, "plain")]
if os.environ.get("ONE"):
    M, N, K, kind = cases[int(os.environ["ONE"])]
    fn, keep = make(M, N, K, kind)
    for _ in range(3): fn()
    torch.cuda.synchronize()
else:
    for M, N, K, kind in cases:
        fn, keep = make(M, N, K, kind)
        print(f"M={M} N={N} K={K} {kind}: {bench(fn):.1f} us", flush=True)
