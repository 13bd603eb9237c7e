"""Dev: time the fused MLP kernels (mdv_mlp_fwd / mdv_mlp_bwd) at the stage-0/1 shapes of the bench (B=128 stacked)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib()
dev = "cuda"
torch.manual_seed(0)
rng = torch.tensor([1, 2], dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); t.record(); t.synchronize()
        ts.append(s.elapsed_time(t) * 1e3)
    return sum(ts) / len(ts)


only = os.environ.get("ONE")
for M, C, hidden in ((524288, 64, 512), (131072, 128, 1024)):
    a = torch.randn(M, C, device=dev).bfloat16(); w1 = (torch.randn(hidden, C, device=dev) / C ** 0.5).bfloat16(); b1 = torch.randn(hidden, device=dev)
    w2 = (torch.randn(C, hidden, device=dev) / hidden ** 0.5).bfloat16(); b2 = torch.randn(C, device=dev); res = torch.randn(M, C, device=dev)
    out = torch.empty(M, C, device=dev); hact = torch.empty(M, hidden, device=dev, dtype=torch.bfloat16); u = torch.empty_like(hact)
    du = torch.empty_like(hact); cs = torch.zeros(hidden, device=dev); rs = torch.ones(M // 4096, device=dev)
    st = L.stream()
    f_inf = lambda: lib.mdv_mlp_fwd(L.ptr(a), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(res), L.ptr(out), None, None, M, C, hidden, 0.0, None, 0, 0, None, 1, st)
    f_trn = lambda: lib.mdv_mlp_fwd(L.ptr(a), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(res), L.ptr(out), L.ptr(hact), L.ptr(u), M, C, hidden, 0.1, L.ptr(rng), 3, 4, L.ptr(rs), 4096, st)
    b_full = lambda: lib.mdv_mlp_bwd(L.ptr(a), L.ptr(w1), L.ptr(u), L.ptr(w2), L.ptr(du), L.ptr(out), L.ptr(cs), M, C, hidden, st)
    b_act = lambda: lib.mdv_mlp_bwd(L.ptr(a), L.ptr(w1), L.ptr(u), L.ptr(w2), None, L.ptr(out), None, M, C, hidden, st)
    cases = [("fwd inference", f_inf, M * C * (2 + 4 + 4)), ("fwd training", f_trn, M * C * 10 + 2 * M * hidden * 2),
             ("bwd + du + colsum", b_full, M * C * 6 + 2 * M * hidden * 2), ("bwd activation-only", b_act, M * C * 6 + M * hidden * 2)]
    for i, (name, fn, nbytes) in enumerate(cases):
        if only is not None:
            if int(only) == i and C == int(os.environ.get("ONE_C", 64)):
                for _ in range(3): fn()
                torch.cuda.synchronize()
            continue
        us = bench(fn)
        fl = 4.0 * M * C * hidden
        print(f"M={M} C={C} hidden={hidden} {name:22s}: {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s algorithmic  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
