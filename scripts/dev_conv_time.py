"""Dev: time the depthwise-conv / resize kernels at the stage shapes (B=32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib(); dev = "cuda"; P = L.ptr
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    t.record(); torch.cuda.synchronize()
    return s.elapsed_time(t) / n * 1e3
B = 32; st = L.stream()
for (H, C) in ((64, 64), (32, 128), (16, 320), (8, 512)):
    x = torch.randn(B, H, H, C, device=dev); w = torch.randn(C, 1, 3, 3, device=dev); b = torch.randn(C, device=dev)
    out = torch.empty_like(x); dw = torch.zeros_like(w); db = torch.zeros_like(b)
    mb = x.numel() * 4 / 1e6
    f = lambda: lib.mdv_dwconv3(P(x), P(w), P(b), P(out), 0, B, H, H, H, H, C, 1, 0, 1, st)
    g = lambda: lib.mdv_dwconv3_wgrad(P(out), P(x), P(dw), P(db), B, H, H, H, H, C, 1, st)
    print(f"dwconv3 H={H} C={C}: fwd {bench(f):.1f} us (ideal {2*mb/6.45:.1f}), wgrad {bench(g):.1f} us (ideal {2*mb/6.45:.1f})", flush=True)
