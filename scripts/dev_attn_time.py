"""Dev: time mdv_attn_fwd / mdv_attn_bwd at the four stage shapes (B=32)."""
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
B = 32
for (H, C) in ((64, 64), (32, 128), (16, 320), (8, 512)):
    Ch, N = C // 8, H * H
    qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
    cw = []
    for win, hh in ((3, 2), (5, 3), (7, 3)):
        cw += [torch.randn(hh * Ch, 1, win, win, device=dev) * 0.2, torch.randn(hh * Ch, device=dev) * 0.2]
    gate = torch.softmax(torch.randn(B, 8, Ch, device=dev), dim=1).reshape(B, C).contiguous()
    stats = torch.empty(lib.mdv_attn_stats_floats(B, C, 8), device=dev); ws = torch.empty(lib.mdv_attn_ws_floats(B, C, 8), device=dev)
    y = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16); dy = torch.randn(B, N, C, device=dev).bfloat16()
    ec = torch.empty_like(y); dqkv = torch.empty_like(qkv); dgate = torch.zeros(B, C, device=dev); gcw = [torch.zeros_like(t) for t in cw]
    st = L.stream()
    f = lambda: lib.mdv_attn_fwd(P(qkv), P(gate), *[P(t) for t in cw], P(stats), P(ws), P(y), P(ec), B, H, H, C, 8, st)
    bw = lambda: lib.mdv_attn_bwd(P(qkv), P(dy), P(y), P(ec), P(gate), *[P(t) for t in cw], P(stats), P(dqkv), P(dgate), *[P(t) for t in gcw], None, P(ws), B, H, H, C, 8, st)
    bd = lambda: lib.mdv_attn_bwd(P(qkv), P(dy), P(y), P(ec), P(gate), *[P(t) for t in cw], P(stats), P(dqkv), P(dgate), None, None, None, None, None, None, None, P(ws), B, H, H, C, 8, st)
    mb = B * N * C * 2 / 1e6
    print(f"H={H} C={C}: fwd {bench(f):.1f} us (ideal {5*mb/6.45:.1f}), bwd {bench(bw):.1f} us, bwd(no wgrad) {bench(bd):.1f} us (ideal {9*mb/6.45:.1f})", flush=True)
