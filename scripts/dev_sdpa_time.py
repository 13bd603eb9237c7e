"""Dev: time / profile mdv_sdpa_fwd and mdv_sdpa_bwd (softmax(QK^T)V + DA gate, TransFuse DeiT-S shape: N=256, 6 heads x 64)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib(); dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in (32, 128, 512):
    N, C, H = 256, 384, 6
    qkv = (torch.randn(B * N, 3 * C, device=dev) * 1.5).bfloat16()
    gate = torch.softmax(torch.randn(B, H, 64, device=dev), dim=1).reshape(B, C).contiguous()
    out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: lib.mdv_sdpa_fwd(L.ptr(qkv), L.ptr(gate), L.ptr(out), None, B, N, C, H, ctypes.c_float(0.125), L.stream())
    if os.environ.get("ONE"):
        if B == int(os.environ["ONE"]):
            for _ in range(3): fn()
            torch.cuda.synchronize()
        continue
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); t.record(); t.synchronize(); ts.append(s.elapsed_time(t) * 1e3)
    us = sum(ts) / len(ts)
    fl = 4.0 * B * H * N * N * 64
    by = B * N * C * 2 * 4
    print(f"B={B}: {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s  {by / us / 1e3:.1f} GB/s algorithmic (qkv read + out write)", flush=True)
    # backward: 7 GEMMs per head (S, dP twice; dQ, dK, dV), reads qkv + out + dout, writes dqkv
    lse = torch.empty(B, H, N, device=dev); dy = torch.randn(B * N, C, device=dev).bfloat16(); dqkv = torch.empty_like(qkv); dgate = torch.empty(B, C, device=dev)
    lib.mdv_sdpa_fwd(L.ptr(qkv), L.ptr(gate), L.ptr(out), L.ptr(lse), B, N, C, H, ctypes.c_float(0.125), L.stream())
    fb = lambda: lib.mdv_sdpa_bwd(L.ptr(qkv), L.ptr(gate), L.ptr(out), L.ptr(lse), L.ptr(dy), L.ptr(dqkv), L.ptr(dgate), B, N, C, H, ctypes.c_float(0.125), L.stream())
    for _ in range(3): fb()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fb(); t.record(); t.synchronize(); ts.append(s.elapsed_time(t) * 1e3)
    us = sum(ts) / len(ts)
    print(f"B={B}: bwd {us:.1f} us  {14.0 * B * H * N * N * 64 / us / 1e6:.1f} TFLOP/s executed  {B * N * C * 2 * 8 / us / 1e3:.1f} GB/s algorithmic (qkv, out, dout read + dqkv write)", flush=True)
