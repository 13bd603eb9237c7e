"""Development aid: the three implicit-GEMM 3x3 convolution launches (forward TF32, input gradient bf16, weight gradient bf16) at
resnet34.layer1's shape for a stacked TransFuse step (128 images, 64 x 64 x 64 -> 64), timed with CUDA events and compared with
the im2col formulation; run under `ncu --set full -k regex:gemm_kernel` for the profile in profiles/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdvit_b200 import _lib as L, ops      # noqa: E402

B, H, W, Cin, Cout = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (128, 64, 64, 64, 64)))
dev = torch.device("cuda")
M = B * H * W
x = torch.randn(M, Cin, device=dev)
w = torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5
dz = torch.randn(M, Cout, device=dev).bfloat16()
lib = L.lib()
Wf = ops.prep_weight(w, 2 | 8, Cout, 9 * Cin, cin=Cin)
Wd = ops.prep_weight(w, 4, Cout, 9 * Cin)
xb = x.bfloat16()
z = torch.empty(M, Cout, device=dev)
dx = torch.empty(M, Cin, device=dev)
gw = torch.zeros(Cout, 9 * Cin, device=dev)
col = torch.empty(M, 9 * Cin, device=dev)


def fwd():
    ops.conv3_gemm(x, Wf, B, H, W, Cin, Cout, z)


def dgrad():
    ops.conv3_gemm(dz, Wd, B, H, W, Cout, Cin, dx, flip=True)


def wgrad():
    L.check(lib.mdv_conv3_wgrad(L.ptr(dz), Cout, L.ptr(xb), Cin, B, H, W, Cin, Cout, L.ptr(gw), 9 * Cin, L.stream()), "wgrad")


def fwd_im2col():
    L.check(lib.mdv_im2col3(L.ptr(x), 0, L.ptr(col), 0, B, H, W, H, W, Cin, 1, 9 * Cin, L.stream()), "im2col3")
    ops.gemm_nt(col, Wf, M, Cout, 9 * Cin, z, tf32=True)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


flops = 2.0 * M * Cout * 9 * Cin
for name, fn, by in (("forward (TF32, implicit)", fwd, 4 * M * (Cin + Cout)), ("input gradient (bf16, implicit)", dgrad, 2 * M * Cout + 4 * M * Cin),
                     ("weight gradient (bf16, implicit)", wgrad, 2 * M * (Cin + Cout)), ("forward via im2col + GEMM", fwd_im2col, 4 * M * (Cin + Cout))):
    us = t(fn)
    print(f"{name:36s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  algorithmic {by / 1e6:7.1f} MB -> {by / us / 1e3:7.1f} GB/s")
