import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib(); dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
M, C, hidden = 60000, 128, 1024
torch.manual_seed(1)
for variant in ("u_ones", "T_ones", "random"):
    if variant == "u_ones":
        dy = torch.randn(M, C, device=dev).bfloat16(); w2t = (torch.randn(hidden, C, device=dev) / C ** 0.5).bfloat16()
        u = torch.ones(M, hidden, device=dev).bfloat16()
    elif variant == "T_ones":
        dy = torch.ones(M, C, device=dev).bfloat16(); w2t = torch.full((hidden, C), 1.0 / C, device=dev).bfloat16()
        u = torch.randn(M, hidden, device=dev).bfloat16()
    else:
        dy = torch.randn(M, C, device=dev).bfloat16(); w2t = (torch.randn(hidden, C, device=dev) / C ** 0.5).bfloat16()
        u = torch.randn(M, hidden, device=dev).bfloat16()
    w1t = (torch.randn(C, hidden, device=dev) / hidden ** 0.5).bfloat16()
    du_ref = (dy.float() @ w2t.float().t()) * u.float()
    dx_ref = du_ref.bfloat16().float() @ w1t.float().t()
    for it in range(12):
        dx = torch.empty(M, C, device=dev)
        L.check(lib.mdv_mlp_bwd(L.ptr(dy), L.ptr(w2t), L.ptr(u), L.ptr(w1t), None, L.ptr(dx), None, M, C, hidden, L.stream()), "bwd")
        e2 = (dx - dx_ref).abs().amax(dim=1)
        badr = (e2 > 0.02 * dx_ref.abs().max()).nonzero().flatten()
        if badr.numel():
            t = torch.unique(badr // 128)
            print(f"{variant} ACT-ONLY it={it} bad rows {badr.numel()} tiles {t[:10].tolist()} lt {[int(x) // 148 for x in t[:10]]} rows-in-tile {torch.unique(badr % 128)[:24].tolist()}", flush=True)
        else:
            print(f"{variant} ACT-ONLY it={it} ok", flush=True)
    for it in range(2):
        du = torch.zeros(M, hidden, device=dev, dtype=torch.bfloat16)
        dx = torch.empty(M, C, device=dev)
        L.check(lib.mdv_mlp_bwd(L.ptr(dy), L.ptr(w2t), L.ptr(u), L.ptr(w1t), L.ptr(du), L.ptr(dx), None, M, C, hidden, L.stream()), "bwd")
        err = (du.float() - du_ref).abs()
        bad = (err > 0.02 * du_ref.abs().max()).nonzero()
        msg = f"{variant} it={it} bad elems {bad.shape[0]}"
        if bad.shape[0]:
            rows = torch.unique(bad[:, 0]); cols = torch.unique(bad[:, 1])
            t = torch.unique(rows // 128)
            msg += f" rows {rows.numel()} tiles {t[:8].tolist()} lt {[int(x) // 148 for x in t[:8]]} rows-in-tile {torch.unique(rows % 128)[:20].tolist()} chunks {torch.unique(cols // 64)[:16].tolist()} cols-in-chunk {torch.unique(cols % 64)[:8].tolist()}..{int((cols % 64).max())}"
            r0, c0 = int(bad[0, 0]), int(bad[0, 1])
            msg += f" | first bad ({r0},{c0}) got {du[r0, c0].item():.4f} want {du_ref[r0, c0].item():.4f}"
            # does the bad value equal the reference value of another tile's same position?
            for dt in (-2, -1, 1, 2):
                rr = r0 + dt * 148 * 128
                if 0 <= rr < M: msg += f" [tile{dt:+d}: {du_ref[rr, c0].item():.4f}]"
        print(msg, flush=True)
