import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L
lib = L.lib(); dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
for (M, C, hidden) in ((40000, 64, 512), (60000, 128, 1024)):
    torch.manual_seed(M)
    dy = torch.randn(M, C, device=dev).bfloat16()
    w2t = (torch.randn(hidden, C, device=dev) / C ** 0.5).bfloat16()
    w1t = (torch.randn(C, hidden, device=dev) / hidden ** 0.5).bfloat16()
    u = torch.randn(M, hidden, device=dev).bfloat16()
    du_ref = (dy.float() @ w2t.float().t()) * u.float()
    dx_ref = du_ref.bfloat16().float() @ w1t.float().t()
    for it in range(int(os.environ.get("ITERS", 6))):
        with_w = it % 2 == 0
        du = torch.zeros(M, hidden, device=dev, dtype=torch.bfloat16) if with_w else None
        cs = torch.zeros(hidden, device=dev) if with_w else None
        dx = torch.full((M, C), float("nan"), device=dev)
        L.check(lib.mdv_mlp_bwd(L.ptr(dy), L.ptr(w2t), L.ptr(u), L.ptr(w1t), L.ptr(du), L.ptr(dx), L.ptr(cs), M, C, hidden, L.stream()), "bwd")
        err = (dx - dx_ref).abs().nan_to_num(1e9).amax(dim=1)
        bad = (err > 0.02 * dx_ref.abs().max()).nonzero().flatten()
        msg = f"M={M} C={C} it={it} with_w={with_w} max err {err.max().item():.3e} bad rows {bad.numel()}"
        if bad.numel():
            tiles = torch.unique(bad // 128)
            msg += f" tiles {tiles[:12].tolist()} (cta {[(t % 148) for t in tiles[:12].tolist()]}, lt {[(t // 148) for t in tiles[:12].tolist()]}) rows-in-tile {torch.unique(bad % 128)[:16].tolist()}"
            cols = ((dx - dx_ref).abs()[bad[0]] > 0.02 * dx_ref.abs().max()).nonzero().flatten()
            msg += f" bad cols of first row {cols[:16].tolist()}"
        if with_w:
            e2 = ((du.float() - du_ref).abs().amax(dim=1) > 0.02 * du_ref.abs().max()).nonzero().flatten()
            msg += f" | du bad rows {e2.numel()}"
        print(msg, flush=True)
