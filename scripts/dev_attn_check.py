"""Dev check (GPU): fused factorized attention fwd/bwd vs the oracle's torch ops."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from mdvit_b200 import _lib as L
from oracle import mdvit_oracle as O
lib = L.lib()
dev = torch.device("cuda")
P = L.ptr

def run(B, H, W, C, sup=True):
    torch.manual_seed(0)
    Ch, N = C // 8, H * W
    qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
    ws = {}
    sd = {}
    for i, (win, hh) in enumerate(((3, 2), (5, 3), (7, 3))):
        sd[f"crpe.conv_list.{i}.weight"] = (torch.randn(hh * Ch, 1, win, win, device=dev) * 0.2).requires_grad_()
        sd[f"crpe.conv_list.{i}.bias"] = (torch.randn(hh * Ch, device=dev) * 0.2).requires_grad_()
    gate = torch.softmax(torch.randn(B, 8, Ch, device=dev), dim=1).reshape(B, C).contiguous() if sup else None
    # ---- reference (fp32 math on the bf16-rounded qkv)
    q32 = qkv.float().requires_grad_()
    t = q32.reshape(B, N, 3, 8, Ch).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    ks = k.softmax(dim=2)
    fa = (Ch ** -0.5) * torch.einsum("bhnk,bhkv->bhnv", q, torch.einsum("bhnk,bhnv->bhkv", ks, v)) + O.conv_rel_pos_enc(sd, "crpe", q, v, H, W)
    gref = gate.clone().requires_grad_() if sup else None
    if sup: fa = gref.reshape(B, 8, 1, Ch) * fa
    yref = fa.transpose(1, 2).reshape(B, N, C)
    # ---- ours
    stats = torch.empty(lib.mdv_attn_stats_floats(B, C, 8), device=dev)
    y = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
    cw = [sd[f"crpe.conv_list.{i}.{n}"].detach() for i in range(3) for n in ("weight", "bias")]
    L.check(lib.mdv_attn_fwd(P(qkv), P(gate), *[P(t_) for t_ in cw], P(stats), P(y), B, H, W, C, 8, L.stream()), "attn_fwd")
    torch.cuda.synchronize()
    e_f = ((y.float() - yref).abs().max() / yref.abs().max()).item()
    dy = torch.randn(B, N, C, device=dev).bfloat16()
    yref.backward(dy.float())
    dqkv = torch.empty_like(qkv)
    dgate = torch.zeros(B, C, device=dev) if sup else None
    gcw = [torch.zeros_like(t_) for t_ in cw]
    wsb = torch.empty(B * C * (2 * Ch + 1), device=dev)
    L.check(lib.mdv_attn_bwd(P(qkv), P(dy), P(y), P(gate), *[P(t_) for t_ in cw], P(stats), P(dqkv), P(dgate), *[P(t_) for t_ in gcw],
                             P(wsb), B, H, W, C, 8, L.stream()), "attn_bwd")
    torch.cuda.synchronize()
    r = lambda a, b: ((a.float() - b).abs().max() / (b.abs().max() + 1e-12)).item()
    g = q32.grad
    msg = f"B={B} H={H} W={W} C={C} sup={sup}: fwd {e_f:.2e} dq {r(dqkv[..., :C], g[..., :C]):.2e} dk {r(dqkv[..., C:2*C], g[..., C:2*C]):.2e} dv {r(dqkv[..., 2*C:], g[..., 2*C:]):.2e}"
    for i in range(3):
        msg += f" dw{i} {r(gcw[2*i], sd[f'crpe.conv_list.{i}.weight'].grad):.2e} db{i} {r(gcw[2*i+1], sd[f'crpe.conv_list.{i}.bias'].grad):.2e}"
    if sup: msg += f" dgate {r(dgate, gref.grad):.2e}"
    print(msg, flush=True)

for cfg in [(2, 16, 16, 64), (2, 8, 8, 128), (2, 4, 4, 320), (2, 2, 2, 512), (3, 12, 20, 64), (1, 64, 64, 64), (2, 16, 16, 320), (2, 8, 8, 512)]:
    run(*cfg)
run(2, 16, 16, 128, sup=False)
