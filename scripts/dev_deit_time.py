"""Dev: forward+backward time of the TransFuse_S_adapt transformer branch (deit_small_patch16_224_adapt, 256x256 input) through
mdvit_b200.transfuse vs the same arithmetic in stock PyTorch eager (fp32 and bf16 autocast) on the same GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from mdvit_b200.transfuse import deit_small_patch16_224_adapt
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False

def eager_forward(m, x, label):
    """vision_transformer.py:149-211 / DeiT.py:121-139 with torch ops on the module's own parameters."""
    B = x.shape[0]
    t = F.conv2d(x, m.patch_embed.proj.weight, m.patch_embed.proj.bias, stride=16).flatten(2).transpose(1, 2) + m.pos_embed
    for blk in m.blocks:
        a = blk.attn
        n1 = F.layer_norm(t, (384,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv = a.qkv(n1).reshape(B, -1, 3, a.num_heads, 64).permute(2, 0, 3, 1, 4)
        att = ((qkv[0] @ qkv[1].transpose(-2, -1)) * a.scale).softmax(dim=-1)
        o = att @ qkv[2]
        g = torch.softmax(a.domain_layer(label).reshape(B, a.num_heads, 1, 64), dim=1)
        t = t + a.proj((g * o).transpose(1, 2).reshape(B, -1, 384))
        n2 = F.layer_norm(t, (384,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        t = t + blk.mlp.fc2(F.gelu(blk.mlp.fc1(n2)))
    return F.layer_norm(t, (384,), m.norm.weight, m.norm.bias, m.norm.eps)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

for B in (32, 128):
    torch.manual_seed(0)
    m = deit_small_patch16_224_adapt(num_domains=4).to(dev).train()
    x = torch.randn(B, 3, 256, 256, device=dev); label = F.one_hot(torch.arange(B, device=dev) % 4, 4).float()
    R = torch.randn(B, 256, 384, device=dev)
    def ours():
        m.zero_grad(set_to_none=True); (m(x, label) * R).sum().backward()
    def eager():
        m.zero_grad(set_to_none=True); (eager_forward(m, x, label) * R).sum().backward()
    def eager_bf16():
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = eager_forward(m, x, label)
        (y.float() * R).sum().backward()
    with torch.no_grad():
        err = ((m(x, label) - eager_forward(m, x, label)).abs().max() / eager_forward(m, x, label).abs().max()).item()
    print(f"B={B}: mdvit_b200 {timeit(ours):.2f} ms  eager fp32 {timeit(eager):.2f} ms  eager bf16 autocast {timeit(eager_bf16):.2f} ms  (fwd max-abs/abs-max vs eager fp32: {err:.2e})", flush=True)
