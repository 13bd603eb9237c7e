"""Drop-ins for the attention modules of TransFuse_S_adapt's DeiT-S branch (BASELINE config 4):
Models/Hybrid_models/TransFuseFolder/vision_transformer.py:96-122 (Attention) and :125-169 (Attention_Sup).  Same constructor,
parameter names and forward signature; the arithmetic is ops.SdpaAttentionFn (tcgen05 GEMMs + the softmax(QK^T)V kernels).
The rest of TransFuse_S_adapt (ResNet34 branch, BiFusion, structure_loss) is not built."""
import torch.nn as nn

from . import ops


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if dim % num_heads or dim // num_heads != 64 or attn_drop or proj_drop:
            raise NotImplementedError("mdvit_b200 implements head_dim 64 without attention / projection dropout (DeiT-S as TransFuse uses it)")
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def _run(self, x, label, da):
        if not x.is_cuda:
            raise RuntimeError("mdvit_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        return ops.SdpaAttentionFn.apply(x, label, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, *da, self.num_heads,
                                         float(self.scale))

    def forward(self, x):
        return self._run(x, None, (None, None, None, None))


class Attention_Sup(Attention):
    """add domain attention adaption (vision_transformer.py:125-169): the per-head softmax gate of the domain label."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., r=2, num_domains=4):
        super().__init__(dim, num_heads, qkv_bias, qk_scale, attn_drop, proj_drop)
        hidden_dim = max(dim // r, 4)
        self.domain_layer = nn.Sequential(nn.Linear(num_domains, hidden_dim), nn.ReLU(inplace=True), nn.Linear(hidden_dim, dim))

    def forward(self, x, domain_label):
        dl = self.domain_layer
        return self._run(x, domain_label, (dl[0].weight, dl[0].bias, dl[2].weight, dl[2].bias))
