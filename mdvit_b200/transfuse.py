"""Drop-ins for the attention modules of TransFuse_S_adapt's DeiT-S branch (BASELINE config 4):
Models/Hybrid_models/TransFuseFolder/vision_transformer.py:96-122 (Attention) and :125-169 (Attention_Sup).  Same constructor,
parameter names and forward signature; the arithmetic is ops.SdpaAttentionFn (tcgen05 GEMMs + the softmax(QK^T)V kernels).
The rest of TransFuse_S_adapt (ResNet34 branch, BiFusion, structure_loss) is not built."""
import math
from functools import partial

import torch
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


class Mlp(nn.Module):
    """vision_transformer.py:78-94 (parameter container)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if drop or act_layer is not nn.GELU or (out_features or in_features) != in_features:
            raise NotImplementedError("mdvit_b200 implements the GELU Mlp without dropout")
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, in_features)
        self.drop = nn.Dropout(drop)


class _BlockBase(nn.Module):
    def _build(self, dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, act_layer, norm_layer, attn):
        if drop or attn_drop or drop_path:
            raise NotImplementedError("mdvit_b200 implements the DeiT blocks without dropout / DropPath (TransFuse_S_adapt's setting)")
        self.norm1 = norm_layer(dim)
        self.attn = attn
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _run(self, x, label):
        a, m = self.attn, self.mlp
        if not x.is_cuda:
            raise RuntimeError("mdvit_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        dl = getattr(a, "domain_layer", None)
        da = (dl[0].weight, dl[0].bias, dl[2].weight, dl[2].bias) if (dl is not None and label is not None) else (None,) * 4
        return ops.DeiTBlockFn.apply(x, label, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, *da,
                                     self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, a.num_heads,
                                     float(a.scale), float(self.norm1.eps))


class Block(_BlockBase):
    """vision_transformer.py:172-188."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self._build(dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, act_layer, norm_layer,
                    Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop))

    def forward(self, x):
        return self._run(x, None)


class Block_adapt(_BlockBase):
    """vision_transformer.py:191-211."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, num_domains=4):
        super().__init__()
        self._build(dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, act_layer, norm_layer,
                    Attention_Sup(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                                  num_domains=num_domains))

    def forward(self, x, domain_label):
        return self._run(x, domain_label)


class PatchEmbed(nn.Module):
    """vision_transformer.py:214-236 (parameter container)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class DeiT_adapt(nn.Module):
    """DeiT.py:116-139 over VisionTransformer_adapt (vision_transformer.py:322-389): the transformer branch of TransFuse_S_adapt.
    forward(x [B,3,H,W], domain_label [B,num_domains]) -> tokens [B, (H/16)(W/16), embed_dim].  The number of tokens must be 128
    or 256 (256x256 images give 256) and must match pos_embed."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.,
                 qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., hybrid_backbone=None,
                 norm_layer=nn.LayerNorm, num_domains=4):
        super().__init__()
        if hybrid_backbone is not None or drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("mdvit_b200 implements the patch-embedding DeiT without dropout / DropPath")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([
            Block_adapt(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                        attn_drop=attn_drop_rate, drop_path=0., norm_layer=norm_layer, num_domains=num_domains) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, self.embed_dim))      # DeiT.py:119-120

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x, domain_label):
        B, Cin, H, W = x.shape
        p = self.patch_embed.patch_size[0]
        n = (H // p) * (W // p)
        if self.pos_embed.shape[1] != n:
            raise ValueError(f"pos_embed holds {self.pos_embed.shape[1]} positions, the image gives {n} patches")
        # patches in the (c, i, j) order of the flattened Conv2d weight: a pure permutation of the image (host-side glue)
        patches = x.reshape(B, Cin, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B * n, Cin * p * p)
        w = self.patch_embed.proj
        t = ops.DeiTEmbedFn.apply(patches, w.weight, w.bias, self.pos_embed, B, n)
        for blk in self.blocks:
            t = blk(t, domain_label)
        return ops.LayerNormOutFn.apply(t, self.norm.weight, self.norm.bias, float(self.norm.eps))


def deit_small_patch16_224_adapt(pretrained=False, pretrained_folder=None, num_domains=4, **kwargs):
    """DeiT.py:157-181: DeiT-S (384 wide, depth 8, 6 heads) with pos_embed resampled to a 16 x 16 grid (256 tokens), no head."""
    if pretrained:
        raise NotImplementedError("loading the ImageNet checkpoint is the reference's job: load_state_dict() the result here")
    model = DeiT_adapt(patch_size=16, embed_dim=384, depth=8, num_heads=6, mlp_ratio=4, qkv_bias=True,
                       norm_layer=partial(nn.LayerNorm, eps=1e-6), num_domains=num_domains, **kwargs)
    pe = model.pos_embed[:, 1:, :].detach().transpose(-1, -2)
    g = int(math.sqrt(pe.shape[2]))
    pe = torch.nn.functional.interpolate(pe.reshape(pe.shape[0], pe.shape[1], g, g), size=(16, 16), mode='bilinear', align_corners=True)
    model.pos_embed = nn.Parameter(pe.flatten(2).transpose(-1, -2))
    model.head = nn.Identity()
    return model
