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


def load_pretrain(model, pre_s_dict):
    """DeiT.py:74-90: copy the entries of `pre_s_dict` whose keys the model has, keep the model's own values for the rest
    (the reference passes the checkpoint file's top-level dict, so with the published DeiT file — {'model': ...} — nothing
    matches and the model keeps its initialisation; the behaviour is reproduced as is)."""
    s_dict = model.state_dict()
    missing = [k for k in s_dict if k not in pre_s_dict]
    print('{} keys are not in the pretrain model:'.format(len(missing)), missing)
    model.load_state_dict({k: (pre_s_dict[k] if k in pre_s_dict else v) for k, v in s_dict.items()})
    return model


def deit_small_patch16_224_adapt(pretrained=False, pretrained_folder=None, num_domains=4, **kwargs):
    """DeiT.py:157-181: DeiT-S (384 wide, depth 8, 6 heads) with pos_embed resampled to a 16 x 16 grid (256 tokens), no head."""
    model = DeiT_adapt(patch_size=16, embed_dim=384, depth=8, num_heads=6, mlp_ratio=4, qkv_bias=True,
                       norm_layer=partial(nn.LayerNorm, eps=1e-6), num_domains=num_domains, **kwargs)
    if pretrained:      # DeiT.py:125-127
        load_pretrain(model, torch.load(pretrained_folder + '/pretrained/deit_small_patch16_224-cd65a155.pth'))
    pe = model.pos_embed[:, 1:, :].detach().transpose(-1, -2)
    g = int(math.sqrt(pe.shape[2]))
    pe = torch.nn.functional.interpolate(pe.reshape(pe.shape[0], pe.shape[1], g, g), size=(16, 16), mode='bilinear', align_corners=True)
    model.pos_embed = nn.Parameter(pe.flatten(2).transpose(-1, -2))
    model.head = nn.Identity()
    return model


# ================================================================================================ TransFuse_S_adapt
# Models/Hybrid_models/TransFuseFolder/TransFuse.py:20-78 (ChannelPool, BiFusion_block), :182-283 (TransFuse_S_adapt), :533-650
# (init_weights, Up, Attention_block, DoubleConv, Residual, Conv).  Same class names, constructor signatures, registration
# order (=> the same state_dict keys and, under one seed, the same initial weights) and forward signatures.  Inside, maps are
# NHWC fp32 ([B, H*W, C]); every convolution / BatchNorm / pooling / resize is a C-ABI kernel (ops.ConvBnActFn, ops.BnActFn,
# ops.MaxPool3s2Fn, ops.ResizeACFn, ops.GateCatFn, ops.ChannelPoolFn, ops.Dropout2dFn); what remains as torch tensor expressions are
# the small per-pixel / per-sample pieces (the squeeze-and-excitation mean + Linears on [B, C], the two single-channel
# BatchNorms + sigmoids on [B, H*W, 1]) and the W_g * W_x product / Up-block concatenations.  Each block's `forward` keeps the reference's NCHW signature; `run` is the NHWC form the model
# chains internally.
import torch.nn.functional as F

ACT_NONE, ACT_RELU = ops.ACT_NONE, ops.ACT_RELU


def _to_nhwc(x):
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C), H, W


def _to_nchw(x, H, W):
    B, _, C = x.shape
    return x.view(B, H, W, C).permute(0, 3, 1, 2)


def _bn_bufs(bn):
    return (bn.running_mean, bn.running_var, bn.num_batches_tracked)


def _cba(x, H, W, conv, bn=None, act=ACT_NONE, residual=None, nchw=False):
    """act(BN?(conv(x)) + residual?) -> (y [B, Ho*Wo, Cout], Ho, Wo)"""
    k, s = conv.kernel_size[0], conv.stride[0]
    if conv.kernel_size[0] != conv.kernel_size[1] or conv.padding[0] != (k - 1) // 2 or conv.groups != 1 or conv.dilation[0] != 1:
        raise NotImplementedError("mdvit_b200: square dense convolutions with 'same' padding only")
    B = x.shape[0]
    training = bn.training if bn is not None else False
    y = ops.ConvBnActFn.apply(x, conv.weight, conv.bias, bn.weight if bn is not None else None, bn.bias if bn is not None else None,
                              residual, _bn_bufs(bn) if bn is not None else None, B, H, W, s, act, training, nchw)
    Ho, Wo, _ = ops.conv_geom(H, W, k, s)
    return y, Ho, Wo


_BN1_COEF = {}


def _bn1(z, bn, H, W):
    """BatchNorm2d(1) on a single-channel NHWC map [B, H*W, 1].  Under ops.bn_groups(G) (the stacked multi-dataset forward) the
    batch statistics are taken per group of B/G consecutive samples and the running statistics receive the G momentum updates
    of G consecutive forwards, in group order."""
    B = z.shape[0]
    G = ops.current_bn_groups() if bn.training else 1
    if G == 1:
        if bn.training:
            bn.num_batches_tracked.add_(1)      # (nn.BatchNorm2d.forward does this; the functional does not)
        return F.batch_norm(z.view(B, 1, H, W), bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, bn.momentum, bn.eps).view(B, H * W, 1)
    zg = z.reshape(1, G, -1)      # groups as the channel axis of one batch_norm call (one kernel forward, one backward)
    n = zg.shape[2]
    y = F.batch_norm(zg, None, None, bn.weight.expand(G), bn.bias.expand(G), True, 0.0, bn.eps)
    with torch.no_grad():
        var, mean = torch.var_mean(zg[0], dim=1, unbiased=False)
    mom = bn.momentum
    key = (G, mom, z.device)
    coef = _BN1_COEF.get(key)
    if coef is None:
        coef = _BN1_COEF[key] = torch.tensor([mom * (1 - mom) ** (G - 1 - g) for g in range(G)], dtype=torch.float32, device=z.device)
    with torch.no_grad():
        bn.running_mean.mul_((1 - mom) ** G).add_((coef * mean.flatten()).sum())
        bn.running_var.mul_((1 - mom) ** G).add_((coef * var.flatten()).sum() * (n / (n - 1)))
        bn.num_batches_tracked.add_(G)
    return y.view(B, H * W, 1)


def _drop2d(x, p, training):
    """nn.Dropout2d on an NHWC map: whole (sample, channel) planes"""
    if p <= 0. or not training:
        return x
    return ops.Dropout2dFn.apply(x, float(p))


class ChannelPool(nn.Module):
    def run(self, x):
        return ops.ChannelPoolFn.apply(x)

    def forward(self, x):
        return torch.cat((torch.max(x, 1)[0].unsqueeze(1), torch.mean(x, 1).unsqueeze(1)), dim=1)


class Conv(nn.Module):
    def __init__(self, inp_dim, out_dim, kernel_size=3, stride=1, bn=False, relu=True, bias=True):
        super().__init__()
        self.inp_dim = inp_dim
        self.conv = nn.Conv2d(inp_dim, out_dim, kernel_size, stride, padding=(kernel_size - 1) // 2, bias=bias)
        self.relu = None
        self.bn = None
        if relu:
            self.relu = nn.ReLU(inplace=True)
        if bn:
            self.bn = nn.BatchNorm2d(out_dim)

    def run(self, x, H, W, residual=None):
        assert x.shape[2] == self.inp_dim, "{} {}".format(x.shape[2], self.inp_dim)
        act = ACT_RELU if self.relu is not None else ACT_NONE
        if self.conv.out_channels == 1:      # single-channel output: row-dot kernel, BatchNorm2d(1) / ReLU as tensor expressions
            y, Ho, Wo = _cba(x, H, W, self.conv)
            if self.bn is not None:
                y = _bn1(y, self.bn, Ho, Wo)
            return (torch.relu(y) if self.relu is not None else y), Ho, Wo
        return _cba(x, H, W, self.conv, self.bn, act, residual)

    def forward(self, x):
        t, H, W = _to_nhwc(x)
        y, Ho, Wo = self.run(t, H, W)
        return _to_nchw(y, Ho, Wo)


class Residual(nn.Module):
    def __init__(self, inp_dim, out_dim):
        super().__init__()
        self.relu = nn.ReLU(inplace=True)
        self.bn1 = nn.BatchNorm2d(inp_dim)
        self.conv1 = Conv(inp_dim, int(out_dim / 2), 1, relu=False)
        self.bn2 = nn.BatchNorm2d(int(out_dim / 2))
        self.conv2 = Conv(int(out_dim / 2), int(out_dim / 2), 3, relu=False)
        self.bn3 = nn.BatchNorm2d(int(out_dim / 2))
        self.conv3 = Conv(int(out_dim / 2), out_dim, 1, relu=False)
        self.skip_layer = Conv(inp_dim, out_dim, 1, relu=False)
        self.need_skip = inp_dim != out_dim

    def run(self, x, H, W):
        residual = self.skip_layer.run(x, H, W)[0] if self.need_skip else x
        out = ops.BnActFn.apply(x, self.bn1.weight, self.bn1.bias, _bn_bufs(self.bn1), ACT_RELU, self.bn1.training)
        out, _, _ = _cba(out, H, W, self.conv1.conv, self.bn2, ACT_RELU)      # conv1 -> bn2 -> relu
        out, _, _ = _cba(out, H, W, self.conv2.conv, self.bn3, ACT_RELU)      # conv2 -> bn3 -> relu
        out, _, _ = _cba(out, H, W, self.conv3.conv, None, ACT_NONE, residual)      # conv3, out += residual
        return out

    def forward(self, x):
        t, H, W = _to_nhwc(x)
        return _to_nchw(self.run(t, H, W), H, W)


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels)
        )
        self.identity = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0),
            nn.BatchNorm2d(out_channels)
        )
        self.relu = nn.ReLU(inplace=True)

    def run(self, x, H, W):
        dc, idt = self.double_conv, self.identity
        i, _, _ = _cba(x, H, W, idt[0], idt[1], ACT_NONE)
        y, _, _ = _cba(x, H, W, dc[0], dc[1], ACT_RELU)
        y, _, _ = _cba(y, H, W, dc[3], dc[4], ACT_RELU, i)      # relu(double_conv(x) + identity(x))
        return y

    def forward(self, x):
        t, H, W = _to_nhwc(x)
        return _to_nchw(self.run(t, H, W), H, W)


class Attention_block(nn.Module):
    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, kernel_size=1, stride=1, padding=0, bias=True), nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, kernel_size=1, stride=1, padding=0, bias=True), nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, kernel_size=1, stride=1, padding=0, bias=True), nn.BatchNorm2d(1), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)

    def run(self, g, x, H, W):
        g1, _, _ = _cba(g, H, W, self.W_g[0], self.W_g[1], ACT_NONE)
        s, _, _ = _cba(x, H, W, self.W_x[0], self.W_x[1], ACT_RELU, g1)      # relu(g1 + x1)
        p, _, _ = _cba(s, H, W, self.psi[0])
        p = torch.sigmoid(_bn1(p, self.psi[1], H, W))
        return ops.GateCatFn.apply(x, p, None, None, None)      # x * psi

    def forward(self, g, x):
        gt, H, W = _to_nhwc(g)
        xt, _, _ = _to_nhwc(x)
        return _to_nchw(self.run(gt, xt, H, W), H, W)


class Up(nn.Module):
    """Upscaling then double conv"""

    def __init__(self, in_ch1, out_ch, in_ch2=0, attn=False):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.conv = DoubleConv(in_ch1 + in_ch2, out_ch)
        self.attn_block = Attention_block(in_ch1, in_ch2, out_ch) if attn else None

    def run(self, x1, H, W, x2=None):
        """x1 at H x W is upsampled to 2H x 2W, where x2 (if given) lives; returns the map at 2H x 2W"""
        Ho, Wo = 2 * H, 2 * W
        x1 = ops.ResizeACFn.apply(x1, H, W, Ho, Wo)
        if x2 is not None:
            if x2.shape[1] != Ho * Wo:
                raise NotImplementedError("mdvit_b200: Up expects the skip map at exactly twice the resolution")
            if self.attn_block is not None:
                x2 = self.attn_block.run(x1, x2, Ho, Wo)
            x1 = torch.cat([x2, x1], dim=2)
        return self.conv.run(x1, Ho, Wo)

    def forward(self, x1, x2=None):
        t1, H, W = _to_nhwc(x1)
        t2 = _to_nhwc(x2)[0] if x2 is not None else None
        return _to_nchw(self.run(t1, H, W, t2), 2 * H, 2 * W)


class BiFusion_block(nn.Module):
    def __init__(self, ch_1, ch_2, r_2, ch_int, ch_out, drop_rate=0.):
        super().__init__()
        # channel attention for F_g, use SE Block
        self.fc1 = nn.Conv2d(ch_2, ch_2 // r_2, kernel_size=1)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(ch_2 // r_2, ch_2, kernel_size=1)
        self.sigmoid = nn.Sigmoid()
        # spatial attention for F_l
        self.compress = ChannelPool()
        self.spatial = Conv(2, 1, 7, bn=True, relu=False, bias=False)
        # bi-linear modelling for both
        self.W_g = Conv(ch_1, ch_int, 1, bn=True, relu=False)
        self.W_x = Conv(ch_2, ch_int, 1, bn=True, relu=False)
        self.W = Conv(ch_int, ch_int, 3, bn=True, relu=True)
        self.relu = nn.ReLU(inplace=True)
        self.residual = Residual(ch_1 + ch_2 + ch_int, ch_out)
        self.dropout = nn.Dropout2d(drop_rate)
        self.drop_rate = drop_rate

    def run(self, g, x, H, W):
        # bilinear pooling
        W_g, _, _ = self.W_g.run(g, H, W)
        W_x, _, _ = self.W_x.run(x, H, W)
        bp, _, _ = self.W.run(W_g * W_x, H, W)
        # spatial attention for cnn branch: per-pixel gate sigmoid(BN(conv7x7(ChannelPool(g))))
        s, _, _ = self.spatial.run(self.compress.run(g), H, W)
        # channel attention for transformer branch (the 1x1 convs act on the [B, C] pooled vector)
        v = x.mean(dim=1)
        v = torch.relu(F.linear(v, self.fc1.weight.flatten(1), self.fc1.bias))
        v = torch.sigmoid(F.linear(v, self.fc2.weight.flatten(1), self.fc2.bias))
        # both gates and torch.cat([g, x, bp], 1) in one kernel
        fuse = self.residual.run(ops.GateCatFn.apply(g, torch.sigmoid(s), x, v, bp), H, W)
        return _drop2d(fuse, self.drop_rate, self.training)

    def forward(self, g, x):
        gt, H, W = _to_nhwc(g)
        xt, _, _ = _to_nhwc(x)
        return _to_nchw(self.run(gt, xt, H, W), H, W)


def init_weights(m):
    """TransFuse.py:533-553"""
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
        if m.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(m.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(m.bias, -bound, bound)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)


class TransFuse_S_adapt(nn.Module):
    """TransFuse.py:182-283: ResNet34 (layers 1-3) || DeiT-S-adapt, fused by BiFusion blocks and attention-gated Up blocks; returns
    the three logit maps (map_x, map_1, map_2), each [B, num_classes, H, W].  256 x 256 inputs (16 x 16 tokens)."""

    def __init__(self, num_classes=1, drop_rate=0.2, normal_init=True, pretrained=False,
                 pretrained_folder='/bigdata/siyiplace/data/skin_lesion', num_domains=4):
        super().__init__()
        from torchvision.models import resnet34      # parameter container only (same keys / init as the reference); forward is ours
        if num_classes != 1:
            raise NotImplementedError("mdvit_b200 implements the binary-segmentation heads (num_classes=1)")
        self.resnet = resnet34()
        if pretrained:
            self.resnet.load_state_dict(torch.load(pretrained_folder + '/pretrained/resnet34-333f7ec4.pth'))
        self.resnet.fc = nn.Identity()
        self.resnet.layer4 = nn.Identity()
        self.transformer = deit_small_patch16_224_adapt(pretrained=pretrained, pretrained_folder=pretrained_folder, num_domains=num_domains)
        self.up1 = Up(in_ch1=384, out_ch=128)
        self.up2 = Up(128, 64)
        self.final_x = nn.Sequential(Conv(256, 64, 1, bn=True, relu=True), Conv(64, 64, 3, bn=True, relu=True),
                                     Conv(64, num_classes, 3, bn=False, relu=False))
        self.final_1 = nn.Sequential(Conv(64, 64, 3, bn=True, relu=True), Conv(64, num_classes, 3, bn=False, relu=False))
        self.final_2 = nn.Sequential(Conv(64, 64, 3, bn=True, relu=True), Conv(64, num_classes, 3, bn=False, relu=False))
        self.up_c = BiFusion_block(ch_1=256, ch_2=384, r_2=4, ch_int=256, ch_out=256, drop_rate=drop_rate / 2)
        self.up_c_1_1 = BiFusion_block(ch_1=128, ch_2=128, r_2=2, ch_int=128, ch_out=128, drop_rate=drop_rate / 2)
        self.up_c_1_2 = Up(in_ch1=256, out_ch=128, in_ch2=128, attn=True)
        self.up_c_2_1 = BiFusion_block(ch_1=64, ch_2=64, r_2=1, ch_int=64, ch_out=64, drop_rate=drop_rate / 2)
        self.up_c_2_2 = Up(128, 64, 64, attn=True)
        self.drop = nn.Dropout2d(drop_rate)
        if normal_init:
            self.init_weights()

    def _basic_block(self, blk, x, H, W):
        """torchvision BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + downsample(x))"""
        y, Ho, Wo = _cba(x, H, W, blk.conv1, blk.bn1, ACT_RELU)
        idt = x if blk.downsample is None else _cba(x, H, W, blk.downsample[0], blk.downsample[1], ACT_NONE)[0]
        y, _, _ = _cba(y, Ho, Wo, blk.conv2, blk.bn2, ACT_RELU, idt)
        return y, Ho, Wo

    def _layer(self, layer, x, H, W):
        for blk in layer:
            x, H, W = self._basic_block(blk, x, H, W)
        return x, H, W

    @staticmethod
    def _head(seq, x, H, W, Ho, Wo):
        for m in seq:
            x, H, W = m.run(x, H, W)
        B = x.shape[0]
        return ops.ResizeACFn.apply(x, H, W, Ho, Wo).view(B, 1, Ho, Wo)      # one channel: NHWC == NCHW

    def forward_multi(self, imgs, domain_label, groups):
        """The mini-batches of `groups` datasets stacked along the batch axis in ONE pass (multi_train_TransFuse.py:151-172 runs one
        forward per dataset): every per-sample kernel runs once on all samples; BatchNorm batch statistics are taken per group of
        B/groups consecutive samples and the running statistics are updated group by group, exactly as consecutive forwards do."""
        with ops.bn_groups(groups):
            return self.forward(imgs, domain_label)

    def forward(self, imgs, domain_label, labels=None):
        B, _, Hi, Wi = imgs.shape
        if Hi % 16 or Wi % 16:
            raise ValueError("image sides must be multiples of 16")
        p, tr = self.drop.p, self.training
        r = self.resnet
        # bottom-up path: the DeiT tokens [B, (H/16)(W/16), 384] are already the NHWC map the reference builds by transpose + view
        h, w = Hi // 16, Wi // 16
        x_b = _drop2d(self.transformer(imgs, domain_label), p, tr)
        x_b_1 = _drop2d(self.up1.run(x_b, h, w), p, tr)
        x_b_2 = _drop2d(self.up2.run(x_b_1, 2 * h, 2 * w), p, tr)      # transformer pred supervise here
        # top-down path
        x_u, H, W = _cba(imgs.float(), Hi, Wi, r.conv1, r.bn1, ACT_RELU, nchw=True)
        x_u = ops.MaxPool3s2Fn.apply(x_u, H, W)
        H, W = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        x_u_2, H2, W2 = self._layer(r.layer1, x_u, H, W)
        x_u_2 = _drop2d(x_u_2, p, tr)
        x_u_1, H1, W1 = self._layer(r.layer2, x_u_2, H2, W2)
        x_u_1 = _drop2d(x_u_1, p, tr)
        x_u, H0, W0 = self._layer(r.layer3, x_u_1, H1, W1)
        x_u = _drop2d(x_u, p, tr)
        # joint path
        x_c = self.up_c.run(x_u, x_b, H0, W0)
        x_c_1_1 = self.up_c_1_1.run(x_u_1, x_b_1, H1, W1)
        x_c_1 = self.up_c_1_2.run(x_c, H0, W0, x_c_1_1)
        x_c_2_1 = self.up_c_2_1.run(x_u_2, x_b_2, H2, W2)
        x_c_2 = self.up_c_2_2.run(x_c_1, H1, W1, x_c_2_1)      # joint predict low supervise here
        # decoder part
        map_x = self._head(self.final_x, x_c, H0, W0, Hi, Wi)
        map_1 = self._head(self.final_1, x_b_2, H2, W2, Hi, Wi)
        map_2 = self._head(self.final_2, x_c_2, H2, W2, Hi, Wi)
        return map_x, map_1, map_2

    def init_weights(self):
        for m in (self.up1, self.up2, self.final_x, self.final_1, self.final_2, self.up_c, self.up_c_1_1, self.up_c_1_2, self.up_c_2_1,
                  self.up_c_2_2):
            m.apply(init_weights)
