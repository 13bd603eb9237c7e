"""TEST INFRASTRUCTURE ONLY — never imported by the product path (mdvit_b200/).

CPU restatement (plain torch fp32) of what each C-ABI Function of TransFuse_S_adapt computes, on the same NHWC maps and with the
same call signatures as ops.ConvBnActFn / BnActFn / MaxPool3s2Fn / ResizeACFn / GateCatFn / ChannelPoolFn, plus the DeiT-S-adapt
forward over a module's parameters and structure_loss.  Every piece is the reference's own torch call at the cited line:
  nn.Conv2d + nn.BatchNorm2d (+ `out += identity`) + nn.ReLU   TransFuse.py:574-650, torchvision BasicBlock
  nn.MaxPool2d(3, 2, 1)                                          TransFuse.py:234
  F.interpolate / nn.Upsample(bilinear, align_corners=True)      TransFuse.py:559, 262-264
  ChannelPool, sigmoid gates, torch.cat([g, x, bp], 1)           TransFuse.py:20-22, 63-73, 620
  DeiT-S-adapt blocks (softmax(QK^T)V, DA head gate)             vision_transformer.py:125-211, 322-389; DeiT.py:116-139
  structure_loss                                                 multi_train_TransFuse.py:29-38
PINNED: with these in place of the kernels, mdvit_b200.transfuse.TransFuse_S_adapt reproduces the goldens of the UNMODIFIED
reference (oracle/make_golden_transfuse_model.py: maps 2e-4, losses 2e-5, 413 gradient norms) — tests/test_transfuse_wiring.py.
The GPU tests use the same classes as the torch-fp32 reference of each kernel and, with TF32 allowed, as the "stock PyTorch on
this GPU" yardstick."""
import torch
import torch.nn.functional as F

from mdvit_b200 import ops
from oracle.make_golden_transfuse_model import structure_loss_ref      # noqa: F401  (multi_train_TransFuse.py:29-38)


def emu_bn(z, bufs, gamma, beta, training):
    """nn.BatchNorm2d on NCHW; under ops.bn_groups(G): G consecutive calls on the G batch chunks (what the grouped kernels compute)"""
    G = ops.current_bn_groups() if training else 1
    out = torch.cat([F.batch_norm(c, bufs[0], bufs[1], gamma, beta, training, 0.1, 1e-5) for c in z.chunk(G)])
    if training:
        bufs[2].add_(G)
    return out


class EmuConv:
    @staticmethod
    def apply(x, w, cbias, gamma, beta, residual, bufs, B, H, W, stride, act, training, nchw):
        k = w.shape[2]
        xin = x if nchw else x.view(B, H, W, -1).permute(0, 3, 1, 2)
        z = F.conv2d(xin, w, cbias, stride, (k - 1) // 2)
        if gamma is not None:
            z = emu_bn(z, bufs, gamma, beta, training)
        y = z.permute(0, 2, 3, 1).reshape(B, -1, w.shape[0])
        if gamma is None and act == ops.ACT_RELU:
            y = torch.relu(y)
        if residual is not None:
            y = y + residual
        if gamma is not None and act == ops.ACT_RELU:
            y = torch.relu(y)
        return y


class EmuBn:
    @staticmethod
    def apply(x, gamma, beta, bufs, act, training):
        y = emu_bn(x.transpose(1, 2).unsqueeze(-1), bufs, gamma, beta, training).squeeze(-1).transpose(1, 2)
        return torch.relu(y) if act == ops.ACT_RELU else y


class EmuPool:
    @staticmethod
    def apply(x, H, W):
        B, _, C = x.shape
        y = F.max_pool2d(x.view(B, H, W, C).permute(0, 3, 1, 2), 3, 2, 1)
        return y.permute(0, 2, 3, 1).reshape(B, -1, C)


class EmuResize:
    @staticmethod
    def apply(x, H, W, Ho, Wo):
        B, _, C = x.shape
        y = F.interpolate(x.view(B, H, W, C).permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=True)
        return y.permute(0, 2, 3, 1).reshape(B, -1, C)


class EmuGateCat:
    @staticmethod
    def apply(g, p, x, v, bp):
        parts = [g * p]
        if x is not None:
            parts.append(x * v.unsqueeze(1))
        if bp is not None:
            parts.append(bp)
        return torch.cat(parts, dim=2) if len(parts) > 1 else parts[0]


class EmuChannelPool:
    @staticmethod
    def apply(x):
        return torch.cat((x.max(dim=2, keepdim=True)[0], x.mean(dim=2, keepdim=True)), dim=2)


def deit_forward_torch(tr, imgs, label):
    """plain-torch DeiT-S-adapt forward over the module's parameters (vision_transformer.py:125-211,322-389; DeiT.py:116-139)"""
    x = F.conv2d(imgs, tr.patch_embed.proj.weight, tr.patch_embed.proj.bias, stride=16).flatten(2).transpose(1, 2) + tr.pos_embed
    for blk in tr.blocks:
        a = blk.attn
        B, N, C = x.shape
        h = F.layer_norm(x, (C,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv = a.qkv(h).reshape(B, N, 3, a.num_heads, C // a.num_heads).permute(2, 0, 3, 1, 4)
        att = ((qkv[0] @ qkv[1].transpose(-2, -1)) * a.scale).softmax(dim=-1)
        o = att @ qkv[2]                                                              # [B, heads, N, 64]
        gate = a.domain_layer(label).reshape(B, a.num_heads, 1, C // a.num_heads).softmax(dim=1)
        o = (o * gate).transpose(1, 2).reshape(B, N, C)
        x = x + a.proj(o)
        h = F.layer_norm(x, (C,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        x = x + blk.mlp.fc2(F.gelu(blk.mlp.fc1(h)))
    return F.layer_norm(x, (x.shape[-1],), tr.norm.weight, tr.norm.bias, tr.norm.eps)
