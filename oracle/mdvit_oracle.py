"""TEST INFRASTRUCTURE ONLY — the parity oracle for the MDViT hot path.

A from-scratch, functional fp32 PyTorch restatement of the reference algorithm.  It takes a
plain ``state_dict`` (the reference's own key names) and tensors, and is
differentiable through torch autograd, so it is the checker for forward values AND gradients.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product (mdvit_b200/) never does.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.npz, which were
produced by the unmodified reference imported in the build container (oracle/make_golden.py).

Every function cites the reference lines it restates (paths relative to the reference root).
"""
import math

import torch
import torch.nn.functional as F

NUM_HEADS = 8
CRPE_WINDOWS = ((3, 2), (5, 3), (7, 3))  # (window, heads) — mdvit.py:423


# ----------------------------------------------------------------------------- small pieces
def hardswish(x):
    return x * F.relu6(x + 3.0) / 6.0


def batchnorm(sd, prefix, x, training, momentum=0.1, eps=1e-5, update=True):
    """nn.BatchNorm2d on NCHW (mpvit.py:107; mdvit.py:101,559,562; Decoders.py:40,306).

    Train: batch mean / biased var; running <- (1-m) running + m batch (unbiased var)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if update:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                rm.mul_(1 - momentum).add_(momentum * mean.detach())
                rv.mul_(1 - momentum).add_(momentum * var.detach() * n / max(n - 1, 1))
                key = prefix + ".num_batches_tracked"
                if key in sd:
                    sd[key] += 1
    else:
        mean, var = rm, rv
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + eps)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def bilinear(x, size):
    """nn.functional.interpolate(mode='bilinear', align_corners=False) (mdvit.py:699, Decoders.py:196,319-336)."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


# ----------------------------------------------------------------------------- transformer block
def conv_pos_enc(sd, prefix, x, H, W):
    """ConvPosEnc.forward, mpvit.py:239-248: tokens -> image, depthwise 3x3 (+bias) + identity."""
    B, N, C = x.shape
    feat = x.transpose(1, 2).reshape(B, C, H, W)
    y = F.conv2d(feat, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"], padding=1, groups=C) + feat
    return y.flatten(2).transpose(1, 2)


def conv_rel_pos_enc(sd, prefix, q, v, H, W):
    """ConvRelPosEnc.forward, mpvit.py:296-318. q, v: [B,h,N,Ch]. Returns q * dwconv(v)."""
    B, h, N, Ch = q.shape
    v_img = v.permute(0, 1, 3, 2).reshape(B, h * Ch, H, W)
    outs, c0 = [], 0
    for i, (win, heads) in enumerate(CRPE_WINDOWS):
        c1 = c0 + heads * Ch
        outs.append(F.conv2d(v_img[:, c0:c1], sd[f"{prefix}.conv_list.{i}.weight"],
                             sd[f"{prefix}.conv_list.{i}.bias"], padding=win // 2, groups=heads * Ch))
        c0 = c1
    conv_v = torch.cat(outs, dim=1).reshape(B, h, Ch, N).permute(0, 1, 3, 2)
    return q * conv_v


def domain_gate(sd, prefix, domain_label, h):
    """DA head gate, mdvit.py:272-276,301-303: MLP(onehot) -> [B,h,Ch] -> softmax over heads."""
    z = F.linear(domain_label, sd[prefix + ".domain_layer.0.weight"], sd[prefix + ".domain_layer.0.bias"])
    z = F.linear(F.relu(z), sd[prefix + ".domain_layer.2.weight"], sd[prefix + ".domain_layer.2.bias"])
    B, C = z.shape
    return torch.softmax(z.reshape(B, h, C // h), dim=1)  # [B,h,Ch]


def factor_attention(sd, prefix, crpe_prefix, x, H, W, domain_label, drop_p=0.0, training=False):
    """FactorAtt_ConvRelPosEnc_Sup.forward (mdvit.py:281-313) / FactorAtt_ConvRelPosEnc.forward
    (mpvit.py:347-373) when domain_label is None."""
    B, N, C = x.shape
    h = NUM_HEADS
    Ch = C // h
    qkv = F.linear(x, sd[prefix + ".qkv.weight"], sd[prefix + ".qkv.bias"])
    qkv = qkv.reshape(B, N, 3, h, Ch).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]                       # [B,h,N,Ch]
    k_soft = k.softmax(dim=2)                              # over tokens
    ktv = torch.einsum("bhnk,bhnv->bhkv", k_soft, v)
    fa = torch.einsum("bhnk,bhkv->bhnv", q, ktv)
    fa = (Ch ** -0.5) * fa + conv_rel_pos_enc(sd, crpe_prefix, q, v, H, W)
    if domain_label is not None:
        g = domain_gate(sd, prefix, domain_label, h)       # [B,h,Ch]
        fa = g[:, :, None, :] * fa
    y = fa.transpose(1, 2).reshape(B, N, C)
    y = F.linear(y, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])
    return F.dropout(y, drop_p, training)


def attention_sup(sd, prefix, x, domain_label, num_heads, return_pre_proj=False):
    """Attention_Sup.forward, Models/Hybrid_models/TransFuseFolder/vision_transformer.py:149-169 (TransFuse_S_adapt's DeiT
    branch; plain Attention :110-122 when domain_label is None): softmax(q k^T * scale) v, the softmax-over-heads DA gate,
    then proj.  attn_drop / proj_drop are 0 in the TransFuse trainer's configuration."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[prefix + ".qkv.weight"], sd.get(prefix + ".qkv.bias")).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
    y = attn @ v                                                    # [B,h,N,hd]
    if domain_label is not None:
        y = domain_gate(sd, prefix, domain_label, num_heads)[:, :, None, :] * y
    y = y.transpose(1, 2).reshape(B, N, C)
    out = F.linear(y, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])
    return (out, y) if return_pre_proj else out


def mlp(sd, prefix, x, drop_p=0.0, training=False):
    """Mlp.forward, mpvit.py:71-78 (GELU = exact erf)."""
    x = F.gelu(F.linear(x, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))
    x = F.dropout(x, drop_p, training)
    x = F.linear(x, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])
    return F.dropout(x, drop_p, training)


def drop_path(x, p, training):
    """timm DropPath: per-sample Bernoulli keep mask scaled 1/keep (mdvit.py:339)."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep) / keep
    return x * m


def serial_block(sd, stage, j, x, H, W, domain_label, drop=0.0, dpr=0.0, training=False):
    """SerialBlock_adapt.forward, mdvit.py:346-361 (LayerNorm eps 1e-6, mdvit.py:498)."""
    blk = f"{stage}.mhca_blks.{j}"
    C = x.shape[-1]
    x = conv_pos_enc(sd, f"{stage}.cpe", x, H, W)
    cur = F.layer_norm(x, (C,), sd[blk + ".norm1.weight"], sd[blk + ".norm1.bias"], 1e-6)
    cur = factor_attention(sd, blk + ".factoratt_crpe", f"{stage}.crpe", cur, H, W, domain_label, drop, training)
    x = x + drop_path(cur, dpr, training)
    cur = F.layer_norm(x, (C,), sd[blk + ".norm2.weight"], sd[blk + ".norm2.bias"], 1e-6)
    cur = mlp(sd, blk + ".mlp", cur, drop, training)
    return x + drop_path(cur, dpr, training)


def mhsa_stage(sd, stage, x, H, W, domain_label, num_layers=2, **kw):
    """MHSA_stage_adapt.forward, mdvit.py:437-440."""
    for j in range(num_layers):
        x = serial_block(sd, stage, j, x, H, W, domain_label, **kw)
    return x


# ----------------------------------------------------------------------------- conv pieces
def conv_bn_act(sd, prefix, x, stride, pad, training, act):
    """mpvit.Conv2d_BN.forward, mpvit.py:117-124 (conv no bias -> BN -> act)."""
    x = F.conv2d(x, sd[prefix + ".conv.weight"], None, stride=stride, padding=pad)
    return act(batchnorm(sd, prefix + ".bn", x, training))


def patch_embed(sd, prefix, x, stride, training):
    """DWCPatchEmbed / mdvit.DWConv2d_BN.forward, mdvit.py:114-123: dw3x3(groups=in) -> pw1x1 -> BN -> Hardswish."""
    p = prefix + ".patch_conv"
    x = F.conv2d(x, sd[p + ".dwconv.weight"], None, stride=stride, padding=1, groups=x.shape[1])
    x = F.conv2d(x, sd[p + ".pwconv.weight"], None)
    return hardswish(batchnorm(sd, p + ".bn", x, training))


def decoder_block(sd, prefix, x, skip, domain_label, training, **kw):
    """UnetDecodingBlockTransformer.forward (use_res=False), Decoders.py:194-214, with
    Decoders.DWConv2d_BN (groups=out_ch over 2*out_ch inputs), Decoders.py:54-63."""
    H, W = skip.shape[2:]
    out = bilinear(x, (H, W))
    out = F.conv2d(out, sd[prefix + ".conv_before.weight"], sd[prefix + ".conv_before.bias"])
    out = torch.cat((skip, out), dim=1)
    Cout = skip.shape[1]
    out = F.conv2d(out, sd[prefix + ".conv_after.dwconv.weight"], None, padding=1, groups=Cout)
    out = F.conv2d(out, sd[prefix + ".conv_after.pwconv.weight"], None)
    out = hardswish(batchnorm(sd, prefix + ".conv_after.bn", out, training))
    B = out.shape[0]
    tok = out.flatten(2).transpose(1, 2)
    tok = mhsa_stage(sd, prefix + ".mhsa_block", tok, H, W, domain_label, training=training, **kw)
    return tok.transpose(1, 2).reshape(B, Cout, H, W)


def mlp_decoder_fm(sd, prefix, feats, img_size, training, drop2d=0.1):
    """MLPDecoderFM.forward, Decoders.py:315-339."""
    x1, x2, x3, x4, x5 = feats
    h, w = x1.shape[2:]
    ups = []
    for i, xi in enumerate((x1, x2, x3, x4), start=1):
        y = F.conv2d(xi, sd[f"{prefix}.linear{i}.weight"], sd[f"{prefix}.linear{i}.bias"])
        ups.append(bilinear(y, (h, w)))
    out = torch.cat(ups + [x5], dim=1)
    out = F.conv2d(out, sd[prefix + ".linear_fuse.0.weight"], sd[prefix + ".linear_fuse.0.bias"])
    out = F.relu(batchnorm(sd, prefix + ".linear_fuse.1", out, training))
    out = F.dropout2d(out, drop2d, training)
    out = bilinear(out, img_size)
    return F.conv2d(out, sd[prefix + ".linear_out.weight"], sd[prefix + ".linear_out.bias"])


# ----------------------------------------------------------------------------- whole model
def mdvit_forward(sd, x, domain_label=None, d=None, training=False, drop=0.0, dpr=0.0, drop2d=0.0,
                  with_aux=True, return_feats=False):
    """MDViT.forward, mdvit.py:667-730 (decoder_name='MLPFM').  BASE.forward (base.py:477-512) is the
    same with with_aux=False and domain_label=None.  `sd` is mutated (BN running stats) when training."""
    kw = dict(drop=drop, dpr=dpr)
    img_size = x.shape[2:]
    x = conv_bn_act(sd, "stem.0", x, 2, 1, training, hardswish)
    x = conv_bn_act(sd, "stem.1", x, 2, 1, training, hardswish)
    enc = []
    for i in range(4):
        x = patch_embed(sd, f"patch_embed_stages.{i}", x, 1 if i == 0 else 2, training)
        B, C, H, W = x.shape
        tok = x.flatten(2).transpose(1, 2)
        tok = mhsa_stage(sd, f"mhsa_stages.{i}", tok, H, W, domain_label, training=training, **kw)
        x = tok.transpose(1, 2).reshape(B, C, H, W)
        enc.append(x)
    # bridge, mdvit.py:557-564,687
    out = F.conv2d(enc[3], sd["bridge.0.weight"], sd["bridge.0.bias"], padding=1)
    out = F.relu(batchnorm(sd, "bridge.1", out, training))
    out = F.conv2d(out, sd["bridge.3.weight"], sd["bridge.3.bias"], padding=1)
    out = F.relu(batchnorm(sd, "bridge.4", out, training))
    for k, skip in zip((1, 2, 3, 4), (enc[3], enc[2], enc[1], enc[0])):
        out = decoder_block(sd, f"decoder{k}", out, skip, domain_label, training, **kw)
    dec4 = out
    out = bilinear(out, img_size)
    out = F.conv2d(out, sd["finalconv.0.weight"], sd["finalconv.0.bias"])
    aux = None
    if with_aux and d in ("0", "1", "2", "3"):
        aux = mlp_decoder_fm(sd, f"debranch{int(d) + 1}", enc + [dec4], img_size, training, drop2d)
    if return_feats:
        return out, aux, enc, dec4
    return out, aux


# ----------------------------------------------------------------------------- losses / step
def dice_loss(score, target):
    """Utils/losses.py:8-16."""
    smooth = 1e-5
    inter = torch.sum(score * target)
    return 1 - (2 * inter + smooth) / (torch.sum(score * score) + torch.sum(target * target) + smooth)


def bce_loss(p, y):
    """nn.BCELoss(mean) (multi_train_MDViT.py:76): log clamped >= -100 in the forward, and torch's own backward
    (p - y) / max(p (1 - p), 1e-12) / n, which stays finite at exactly saturated sigmoids (the hand-written
    -(y log p + (1-y) log(1-p)) formula back-propagates 0 * inf = nan there; at the reference's random init most sigmoids
    ARE saturated)."""
    with torch.autocast(device_type=p.device.type, enabled=False):      # (BCELoss refuses to run under autocast)
        return F.binary_cross_entropy(p.float(), y.float())


def seg_losses(out, aux, label):
    """multi_train_MDViT.py:147-169: (L_seg, L_aux, L_kt) for one domain batch."""
    p, q = torch.sigmoid(out), torch.sigmoid(aux)
    return bce_loss(p, label) + dice_loss(p, label), bce_loss(q, label) + dice_loss(q, label), dice_loss(q, p)


def is_da(name):
    return "domain_layer" in name


def train_step_grads(sd, batches, alpha=0.5, **fw):
    """multi_train_MDViT.py:129-207.  sd values that are float tensors must have requires_grad=True
    (leaf).  Returns (losses, grads) where DA params get only the grad of alpha*kt+(1-alpha)*seg."""
    params = {k: v for k, v in sd.items() if v.requires_grad}
    seg_s = aux_s = kt_s = 0.0
    each = []
    for img, label, dom in batches:
        dl = F.one_hot(torch.full((img.shape[0],), dom, dtype=torch.long, device=img.device), 4).float()
        out, aux = mdvit_forward(sd, img, dl, str(dom), training=True, **fw)
        ls, la, lk = seg_losses(out, aux, label)
        seg_s, aux_s, kt_s = seg_s + ls, aux_s + la, kt_s + lk
        each.append((ls.detach(), la.detach(), lk.detach()))
    names = list(params)
    g_aux = torch.autograd.grad(aux_s, [params[n] for n in names], retain_graph=True, allow_unused=True)
    g_uni = torch.autograd.grad(alpha * kt_s + (1 - alpha) * seg_s, [params[n] for n in names], allow_unused=True)
    grads = {}
    for n, ga, gu in zip(names, g_aux, g_uni):
        tot = None
        if gu is not None:
            tot = gu
        if ga is not None and not is_da(n):
            tot = ga if tot is None else tot + ga
        grads[n] = tot
    return {"seg": seg_s.detach(), "aux": aux_s.detach(), "kt": kt_s.detach(), "each": each}, grads


def adamw_step(p, g, m, v, step, lr=1e-4, wd=0.05, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.AdamW single-tensor math (multi_train_MDViT.py:93-94)."""
    p.mul_(1 - lr * wd)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
