"""Deterministic synthetic weights and inputs (numpy PCG64 — stable across machines), used by the
tests, the golden-vector generator and bench.py.  No reference or oracle dependency.

Shapes follow the reference's state_dict schema; `mdvit_schema()` rebuilds the
608-entry key->shape map of MDViT(adapt_method='Sup', decoder_name='MLPFM') from first principles.
"""
import zlib

import numpy as np
import torch

EMBED = (64, 128, 320, 512)
RATIO = (8, 8, 4, 4)
HEADS = 8
CRPE = ((3, 2), (5, 3), (7, 3))


def _stage_schema(s, prefix, C, r, sup):
    Ch = C // HEADS

    def shared(p):
        s[p + "cpe.proj.weight"] = (C, 1, 3, 3)
        s[p + "cpe.proj.bias"] = (C,)

    def crpe(p):
        for i, (w, hh) in enumerate(CRPE):
            s[f"{p}crpe.conv_list.{i}.weight"] = (hh * Ch, 1, w, w)
            s[f"{p}crpe.conv_list.{i}.bias"] = (hh * Ch,)

    shared(prefix + ".")
    crpe(prefix + ".")
    for j in range(2):
        b = f"{prefix}.mhca_blks.{j}."
        shared(b)
        s[b + "norm1.weight"] = (C,)
        s[b + "norm1.bias"] = (C,)
        a = b + "factoratt_crpe."
        s[a + "qkv.weight"] = (3 * C, C)
        s[a + "qkv.bias"] = (3 * C,)
        s[a + "proj.weight"] = (C, C)
        s[a + "proj.bias"] = (C,)
        if sup:
            hid = max(C // 2, 4)
            s[a + "domain_layer.0.weight"] = (hid, 4)
            s[a + "domain_layer.0.bias"] = (hid,)
            s[a + "domain_layer.2.weight"] = (C, hid)
            s[a + "domain_layer.2.bias"] = (C,)
        crpe(a)
        s[b + "norm2.weight"] = (C,)
        s[b + "norm2.bias"] = (C,)
        s[b + "mlp.fc1.weight"] = (r * C, C)
        s[b + "mlp.fc1.bias"] = (r * C,)
        s[b + "mlp.fc2.weight"] = (C, r * C)
        s[b + "mlp.fc2.bias"] = (C,)


def _bn(s, p, C):
    s[p + ".weight"] = (C,)
    s[p + ".bias"] = (C,)
    s[p + ".running_mean"] = (C,)
    s[p + ".running_var"] = (C,)
    s[p + ".num_batches_tracked"] = ()


def mdvit_schema(sup=True, aux=True):
    """key -> shape, in the reference's registration order (mdvit.py:509-599)."""
    s = {}
    s["stem.0.conv.weight"] = (32, 3, 3, 3)
    _bn(s, "stem.0.bn", 32)
    s["stem.1.conv.weight"] = (64, 32, 3, 3)
    _bn(s, "stem.1.bn", 64)
    cin = 64
    for i, C in enumerate(EMBED):
        p = f"patch_embed_stages.{i}.patch_conv"
        s[p + ".dwconv.weight"] = (cin, 1, 3, 3)
        s[p + ".pwconv.weight"] = (C, cin, 1, 1)
        _bn(s, p + ".bn", C)
        cin = C
    for i, C in enumerate(EMBED):
        _stage_schema(s, f"mhsa_stages.{i}", C, RATIO[i], sup)
    s["bridge.0.weight"] = (512, 512, 3, 3)
    s["bridge.0.bias"] = (512,)
    _bn(s, "bridge.1", 512)
    s["bridge.3.weight"] = (1024, 512, 3, 3)
    s["bridge.3.bias"] = (1024,)
    _bn(s, "bridge.4", 1024)
    chans = [(1024, 512, 3), (512, 320, 2), (320, 128, 1), (128, 64, 0)]
    for k, (ci, co, si) in enumerate(chans, start=1):
        p = f"decoder{k}"
        s[p + ".conv_before.weight"] = (co, ci, 1, 1)
        s[p + ".conv_before.bias"] = (co,)
        s[p + ".conv_after.dwconv.weight"] = (co, 2, 3, 3)
        s[p + ".conv_after.pwconv.weight"] = (co, co, 1, 1)
        _bn(s, p + ".conv_after.bn", co)
        _stage_schema(s, p + ".mhsa_block", co, RATIO[si], sup)
    s["finalconv.0.weight"] = (1, 64, 1, 1)
    s["finalconv.0.bias"] = (1,)
    if aux:
        for k in range(1, 5):
            p = f"debranch{k}"
            for i, C in enumerate(EMBED, start=1):
                s[f"{p}.linear{i}.weight"] = (512, C, 1, 1)
                s[f"{p}.linear{i}.bias"] = (512,)
            s[p + ".linear_fuse.0.weight"] = (512, 2112, 1, 1)
            s[p + ".linear_fuse.0.bias"] = (512,)
            _bn(s, p + ".linear_fuse.1", 512)
            s[p + ".linear_out.weight"] = (1, 512, 1, 1)
            s[p + ".linear_out.bias"] = (1,)
    return s


def canonical_key(k):
    """Shared CPE/CRPE tensors appear under alias keys; map alias -> owner key."""
    for j in ("0", "1"):
        k = k.replace(f".mhca_blks.{j}.cpe.", ".cpe.")
        k = k.replace(f".mhca_blks.{j}.factoratt_crpe.crpe.", ".crpe.")
    return k


def _rng(seed, key):
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def synth_tensor(key, shape, seed=0):
    """Well-conditioned deterministic value for one state_dict entry (activations stay O(1))."""
    key = canonical_key(key)
    g = _rng(seed, key)
    if key.endswith("num_batches_tracked"):
        return torch.zeros((), dtype=torch.long)
    n = lambda std: torch.from_numpy((g.standard_normal(shape) * std).astype(np.float32))  # noqa: E731
    if key.endswith("running_mean"):
        return n(0.1)
    if key.endswith("running_var"):
        return 1.0 + 0.2 * n(1.0).abs()
    leaf = key.rsplit(".", 1)[-1]
    is_norm = any(t in key for t in (".bn.", ".norm1.", ".norm2.", "bridge.1.", "bridge.4.", "linear_fuse.1."))
    if is_norm:
        return 1.0 + n(0.1) if leaf == "weight" else n(0.1)
    if leaf == "bias":
        return n(0.1)
    if len(shape) == 4 and shape[1] in (1, 2) and shape[2] > 1:      # depthwise / 2-per-group convs
        return n(0.5 / shape[2])
    if "domain_layer" in key:
        return n(0.7)
    if ".qkv." in key:
        return n(1.0 / np.sqrt(shape[1]))
    fan_in = int(np.prod(shape[1:]))
    return n(0.7 / np.sqrt(fan_in))


def synth_state_dict(seed=0, sup=True, aux=True):
    return {k: synth_tensor(k, shp, seed) for k, shp in mdvit_schema(sup, aux).items()}


def synth_batch(seed, dom, B, H=256, W=256):
    """Image ~ N(0,1) (the reference feeds ImageNet-normalised images, create_dataset.py:143-144) and a
    binary disk mask; deterministic in (seed, dom)."""
    g = _rng(seed, f"batch{dom}")
    img = torch.from_numpy(g.standard_normal((B, 3, H, W)).astype(np.float32))
    yy, xx = np.mgrid[0:H, 0:W]
    lab = np.zeros((B, 1, H, W), np.float32)
    for b in range(B):
        cy, cx = g.uniform(0.3, 0.7) * H, g.uniform(0.3, 0.7) * W
        r = g.uniform(0.15, 0.35) * min(H, W)
        lab[b, 0] = ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.float32)
    return img, torch.from_numpy(lab)


def dsn_perturb(state_dict, seed=7):
    """Make the domain-specific norm sets of an MDViT_DSN state_dict differ from each other (a fresh model has identical
    gamma=1 / beta=0 in every set, so the domain index would not change the output): deterministic in the key name."""
    out = {}
    for k, v in state_dict.items():
        if any(t in k for t in (".bns.", ".norm1s.", ".norm2s.", "bridge_norms1.", "bridge_norms2.")) and v.is_floating_point():
            g = _rng(seed, k)
            n = torch.from_numpy(g.standard_normal(tuple(v.shape)).astype(np.float32))
            if k.endswith("running_var"):
                v = 1.0 + 0.3 * n.abs()
            elif k.endswith("running_mean"):
                v = 0.2 * n
            elif k.endswith("weight"):
                v = 1.0 + 0.2 * n
            else:
                v = 0.2 * n
        out[k] = v.clone()
    return out
