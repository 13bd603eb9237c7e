"""Shared helpers for the parity tests (fingerprints identical to oracle/make_golden.py)."""
import zlib

import numpy as np
import torch

from mdvit_b200 import synth


def probe(name, numel):
    g = np.random.Generator(np.random.PCG64([77, zlib.crc32(name.encode())]))
    return torch.from_numpy(g.standard_normal(numel).astype(np.float32))


def fingerprint(named):
    rows = []
    for n, t in named:
        t = t.detach().double().flatten().cpu()
        rows.append([t.norm().item(), (t * probe(n, t.numel()).double()).sum().item()])
    return np.asarray(rows, np.float64)


def oracle_state_dict(device="cpu", requires_grad=False, seed=0):
    """synth weights as an oracle state_dict: aliases share storage, float leaves optionally require grad."""
    sd = {k: v.to(device).clone() for k, v in synth.synth_state_dict(seed).items()}
    if requires_grad:
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    for k in list(sd):
        ck = synth.canonical_key(k)
        if ck != k:
            sd[k] = sd[ck]
    return sd


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
