"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/mdvit_golden.npz by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_shim.py) in the build container.

    python oracle/make_golden.py

Weights and inputs come from mdvit_b200/synth.py (numpy PCG64, machine independent), so the file
only has to carry the reference's OUTPUTS: logits, losses, gradient fingerprints, BN running stats
and post-AdamW parameter fingerprints.  The reference's Dropout/DropPath/Dropout2d are disabled for
these runs (torch RNG streams cannot be reproduced by another implementation).
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdvit_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

FULL_GRADS = [
    "stem.0.conv.weight",
    "mhsa_stages.0.cpe.proj.weight",
    "mhsa_stages.0.crpe.conv_list.1.weight",
    "mhsa_stages.0.mhca_blks.0.factoratt_crpe.domain_layer.2.weight",
    "mhsa_stages.0.mhca_blks.0.factoratt_crpe.domain_layer.0.weight",
    "mhsa_stages.1.mhca_blks.1.norm1.weight",
    "decoder4.conv_after.dwconv.weight",
    "decoder4.mhsa_block.mhca_blks.1.mlp.fc2.bias",
    "finalconv.0.weight",
    "debranch3.linear_out.weight",
    "patch_embed_stages.1.patch_conv.bn.weight",
]


def probe(name, numel):
    g = np.random.Generator(np.random.PCG64([77, zlib.crc32(name.encode())]))
    return torch.from_numpy(g.standard_normal(numel).astype(np.float32))


def fingerprint(named):
    """[n,2] array: (l2 norm, dot with a fixed pseudo-random probe) per tensor, in the given order."""
    rows = []
    for n, t in named:
        t = t.detach().double().flatten()
        rows.append([t.norm().item(), (t * probe(n, t.numel()).double()).sum().item()])
    return np.asarray(rows, np.float64)


def build(ref, img):
    m = ref.MDViT(img_size=img, drop_rate=0.0, drop_path_rate=0.0, adapt_method="Sup", num_domains=4,
                  decoder_name="MLPFM")
    for k in range(1, 5):
        getattr(m, f"debranch{k}").dropout.p = 0.0
    m.load_state_dict(synth.synth_state_dict(0), strict=True)
    return m


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = ref_shim.load_reference()
    out = {}
    # ---- eval logits, 64x64, B=2, 4 domains
    m = build(ref, 64).eval()
    batches = [synth.synth_batch(1, d, 2, 64, 64) + (d,) for d in range(4)]
    with torch.no_grad():
        for img, lab, d in batches:
            dl = torch.nn.functional.one_hot(torch.full((2,), d), 4).float()
            o, a = m(img, dl, str(d))
            out[f"eval64_out_{d}"], out[f"eval64_aux_{d}"] = o.numpy(), a.numpy()
        # ---- eval logits 256x256, B=1, domain 3
        img, lab = synth.synth_batch(2, 3, 1, 256, 256)
        o, a = m(img, torch.tensor([[0.0, 0, 0, 1]]), "3")
        out["eval256_out_3"], out["eval256_aux_3"] = o.numpy(), a.numpy()
    # ---- one training step (multi_train_MDViT.py:129-213), train-mode BN
    m = build(ref, 64).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.05)
    opt.zero_grad()
    L = ref_shim.reference_step(m, batches, ref.dice_loss)
    out["train64_losses"] = np.asarray([[s.item(), a.item(), k.item()] for s, a, k in
                                        zip(L["seg_each"], L["aux_each"], L["kt_each"])], np.float64)
    named = [(n, p.grad) for n, p in m.named_parameters()]
    out["train64_grad_names"] = np.asarray([n for n, _ in named])
    out["train64_grad_fp"] = fingerprint(named)
    gd = dict(named)
    for n in FULL_GRADS:
        out["train64_grad/" + n] = gd[n].numpy()
    sd = m.state_dict()
    bn_keys = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
    out["train64_bn_names"] = np.asarray(bn_keys)
    out["train64_bn_fp"] = fingerprint([(k, sd[k]) for k in bn_keys])
    out["train64_bn/stem.0.bn.running_var"] = sd["stem.0.bn.running_var"].clone().numpy()   # clone: the buffer is updated in place later
    opt.step()
    out["train64_param_fp_after_adamw"] = fingerprint(list(m.named_parameters()))
    # train-mode logits of domain 1 after the step (BN batch stats, updated weights)
    with torch.no_grad():
        img, lab, d = batches[1]
        o, a = m(img, torch.nn.functional.one_hot(torch.full((2,), d), 4).float(), str(d))
        out["train64_out_after_1"], out["train64_aux_after_1"] = o.numpy(), a.numpy()
    # ---- BASE (base.py:340-512; BASELINE.json config 2): no DA, no aux branch — eval and train-mode logits, 64x64, B=2
    b = ref.BASE(img_size=64, drop_rate=0.0, drop_path_rate=0.0, adapt_method=False)
    b.load_state_dict(synth.synth_state_dict(0, sup=False, aux=False), strict=True)
    img, lab = synth.synth_batch(4, 0, 2, 64, 64)
    with torch.no_grad():
        out["base_eval64_out"] = b.eval()(img).numpy()
        out["base_train64_out"] = b.train()(img).numpy()
    path = os.path.join(ROOT, "tests", "golden", "mdvit_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()
