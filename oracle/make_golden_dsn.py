"""TEST INFRASTRUCTURE ONLY.  Golden logits of the UNMODIFIED reference MDViT_DSN (domain-specific norms,
Models/Transformer/mdvit.py:735-960) -> tests/golden/mdvit_dsn_golden.npz.   python oracle/make_golden_dsn.py

Weights: torch.manual_seed(0) + the stock constructor (mdvit_b200.model.MDViT_DSN reproduces them bit for bit), then
synth.dsn_perturb so that the four norm sets differ.  Dropout2d of the aux decoders off."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdvit_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import fingerprint  # noqa: E402


def aux_state(model, prefix="debranch"):
    """Deterministic weights for the auxiliary-decoder tensors of a model (synth_tensor handles any key/shape)."""
    return {k: synth.synth_tensor(k, tuple(v.shape)) for k, v in model.state_dict().items() if k.startswith(prefix)}


def mlp_aux_state(model):
    return aux_state(model, "debranch")


def main():
    ref_shim.load_reference()
    from Models.Transformer.mdvit import MDViT_DSN
    torch.manual_seed(0)
    m = MDViT_DSN(img_size=64, adapt_method="Sup", num_domains=4, decoder_name="MLPFM")
    out = {"keys": np.asarray(list(m.state_dict().keys())), "init_fp": fingerprint(list(m.named_parameters()))}
    m.load_state_dict(synth.dsn_perturb(m.state_dict()), strict=True)
    for k in range(1, 5):
        getattr(m, f"debranch{k}").dropout.p = 0.0
    with torch.no_grad():
        for mode in ("eval", "train"):
            m.train(mode == "train")
            for d in (1, 3):
                img, _ = synth.synth_batch(11, d, 2, 64, 64)
                dl = torch.nn.functional.one_hot(torch.full((2,), d), 4).float()
                o, a = m(img, dl, str(d))
                out[f"{mode}_out_{d}"], out[f"{mode}_aux_{d}"] = o.numpy(), a.numpy()
    # ---- MDViT with the 'MLP' auxiliary decoder (Decoders.MLPDecoder, Decoders.py:239-286; mdvit.py:607-611)
    from Models.Transformer.mdvit import MDViT
    m = MDViT(img_size=64, adapt_method="Sup", num_domains=4, decoder_name="MLP")
    m.load_state_dict(synth.synth_state_dict(0, aux=False) | {k: v for k, v in mlp_aux_state(m).items()}, strict=True)
    for k in range(1, 5):
        getattr(m, f"debranch{k}").dropout.p = 0.0
    out["mlp_keys"] = np.asarray(list(m.state_dict().keys()))
    with torch.no_grad():
        for mode in ("eval", "train"):
            m.train(mode == "train")
            img, _ = synth.synth_batch(12, 2, 2, 64, 64)
            dl = torch.nn.functional.one_hot(torch.full((2,), 2), 4).float()
            o, a = m(img, dl, "2")
            out[f"mlp_{mode}_out"], out[f"mlp_{mode}_aux"] = o.numpy(), a.numpy()
    # ---- MDViT with the 'Transformer' auxiliary decoder (mdvit.py:613-642,704-712): one extra 4-stage decoder per domain
    torch.manual_seed(0)
    m = MDViT(img_size=64, adapt_method="Sup", num_domains=4, decoder_name="Transformer")
    out["tr_keys"] = np.asarray(list(m.state_dict().keys()))
    out["tr_init_fp"] = fingerprint(list(m.named_parameters()))
    m.load_state_dict(synth.synth_state_dict(0, aux=False) | aux_state(m, "debranchs"), strict=True)
    with torch.no_grad():
        for mode in ("eval", "train"):
            m.train(mode == "train")
            for d in (0, 3):
                img, _ = synth.synth_batch(14, d, 2, 64, 64)
                dl = torch.nn.functional.one_hot(torch.full((2,), d), 4).float()
                o, a = m(img, dl, str(d))
                out[f"tr_{mode}_out_{d}"], out[f"tr_{mode}_aux_{d}"] = o.numpy(), a.numpy()
    # ---- BASE_DSN (Models/Transformer/base.py:515-696): the DSN trunk without auxiliary branches, Sup and plain attention
    from Models.Transformer.base import BASE_DSN
    for am in ("Sup", None):
        tag = "sup" if am else "plain"
        torch.manual_seed(0)
        m = BASE_DSN(img_size=64, adapt_method=am, num_domains=4)
        out[f"base_dsn_{tag}_keys"] = np.asarray(list(m.state_dict().keys()))
        out[f"base_dsn_{tag}_init_fp"] = fingerprint(list(m.named_parameters()))
        m.load_state_dict(synth.dsn_perturb(m.state_dict()), strict=True)
        with torch.no_grad():
            for mode in ("eval", "train"):
                m.train(mode == "train")
                for d in (0, 2):
                    img, _ = synth.synth_batch(13, d, 2, 64, 64)
                    dl = torch.nn.functional.one_hot(torch.full((2,), d), 4).float() if am else None
                    out[f"base_dsn_{tag}_{mode}_{d}"] = m(img, dl, str(d)).numpy()
    path = os.path.join(ROOT, "tests", "golden", "mdvit_dsn_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
