"""TEST INFRASTRUCTURE ONLY — never imported by the product path (mdvit_b200/).

Imports the *unmodified* reference (siyi-wind/MDViT) from /root/reference so that
golden vectors can be generated and the torch restatement in oracle/mdvit_oracle.py
can be pinned against it.  /root/reference exists only in the build container, so
nothing that runs on the GPU box (-m gpu tests, smoke(), bench.py) may call this.

The reference imports three packages that are absent from this image; they are
replaced by minimal stand-ins that follow the published semantics:
  * timm.models.layers.{DropPath, trunc_normal_, to_2tuple}  (mdvit.py:15, mpvit.py)
  * turtle.forward                                           (Decoders.py:5, stray import)
  * skimage.segmentation                                     (Utils/losses.py:5, unused here)
"""
import collections.abc
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MDVIT_REFERENCE_ROOT", "/root/reference")


class _DropPath(nn.Module):
    """timm semantics: one Bernoulli(keep) draw per sample, scaled by 1/keep; identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


def _to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return (x, x)


def _install(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Models", "Transformer"))


def install_stubs():
    if "timm" not in sys.modules:
        _install("timm")
        _install("timm.models")
        _install("timm.models.layers", DropPath=_DropPath,
                 trunc_normal_=nn.init.trunc_normal_, to_2tuple=_to_2tuple)
        _install("timm.models.registry", register_model=lambda f: f)
        _install("timm.models.helpers", load_pretrained=lambda *a, **k: None)
        _install("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
                 IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    if "turtle" not in sys.modules:
        _install("turtle", forward=None)
    if "skimage" not in sys.modules:
        sk = _install("skimage")
        sk.segmentation = _install("skimage.segmentation")


def load_reference():
    """Returns a namespace with the reference's MDViT, BASE and dice_loss."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    sys.dont_write_bytecode = True  # reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from Models.Transformer.mdvit import MDViT  # noqa
    from Models.Transformer.base import BASE  # noqa
    from Utils.losses import dice_loss  # noqa
    return types.SimpleNamespace(MDViT=MDViT, BASE=BASE, dice_loss=dice_loss)


def reference_step(model, batches, criterion_dice, alpha=0.5):
    """The reference's training-step math, multi_train_MDViT.py:129-207, minus logging.

    batches: list of (img, label, domain_index).  Returns dict of summed losses; grads are
    left on the model's parameters (optimizer.zero_grad / step are the caller's job).
    """
    bce = nn.BCELoss()
    seg_l, aux_l, kt_l = [], [], []
    for img, label, dom in batches:
        dl = torch.nn.functional.one_hot(torch.full((img.shape[0],), dom, dtype=torch.long), 4).float()
        out, aux = model(img, dl, str(dom))
        p, q = torch.sigmoid(out), torch.sigmoid(aux)
        seg_l.append(bce(p, label) + criterion_dice(p, label))
        aux_l.append(bce(q, label) + criterion_dice(q, label))
        kt_l.append(criterion_dice(q, p))
    seg, aux, kt = sum(seg_l), sum(aux_l), sum(kt_l)
    for n, p_ in model.named_parameters():
        if "domain_layer" in n:
            p_.requires_grad = False
    aux.backward(retain_graph=True)
    for n, p_ in model.named_parameters():
        if "domain_layer" in n:
            p_.requires_grad = True
    (alpha * kt + (1 - alpha) * seg).backward()
    return {"seg": seg.detach(), "aux": aux.detach(), "kt": kt.detach(),
            "seg_each": [x.detach() for x in seg_l], "aux_each": [x.detach() for x in aux_l],
            "kt_each": [x.detach() for x in kt_l]}
