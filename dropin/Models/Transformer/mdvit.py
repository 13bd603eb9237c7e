"""Drop-in replacement for the reference's Models/Transformer/mdvit.py.

Put `<repo>/dropin` in front of the reference checkout on PYTHONPATH and the unmodified trainer's
`from Models.Transformer.mdvit import MDViT` (multi_train_MDViT.py:58) resolves to the B200 implementation:
same constructor, same forward(x, domain_label, d, out_feat, out_seg), same parameter names / state_dict keys.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mdvit_b200.model import MDViT, MDViT_DSN  # noqa: E402,F401
