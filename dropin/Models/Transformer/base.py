"""Drop-in replacement for the reference's Models/Transformer/base.py (multi_train_BASE.py:67 imports BASE from it)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mdvit_b200.model import BASE, BASE_DSN  # noqa: E402,F401
