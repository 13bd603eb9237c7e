"""Drop-in replacement for the reference's Models/Hybrid_models/TransFuseFolder/TransFuse.py (the `TransFuse_S_adapt` that
multi_train_TransFuse.py:66-67 builds): same constructor, forward(imgs, domain_label) -> (map_x, map_1, map_2), parameter names /
state_dict keys.  `structure_loss` (multi_train_TransFuse.py:29-38) is exported next to it.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mdvit_b200.ops import structure_loss  # noqa: E402,F401
from mdvit_b200.transfuse import (Attention_block, BiFusion_block, ChannelPool, Conv, DoubleConv, Residual, TransFuse_S_adapt, Up,  # noqa: E402,F401
                                  init_weights)
