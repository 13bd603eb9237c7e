"""Development aid: per-C-ABI-call time of one stacked TransFuse_S_adapt train step (4 datasets x B images).
MDV_PROFILE=1 python scripts/dev_transfuse_profile.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MDV_PROFILE", "1")
from mdvit_b200 import _lib as L, ops, synth      # noqa: E402
from mdvit_b200.train_step import TransFuseTrainer      # noqa: E402
from mdvit_b200.transfuse import TransFuse_S_adapt      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
torch.manual_seed(0)
m = TransFuse_S_adapt(drop_rate=0.2).to(dev).train()
ops.manual_seed(1, dev)
tr = TransFuseTrainer(m)
batches = []
for d in range(4):
    img, lab = synth.synth_batch(3, d, B, 256, 256)
    batches.append((img.to(dev), lab.to(torch.uint8).to(dev), d))
for _ in range(2):
    tr.step(batches)
torch.cuda.synchronize()
L.PROFILE_LOG.clear()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
tr.step(batches)
e.record()
torch.cuda.synchronize()
print(f"step wall (eager, with profiling events): {s.elapsed_time(e):.2f} ms")
rep = L.profile_report(60, clear=False)
print(rep)
# by entry point
torch.cuda.synchronize()
agg = {}
for key, a, b in L.PROFILE_LOG:
    t = agg.setdefault(key[0], [0, 0.0])
    t[0] += 1
    t[1] += a.elapsed_time(b)
print("---- by entry point")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms  n={n:4d}  {k}")
