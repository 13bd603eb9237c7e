"""Per-shape time table of one eager training step (MDV_PROFILE=1: CUDA events around every C-ABI call)."""
import os, sys
os.environ["MDV_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import _lib as L, ops, synth
from mdvit_b200.model import MDViT
from mdvit_b200.train_step import MKDTrainer
B = int(os.environ.get("B", 32))
dev = torch.device("cuda")
torch.manual_seed(0)
model = MDViT(img_size=256, drop_rate=0.1, drop_path_rate=0.1, adapt_method="Sup", num_domains=4, decoder_name="MLPFM").to(dev).train()
tr = MKDTrainer(model, **({"schedule": os.environ["SCHEDULE"]} if "SCHEDULE" in os.environ else {}))
batches = [tuple(t.to(dev) for t in synth.synth_batch(1234, d, B, 256, 256)) + (d,) for d in range(4)]
for _ in range(2):
    tr.step(batches)
torch.cuda.synchronize()
L.PROFILE_LOG.clear()
tr.step(batches)
print(L.profile_report(int(os.environ.get("TOP", 120))))
