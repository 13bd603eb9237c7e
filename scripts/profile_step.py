"""One eager training step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdvit_b200 import ops, synth
from mdvit_b200.model import MDViT
from mdvit_b200.train_step import MKDTrainer
B = int(os.environ.get("B", 32))
dev = torch.device("cuda")
torch.manual_seed(0)
model = MDViT(img_size=256, drop_rate=0.1, drop_path_rate=0.1, adapt_method="Sup", num_domains=4, decoder_name="MLPFM").to(dev).train()
tr = MKDTrainer(model)
batches = [tuple(t.to(dev) for t in synth.synth_batch(1234, d, B, 256, 256)) + (d,) for d in range(4)]
for _ in range(int(os.environ.get("WARM", 1))):
    tr.step(batches)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(batches)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
