"""Dev: per-parameter gradient fingerprints of the DeepLabV3 aux decoder vs the reference golden."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdvit_b200 import synth
from mdvit_b200.model import MDViT
from oracle.make_golden_dsn import aux_state
from tests.helpers import fingerprint
g = np.load("tests/golden/mdvit_dsn_golden.npz")
dev = torch.device("cuda")
m = MDViT(img_size=256, adapt_method="Sup", num_domains=4, decoder_name="DeepLabV3")
m.load_state_dict(synth.synth_state_dict(0, aux=False) | aux_state(m, "debranch"), strict=True)
for k in range(1, 5):
    getattr(m, f"debranch{k}").classifier[0].project[3].p = 0.0
m = m.to(dev).train()
img, _ = synth.synth_batch(15, 1, 2, 256, 256)
dl = torch.nn.functional.one_hot(torch.full((2,), 1), 4).float().to(dev)
o, a = m(img.to(dev), dl, "1")
R = synth.synth_tensor("dl_probe", tuple(a.shape)).to(dev)
(a * R).sum().backward()
names = [str(n) for n in g["dl_grad_names"]]
P = dict(m.named_parameters())
fp = fingerprint([(n, P[n].grad) for n in names]); ref = g["dl_grad_fp"]
for i, n in enumerate(names):
    print(f"{n:55s} norm {fp[i,0]:.4e} ref {ref[i,0]:.4e} ({fp[i,0]/max(ref[i,0],1e-30):.3f})  probe diff/norm {abs(fp[i,1]-ref[i,1])/max(ref[i,0],1e-30):.3f}")
