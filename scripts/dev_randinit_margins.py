"""Developer aid: print the measured parity margins at the reference's random init (the numbers quoted in DESIGN.md section 7)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from tests import test_randinit_gpu as T
from tests.helpers import fingerprint
dev = torch.device("cuda")
rg = np.load(os.path.join(T.ROOT, "tests", "golden", "mdvit_randinit_golden.npz"))
print("== logits (domain, tensor, max/absmax, relL2, sign flips)")
for r in T.logits_errors(dev, rg):
    print("  ", r)
m, tr, losses, dice = T.run_trajectory(dev)
print("== trajectory: max |loss-ref|/max|ref|", np.abs(losses - rg["traj_losses"]).max() / np.abs(rg["traj_losses"]).max())
print("   per-entry rel", np.abs(losses / rg["traj_losses"] - 1).max(axis=(1, 2)))
print("   dice abs err per step", np.abs(dice - rg["traj_dice"]).max(axis=(1, 2)))
names = [str(n) for n in rg["param_names"]]
named = dict(m.named_parameters())
fp = fingerprint([(n, named[n]) for n in names]); ref = rg["traj_param_fp"]
en = np.abs(fp[:, 0] - ref[:, 0]) / np.maximum(ref[:, 0], 1e-6)
ep = np.abs(fp[:, 1] - ref[:, 1]) / np.maximum(ref[:, 0] * np.sqrt([named[n].numel() for n in names]), 1e-6)
print("   param fp: norm err max", en.max(), names[int(en.argmax())], "probe err max", ep.max(), names[int(ep.argmax())])
live = [not T.zero_grad_param(n) for n in names]
print("   zero-gradient biases excluded:", len(names) - sum(live), "max |p| there", max(float(named[n].abs().max()) for n, l in zip(names, live) if not l))
for e, n in sorted((e, n) for e, n, l in zip(np.maximum(en, ep), names, live) if l)[-4:]:
    print(f"      {e:.3e} {n}")
print("   worst non-bias:", sorted((e, n) for e, n, l in zip(np.maximum(en, ep), names, live) if l and not n.endswith(".bias"))[-3:])
for B in (int(os.environ.get("MARGIN_B", 32)),):
    losses, ref_l, rep = T.graph_step_vs_oracle(dev, B=B)
    print(f"== graph step B={B}: loss rel err", np.abs(losses / ref_l - 1).max(), "grads (global, worst tight, worst loose)", rep[:3])
    print("   ", rep[3])
