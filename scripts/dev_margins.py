import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests.test_model_gpu import _grads_of_step, grad_report
dev = torch.device("cuda")
g = np.load("tests/golden/mdvit_golden.npz")
for rep in range(3):
    m, tr, losses, grads = _grads_of_step(dev, "reference" if rep == 0 else "single_sweep")
    full = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("train64_grad/")}
    ga, wt, wl, text = grad_report(grads, full)
    names = [str(n) for n in g["train64_grad_names"]]; ref_fp = g["train64_grad_fp"]; gm = ref_fp[:, 0].max()
    dn = sorted((abs(grads[n].double().norm().item() - ref_fp[i, 0]) / max(ref_fp[i, 0], 1e-3 * gm), n) for i, n in enumerate(names))
    tight = max(d for d, n in dn if not ("bridge." in n or "domain_layer" in n)); loose = max(d for d, n in dn if ("bridge." in n or "domain_layer" in n))
    print(f"rep{rep}: global {ga:.4f} tight {wt:.3f} loose {wl:.3f} | norm-dev median {dn[len(dn)//2][0]:.4f} tight-max {tight:.3f} loose-max {loose:.3f} | loss err {np.abs(losses.cpu().numpy()-g['train64_losses']).max()/np.abs(g['train64_losses']).max():.4f}")
