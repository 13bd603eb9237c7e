"""TEST INFRASTRUCTURE ONLY.  A 5-step training trajectory of the UNMODIFIED reference TransFuse_S_adapt under the loop of
multi_train_TransFuse.py:145-197 (per dataset: forward, 0.5/0.3/0.2 structure_loss deep supervision; summed; zero_grad; backward;
AdamW) at its own random init (seed 0, stock ctor, drop_rate=0), 2 datasets x 2 images of 256 x 256:
per-step per-dataset losses and hard Dice of the joint prediction (sigmoid(map_2) > 0.5), into
tests/golden/transfuse_traj_golden.npz.

    python oracle/make_golden_transfuse_traj.py

LR is 10x the reference config's 1e-4 (Configs/multi_train_local.yml:25) so that five steps move the loss visibly."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden_randinit import hard_dice      # noqa: E402
from oracle.make_golden_transfuse_model import case, structure_loss_ref      # noqa: E402

STEPS, LR, WD, DOMAINS = 5, 1e-3, 0.05, (1, 3)


def batches():
    """2 datasets x 2 images: (img, mask, domain index)"""
    img, mask, _ = case(B=4)
    return [(img[0:2], mask[0:2], DOMAINS[0]), (img[2:4], mask[2:4], DOMAINS[1])]


def main():
    from oracle import ref_shim
    ref_shim.install_stubs()
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from Models.Hybrid_models.TransFuseFolder.TransFuse import TransFuse_S_adapt
    torch.manual_seed(0)
    m = TransFuse_S_adapt(drop_rate=0.0, pretrained=False, num_domains=4).train()
    opt = torch.optim.AdamW(m.parameters(), lr=LR, weight_decay=WD)
    bs = batches()
    losses, dices = [], []
    for step in range(STEPS):
        ls, ds = [], []
        for img, mask, d in bs:
            dl = F.one_hot(torch.full((img.shape[0],), d), 4).float()
            map_x, map_1, map_2 = m(img, dl)      # (lateral_map_4, lateral_map_3, lateral_map_2), multi_train_TransFuse.py:166
            ls.append(0.5 * structure_loss_ref(map_2, mask) + 0.3 * structure_loss_ref(map_1, mask) + 0.2 * structure_loss_ref(map_x, mask))
            ds.append(hard_dice(map_2.detach(), mask))
        opt.zero_grad()
        sum(ls).backward()
        opt.step()
        losses.append([l.item() for l in ls])
        dices.append(ds)
        print("step", step, losses[-1], dices[-1], flush=True)
    path = os.path.join(ROOT, "tests", "golden", "transfuse_traj_golden.npz")
    np.savez_compressed(path, losses=np.asarray(losses, np.float64), dice=np.asarray(dices, np.float64), lr=np.asarray([LR, WD]))
    print("wrote", path)


if __name__ == "__main__":
    main()
