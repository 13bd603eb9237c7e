"""CPU: pins oracle/mdvit_oracle.py (the torch fp32 restatement) to the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  Tolerances are fp32 round-off (the two differ only in op order)."""
import numpy as np
import torch
import torch.nn.functional as F

from mdvit_b200 import synth
from oracle import mdvit_oracle as O
from tests.helpers import fingerprint, oracle_state_dict


def test_eval_logits_64(golden):
    sd = oracle_state_dict()
    for d in range(4):
        img, _ = synth.synth_batch(1, d, 2, 64, 64)
        dl = F.one_hot(torch.full((2,), d), 4).float()
        with torch.no_grad():
            out, aux = O.mdvit_forward(sd, img, dl, str(d), training=False)
        np.testing.assert_allclose(out.numpy(), golden[f"eval64_out_{d}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(aux.numpy(), golden[f"eval64_aux_{d}"], rtol=0, atol=2e-5)


def test_eval_logits_256(golden):
    sd = oracle_state_dict()
    img, _ = synth.synth_batch(2, 3, 1, 256, 256)
    with torch.no_grad():
        out, aux = O.mdvit_forward(sd, img, torch.tensor([[0.0, 0, 0, 1]]), "3", training=False)
    np.testing.assert_allclose(out.numpy(), golden["eval256_out_3"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(aux.numpy(), golden["eval256_aux_3"], rtol=0, atol=5e-5)


def test_train_step_losses_grads_bn_adamw(golden):
    sd = oracle_state_dict(requires_grad=True)
    batches = [synth.synth_batch(1, d, 2, 64, 64) + (d,) for d in range(4)]
    losses, grads = O.train_step_grads(sd, batches)
    got = np.asarray([[a.item(), b.item(), c.item()] for a, b, c in losses["each"]])
    np.testing.assert_allclose(got, golden["train64_losses"], rtol=1e-5, atol=1e-6)
    names = [str(n) for n in golden["train64_grad_names"]]
    fp = fingerprint([(n, grads[n]) for n in names])
    ref = golden["train64_grad_fp"]
    scale = ref[:, 0].max()
    # l2 norms and probe projections agree up to fp32 accumulation noise (relative to the largest gradient)
    assert np.abs(fp - ref).max() / scale < 2e-4
    for key in golden.files:
        if key.startswith("train64_grad/"):
            g = grads[key.split("/", 1)[1]].detach().numpy()
            r = golden[key]
            assert np.abs(g - r).max() <= 2e-4 * max(np.abs(r).max(), 1e-6 * scale) + 1e-9
    bn_names = [str(n) for n in golden["train64_bn_names"]]
    np.testing.assert_allclose(fingerprint([(n, sd[n]) for n in bn_names]), golden["train64_bn_fp"], rtol=1e-5, atol=1e-5)
    assert int(sd["stem.0.bn.num_batches_tracked"]) == 4
    # AdamW (multi_train_MDViT.py:93-94)
    with torch.no_grad():
        after = []
        for n in names:
            p = sd[n].detach().clone()
            O.adamw_step(p, grads[n], torch.zeros_like(p), torch.zeros_like(p), step=1)
            after.append((n, p))
    ref_p = golden["train64_param_fp_after_adamw"]
    # at step 1 AdamW moves every weight by ~lr*sign(g); parameters whose true gradient is 0 (conv biases feeding a
    # batch-stat BN) have a round-off-noise sign, so their fingerprints may differ by ~lr*sqrt(numel)
    np.testing.assert_allclose(fingerprint(after), ref_p, rtol=2e-5, atol=5e-3)


def test_bce_clamp_and_dice_edge_cases():
    p = torch.tensor([0.0, 1.0, 0.5, 1.0])
    y = torch.tensor([1.0, 0.0, 1.0, 1.0])
    ref = torch.nn.BCELoss()(p, y)
    assert torch.allclose(O.bce_loss(p, y), ref)
    z = torch.zeros(4)
    assert abs(O.dice_loss(z, z).item()) < 1e-12      # (0+eps)/(0+eps) -> loss 0


def test_base_logits_64(golden):
    """BASE (base.py:340-512): the oracle with domain_label=None / with_aux=False against the unmodified reference BASE."""
    sd = {k: v.clone() for k, v in synth.synth_state_dict(0, sup=False, aux=False).items()}
    for k in list(sd):
        ck = synth.canonical_key(k)
        if ck != k:
            sd[k] = sd[ck]
    img, _ = synth.synth_batch(4, 0, 2, 64, 64)
    with torch.no_grad():
        out, aux = O.mdvit_forward(sd, img, None, None, training=False, with_aux=False)
        assert aux is None
        np.testing.assert_allclose(out.numpy(), golden["base_eval64_out"], rtol=0, atol=2e-5)
        out, _ = O.mdvit_forward(sd, img, None, None, training=True, with_aux=False)
        np.testing.assert_allclose(out.numpy(), golden["base_train64_out"], rtol=0, atol=5e-5)
