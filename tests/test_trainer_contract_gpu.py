"""GPU (B200): the drop-in nn.Module under the call sequence of the reference trainer, with STOCK PyTorch around it.

The unmodified multi_train_MDViT.py cannot run where this model can (it needs /root/reference, which exists only in the
GPU-less build container) nor the model where the trainer can (the model has no CPU path), so this test restates, line by
line, everything train_val()/test() do TO the module and runs it on the drop-in with stock autograd,
torch.optim.AdamW, nn.BCELoss and torch.save/load — no MKDTrainer, no in-place gradient accumulation:

  model.cuda(); model.train()                                                   multi_train_MDViT.py:70,120
  output = model(img, domain_label, d) -> [out, aux]; sigmoid; BCELoss + dice    :139-169
  optimizer.zero_grad(); requires_grad=False on '*domain_layer*'; aux.backward(retain_graph=True)   :195-201
  requires_grad=True; (alpha*kt + (1-alpha)*seg).backward(); optimizer.step()   :203-213
  model.eval(); with torch.no_grad(): model(img, domain_label, d)               :239-263
  torch.save(model.state_dict()); model.load_state_dict(torch.load(...)) strict :114,332,352
"""
import io

import pytest
import torch
import torch.nn.functional as F

from mdvit_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda")


def dice_loss(score, target):        # Utils/losses.py:8-16
    smooth = 1e-5
    intersect = torch.sum(score * target)
    return 1 - (2 * intersect + smooth) / (torch.sum(score * score) + torch.sum(target * target) + smooth)


def build(dev):
    from mdvit_b200.model import MDViT
    m = MDViT(img_size=64, drop_rate=0.0, drop_path_rate=0.0, adapt_method="Sup", num_domains=4, decoder_name="MLPFM")
    m.load_state_dict(synth.synth_state_dict(0), strict=True)
    for k in range(1, 5):
        getattr(m, f"debranch{k}").dropout.p = 0.0
    return m.cuda()                                                               # :70


def trainer_step(model, optimizer, batches, alpha=0.5, hooks=None):
    criterion = [torch.nn.BCELoss(), dice_loss]                                   # :76
    datas_loss_list, aux_loss_list, kt_loss_list = [], [], []
    for img, label, set_id in batches:
        d = str(set_id[0].item())                                                 # :138
        domain_label = F.one_hot(set_id, 4).float().cuda()                        # :139
        output = model(img, domain_label, d)                                      # :142
        output, aux_out = torch.sigmoid(output[0]), torch.sigmoid(output[1])      # :147-149
        assert output.shape == label.shape and aux_out.shape == label.shape      # :152,160
        datas_loss_list.append(sum(f(output, label) for f in criterion))
        aux_loss_list.append(sum(f(aux_out, label) for f in criterion))
        kt_loss_list.append(dice_loss(aux_out, output))                           # :168
    multi_loss, multi_aux_loss, multi_kt_loss = sum(datas_loss_list), sum(aux_loss_list), sum(kt_loss_list)
    optimizer.zero_grad()                                                         # :195
    for name, params in model.named_parameters():                                 # :198-200
        if 'domain_layer' in name:
            params.requires_grad = False
    multi_aux_loss.backward(retain_graph=True)                                    # :201
    if hooks is not None:
        hooks["after_pass1"](model)
    for name, params in model.named_parameters():                                 # :203-205
        if 'domain_layer' in name:
            params.requires_grad = True
    (alpha * multi_kt_loss + (1 - alpha) * multi_loss).backward()                 # :206-207
    if hooks is not None:
        hooks["after_pass2"](model)
    optimizer.step()                                                              # :213
    return multi_loss.item(), multi_aux_loss.item(), multi_kt_loss.item()


def test_reference_trainer_call_sequence_on_the_dropin_module(dev):
    from mdvit_b200.train_step import MKDTrainer
    torch.manual_seed(0)
    model = build(dev)
    model.train()                                                                 # :120
    optimizer = torch.optim.AdamW(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-4, weight_decay=0.05)   # :93-94
    batches = []
    for d in range(4):
        img, lab = synth.synth_batch(21, d, 2, 64, 64)
        batches.append((img.cuda().float(), lab.cuda().float(), torch.full((2,), d, dtype=torch.long)))

    fired, seen = [], {}

    def after_pass1(m):
        # backward #1 must leave the domain adapter without gradients (multi_train_MDViT.py:198-201) ...
        seen["da_none"] = all(p.grad is None for n, p in m.named_parameters() if "domain_layer" in n)
        seen["fc_none"] = m.finalconv[0].weight.grad is None                       # ... and never reaches finalconv
        seen["trunk"] = m.stem[0].conv.weight.grad is not None and m.debranch1.linear_out.weight.grad is not None

    def after_pass2(m):
        seen["all"] = all(p.grad is not None and torch.isfinite(p.grad).all().item() for p in m.parameters())
        seen["grads"] = {n: p.grad.detach().clone() for n, p in m.named_parameters()}

    # autograd hooks fire (the default gradient path goes through AccumulateGrad; ADVICE r1: in-place accumulation is opt-in)
    h = model.stem[0].conv.weight.register_post_accumulate_grad_hook(lambda p: fired.append(1))
    l0 = trainer_step(model, optimizer, batches, hooks={"after_pass1": after_pass1, "after_pass2": after_pass2})
    h.remove()
    assert seen["da_none"] and seen["fc_none"] and seen["trunk"] and seen["all"]
    assert len(fired) == 2                                                         # once per backward pass
    # same gradients as the fused trainer's reference schedule on the same weights and batches
    m2 = build(dev).train()
    tr = MKDTrainer(m2, schedule="reference", fuse_domains=False)
    tr.grad.zero_()
    tr.backward(tr.forward_losses([(i, l, int(s[0])) for i, l, s in batches]))
    torch.cuda.synchronize()
    num = sum(float((seen["grads"][n] - p.grad).double().norm() ** 2) for n, p in m2.named_parameters())
    den = sum(float(p.grad.double().norm() ** 2) for _, p in m2.named_parameters())
    assert (num / den) ** 0.5 < 2e-2
    # a second step trains (loss moves) and stays finite
    l1 = trainer_step(model, optimizer, batches)
    assert all(map(lambda v: v == v and abs(v) < 1e4, l0 + l1)) and l1 != l0
    # ---- validation pass (:239-263): eval mode uses the BatchNorm running statistics, no autograd
    model.eval()
    img, lab, set_id = batches[1]
    with torch.no_grad():
        out_eval = model(img, F.one_hot(set_id, 4).float().cuda(), "1")
    assert out_eval[0].shape == lab.shape and not out_eval[0].requires_grad
    model.train()
    with torch.no_grad():
        out_train = model(img, F.one_hot(set_id, 4).float().cuda(), "1")
    assert (out_eval[0] - out_train[0]).abs().max().item() > 1e-3
    # ---- checkpoint round trip (:114,332,352): torch.save(state_dict) -> strict load into a fresh module -> same outputs
    model.eval()
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)
    buf.seek(0)
    fresh = build(dev)
    fresh.load_state_dict(torch.load(buf), strict=True)
    fresh.eval()
    with torch.no_grad():
        a = model(img, F.one_hot(set_id, 4).float().cuda(), "1")
        b = fresh(img, F.one_hot(set_id, 4).float().cuda(), "1")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # nn.DataParallel wraps the module when several GPUs are visible (:73-74): keys get a 'module.' prefix, nothing else
    sd = torch.nn.DataParallel(model).state_dict()
    assert len(sd) == 608 and all(k.startswith("module.") for k in sd)


def test_torch_autograd_grad_does_not_touch_param_grads(dev):
    """torch.autograd.grad(loss, params) returns the gradients and leaves .grad alone."""
    from mdvit_b200 import ops
    model = build(dev).train()
    img, lab = synth.synth_batch(3, 2, 2, 64, 64)
    out, aux = model(img.cuda(), F.one_hot(torch.full((2,), 2), 4).float().cuda(), "2")
    losses = ops.seg_losses(out, aux, lab.cuda())
    params = [model.stem[0].conv.weight, model.mhsa_stages[0].mhca_blks[0].mlp.fc1.weight, model.finalconv[0].weight]
    gs = torch.autograd.grad(losses.sum(), params)
    assert all(g is not None and g.abs().sum().item() > 0 for g in gs)
    assert all(p.grad is None for p in model.parameters())
