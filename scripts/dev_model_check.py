"""Dev check (GPU): mdvit_b200.MDViT vs oracle (torch fp32 on the same GPU, TF32 off) — stage-by-stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from mdvit_b200 import synth, ops
from mdvit_b200.model import MDViT
from oracle import mdvit_oracle as O

dev = torch.device("cuda")
IMG = int(os.environ.get("IMG", 64)); B = int(os.environ.get("B", 2))

def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()

ssd = synth.synth_state_dict(0)
m = MDViT(img_size=IMG, adapt_method='Sup', num_domains=4, decoder_name='MLPFM').to(dev)
m.load_state_dict(ssd, strict=True)
for k in range(1, 5): getattr(m, f'debranch{k}').dropout.p = 0.0

def oracle_sd():
    sd = {k: v.to(dev).clone() for k, v in ssd.items()}
    for k in list(sd):
        ck = synth.canonical_key(k)
        if ck != k: sd[k] = sd[ck]
    return sd

for training in (False, True):
    m.train(training)
    sd = oracle_sd()
    img, lab = synth.synth_batch(1, 2, B, IMG, IMG)
    img, lab = img.to(dev), lab.to(dev)
    dl = F.one_hot(torch.full((B,), 2), 4).float().to(dev)
    with torch.no_grad():
        ro, ra, renc, rdec4 = O.mdvit_forward(sd, img, dl, '2', training=training, return_feats=True)
        enc = m._trunk_forward(img, dl)
        for i, (t, H, W) in enumerate(enc):
            r = renc[i].flatten(2).transpose(1, 2)
            print(f"train={training} enc{i}: rel {rel(t, r):.3e}  absmax {r.abs().max().item():.3f}")
        dec4, h, w = m._decode(enc, dl)
        print(f"train={training} dec4: rel {rel(dec4, rdec4.flatten(2).transpose(1, 2)):.3e}")
        out = m._head(dec4, h, w, img.shape[2:])
        print(f"train={training} out: rel {rel(out, ro):.3e} absmax {ro.abs().max().item():.3f}")
        aux = m.debranch3([e[0] for e in enc] + [dec4], [(e[1], e[2]) for e in enc], img.shape[2:])
        print(f"train={training} aux: rel {rel(aux, ra):.3e} absmax {ra.abs().max().item():.3f}")
    if training:
        msd = m.state_dict()
        worst = max(((msd[k].float() - sd[k].float()).abs().max().item(), k) for k in msd if 'running' in k)
        print("BN running worst abs diff", worst, "nbt", msd['stem.0.bn.num_batches_tracked'].item())
        m.load_state_dict(ssd, strict=True)

# ---------------- gradients of one training step (4 domains, MKD two-pass backward)
m.train(True)
m.load_state_dict(ssd, strict=True)
sd = oracle_sd()
for k, v in sd.items():
    if v.is_floating_point() and 'running' not in k: v.requires_grad_(True)
for k in list(sd):
    ck = synth.canonical_key(k)
    if ck != k: sd[k] = sd[ck]
batches = [tuple(t.to(dev) for t in synth.synth_batch(1, d, B, IMG, IMG)) + (d,) for d in range(4)]
t0 = time.time()
Lr, gr = O.train_step_grads(sd, batches)
torch.cuda.synchronize(); print("oracle step s", time.time() - t0)
ops.reset_stream_ids()
seg = aux_l = kt = 0
for img, lab, d in batches:
    dl = F.one_hot(torch.full((B,), d), 4).float().to(dev)
    o, a = m(img, dl, str(d))
    l = ops.seg_losses(o, a, lab)
    seg, aux_l, kt = seg + l[0], aux_l + l[1], kt + l[2]
print("losses mine", seg.item(), aux_l.item(), kt.item(), " oracle", Lr['seg'].item(), Lr['aux'].item(), Lr['kt'].item())
m.zero_grad()
for n, p in m.named_parameters():
    if 'domain_layer' in n: p.requires_grad = False
aux_l.backward(retain_graph=True)
for n, p in m.named_parameters():
    if 'domain_layer' in n: p.requires_grad = True
(0.5 * kt + 0.5 * seg).backward()
torch.cuda.synchronize()
gmax = max(g.abs().max().item() for g in gr.values() if g is not None)
errs = []
for n, p in m.named_parameters():
    g = gr[n]
    if p.grad is None or g is None:
        print("NONE", n, p.grad is None, g is None); continue
    e_rel = rel(p.grad, g)
    e_glob = (p.grad - g).abs().max().item() / gmax
    errs.append((e_rel, e_glob, n, g.abs().max().item()))
errs.sort(reverse=True)
print("gmax", gmax)
for e in errs[:40]: print("grad rel %.3e glob %.3e %s |g|max %.3e" % e)
bad = [e for e in errs if e[0] > 0.05 and e[3] > 1e-6 * gmax]
print("n params", len(errs), "n bad (rel>5%):", len(bad))
