"""Developer aid: which bf16 rounding points produce the logit error at the reference's random init?
The fp32 oracle is run on the GPU with bf16 rounding emulated at selected operand sites (fp32 accumulate, as tcgen05 does),
one site family at a time, and compared with the pure fp32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from mdvit_b200 import synth
from oracle import mdvit_oracle as O
from tests.test_randinit_gpu import build_randinit, oracle_sd_from_model, onehot

dev = torch.device("cuda")
SITE = [None]
ON = set()
bf = lambda t: t.bfloat16().float()


class FProxy:
    def __getattr__(self, n):
        return getattr(TF, n)

    def linear(self, x, w, b=None):
        s = SITE[0]
        cnt = COUNT.setdefault(s, [0])
        cnt[0] += 1
        key = f"{s}.{(cnt[0] - 1) % 2}" if s in ("attn", "mlp") else s
        if key in ON or s in ON:
            x, w = bf(x), bf(w)
        y = TF.linear(x, w, b)
        if key == "attn.0" and "qkv_out" in ON:
            y = bf(y)
        return y

    def conv2d(self, x, w, b=None, **kw):
        if SITE[0] in ON and kw.get("groups", 1) == 1:
            x, w = bf(x), bf(w)
        return TF.conv2d(x, w, b, **kw)


COUNT = {}
O.F = FProxy()


def wrap(name, site):
    f = getattr(O, name)

    def g(*a, **k):
        prev, SITE[0] = SITE[0], site
        try:
            return f(*a, **k)
        finally:
            SITE[0] = prev
    setattr(O, name, g)


wrap("factor_attention", "attn"); wrap("mlp", "mlp"); wrap("patch_embed", "pe"); wrap("conv_bn_act", "stem")
wrap("mlp_decoder_fm", "aux")
_db = O.decoder_block
def decoder_block(sd, prefix, x, skip, domain_label, training, **kw):
    SITE[0] = "dec"
    return _db(sd, prefix, x, skip, domain_label, training, **kw)
O.decoder_block = decoder_block   # (its inner mhsa_stage re-sets the site per attention / mlp call)

m = build_randinit(dev)
sd = oracle_sd_from_model(m, False)
img, _ = synth.synth_batch(4321, 0, 4, 256, 256)
img = img.to(dev)


def run(on):
    ON.clear(); ON.update(on); COUNT.clear(); SITE[0] = None
    s = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        return O.mdvit_forward(s, img, onehot(0, 4, dev), "0", training=True)


ro, ra = run(())
for on in ((), ("attn.0",), ("qkv_out",), ("attn.1",), ("attn.0", "qkv_out", "attn.1"), ("mlp.0",), ("mlp.1",), ("mlp",), ("pe",), ("stem",), ("dec",), ("aux",),
           ("attn.0", "qkv_out", "attn.1", "mlp", "pe", "stem", "dec", "aux")):
    o, a = run(on)
    eo = ((o - ro).abs().max() / ro.abs().max()).item(); ea = ((a - ra).abs().max() / ra.abs().max()).item()
    lo = ((o - ro).norm() / ro.norm()).item(); la = ((a - ra).norm() / ra.norm()).item()
    print(f"{'+'.join(on) or 'none':50s} out max {eo:.2e} l2 {lo:.2e} | aux max {ea:.2e} l2 {la:.2e}", flush=True)
with torch.no_grad():
    o, a = m(img, onehot(0, 4, dev), "0")
print(f"{'mdvit_b200 kernels':50s} out max {((o - ro).abs().max() / ro.abs().max()).item():.2e} l2 {((o - ro).norm() / ro.norm()).item():.2e} | "
      f"aux max {((a - ra).abs().max() / ra.abs().max()).item():.2e} l2 {((a - ra).norm() / ra.norm()).item():.2e}")
