"""TransFuse_S_adapt (BASELINE.json config 4 / SURVEY 8 f-1) on the GPU, through the C ABI:
 * op level — dense conv (+BN, +residual, +ReLU), max-pool, align_corners resize, structure_loss against plain torch fp32 of the
   same op (TF32 forward GEMMs: 2e-3 of the abs-max; bf16 gradient GEMMs: 2e-2);
 * model level — the three logit maps, the deep-supervision loss, every parameter-gradient norm and the BatchNorm running
   statistics against goldens of the UNMODIFIED reference at its own random init (oracle/make_golden_transfuse_model.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.make_golden_transfuse_model import case, structure_loss_ref
from tests.helpers import fingerprint
from tests.test_transfuse_wiring import _EmuConv

GOLD = os.path.join(os.path.dirname(__file__), "golden", "transfuse_model_golden.npz")
pytestmark = pytest.mark.gpu
FWD_TOL, BWD_TOL = 2e-3, 2e-2


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda")


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


# (Cin, Cout, k, stride, H, W, bn, act, residual, nchw, bias)
CONV_CASES = [
    (64, 64, 3, 1, 16, 16, True, True, False, False, False),      # BasicBlock conv1
    (64, 64, 3, 1, 16, 16, True, True, True, False, False),       # BasicBlock conv2 + identity
    (64, 128, 3, 2, 16, 16, True, True, False, False, False),     # layer2.0.conv1 (stride 2)
    (64, 128, 1, 2, 16, 16, True, False, False, False, False),    # layer2.0.downsample
    (3, 64, 7, 2, 32, 32, True, True, False, True, False),        # resnet.conv1 on the NCHW image
    (2, 1, 7, 1, 16, 16, False, False, False, False, False),      # BiFusion_block.spatial.conv
    (64, 1, 3, 1, 16, 16, False, False, False, False, True),      # output heads
    (128, 1, 1, 1, 8, 8, False, False, False, False, True),       # Attention_block.psi
    (384, 128, 1, 1, 8, 8, True, False, False, False, True),      # W_x / identity convs
    (128, 128, 1, 1, 8, 8, True, True, True, False, True),        # relu(g1 + x1)
    (64, 128, 1, 1, 8, 12, False, False, True, False, True),      # Residual.conv3 + residual (GEMM epilogue add)
    (192, 64, 3, 1, 12, 8, True, True, False, False, True),       # DoubleConv first conv, ragged map
]


@pytest.mark.parametrize("case_", CONV_CASES)
def test_conv_bn_act_fwd_bwd_vs_torch(case_):
    from mdvit_b200 import ops
    dev = _dev()
    Cin, Cout, k, s, H, W, bn, act, res, nchw, bias = case_
    B = 3
    g = torch.Generator().manual_seed(k * 1000 + Cin + Cout)
    x = torch.randn((B, Cin, H, W) if nchw else (B, H * W, Cin), generator=g).to(dev).requires_grad_(not nchw)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5).to(dev).requires_grad_(True)
    cb = (0.3 * torch.randn(Cout, generator=g)).to(dev).requires_grad_(True) if bias else None
    gam = (1 + 0.2 * torch.randn(Cout, generator=g)).to(dev).requires_grad_(True) if bn else None
    bet = (0.2 * torch.randn(Cout, generator=g)).to(dev).requires_grad_(True) if bn else None
    Ho, Wo, _ = ops.conv_geom(H, W, k, s)
    r = torch.randn((B, Ho * Wo, Cout), generator=g).to(dev).requires_grad_(True) if res else None
    probe = torch.randn((B, Ho * Wo, Cout), generator=g).to(dev)
    a = ops.ACT_RELU if act else ops.ACT_NONE
    outs = []
    for fn in (ops.ConvBnActFn, _EmuConv):
        bufs = (torch.zeros(Cout, device=dev), torch.ones(Cout, device=dev), torch.zeros((), dtype=torch.long, device=dev)) if bn else None
        for t in (x, w, cb, gam, bet, r):
            if t is not None:
                t.grad = None
        y = fn.apply(x, w, cb, gam, bet, r, bufs, B, H, W, s, a, True, nchw)
        (y * probe).sum().backward()
        outs.append((y.detach(), [None if (t is None or t.grad is None) else t.grad.clone() for t in (x, w, cb, gam, bet, r)], bufs))
    (y0, g0, b0), (y1, g1, b1) = outs
    assert rel(y0, y1) < FWD_TOL
    for name, a_, b_ in zip(("dx", "dw", "dbias", "dgamma", "dbeta", "dres"), g0, g1):
        assert (a_ is None) == (b_ is None), name
        if a_ is None:
            continue
        if name == "dbias" and bn:      # exactly zero in exact arithmetic (BatchNorm removes the bias): only round-off to compare
            assert torch.isfinite(a_).all()
            continue
        assert rel(a_, b_) < BWD_TOL, (name, rel(a_, b_))
    if bn:
        assert rel(b0[0], b1[0]) < FWD_TOL and rel(b0[1], b1[1]) < FWD_TOL and int(b0[2]) == 1


def test_bn_act_fn_vs_torch():
    from mdvit_b200 import ops
    dev = _dev()
    B, N, C = 3, 96, 192
    x = torch.randn(B, N, C, device=dev).requires_grad_(True)
    gam, bet = (1 + 0.2 * torch.randn(C, device=dev)).requires_grad_(True), (0.2 * torch.randn(C, device=dev)).requires_grad_(True)
    probe = torch.randn(B, N, C, device=dev)
    res = []
    for mine in (True, False):
        bufs = (torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros((), dtype=torch.long, device=dev))
        for t in (x, gam, bet):
            t.grad = None
        if mine:
            y = ops.BnActFn.apply(x, gam, bet, bufs, ops.ACT_RELU, True)
        else:
            y = torch.relu(F.batch_norm(x.transpose(1, 2), bufs[0], bufs[1], gam, bet, True, 0.1, 1e-5).transpose(1, 2))
        (y * probe).sum().backward()
        res.append((y.detach(), x.grad.clone(), gam.grad.clone(), bet.grad.clone()))
    for a_, b_ in zip(*res):
        assert rel(a_, b_) < 1e-4


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (3, 10, 14, 8), (1, 7, 9, 4)])
def test_maxpool_fwd_bwd_exact(shape):
    from mdvit_b200 import ops
    dev = _dev()
    B, H, W, C = shape
    x = torch.randn(B, H * W, C, device=dev)
    x = torch.relu(x)      # ties at zero, as after resnet's ReLU
    x.requires_grad_(True)
    y = ops.MaxPool3s2Fn.apply(x, H, W)
    xr = x.detach().view(B, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y.view(B, yr.shape[2], yr.shape[3], C).permute(0, 3, 1, 2), yr)
    probe = torch.randn_like(yr)
    (yr * probe).sum().backward()
    (y * probe.permute(0, 2, 3, 1).reshape(B, -1, C)).sum().backward()
    gr = xr.grad.permute(0, 2, 3, 1).reshape(B, H * W, C)
    # ties (equal maxima inside a window) may be routed to a different tap than cuDNN's; away from ties the routing is identical
    tie = (x.detach() == 0)
    assert torch.allclose(x.grad[~tie], gr[~tie], rtol=1e-6, atol=1e-6)
    assert torch.allclose(x.grad.sum(), gr.sum(), rtol=1e-4)


@pytest.mark.parametrize("shape", [(2, 16, 16, 32, 32, 384), (2, 8, 8, 16, 16, 128), (3, 16, 16, 256, 256, 1), (2, 64, 64, 256, 256, 1),
                                   (1, 5, 7, 10, 14, 4)])
def test_resize_align_corners_fwd_bwd(shape):
    from mdvit_b200 import ops
    dev = _dev()
    B, H, W, Ho, Wo, C = shape
    x = torch.randn(B, H * W, C, device=dev).requires_grad_(True)
    y = ops.ResizeACFn.apply(x, H, W, Ho, Wo)
    xr = x.detach().view(B, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.interpolate(xr, size=(Ho, Wo), mode="bilinear", align_corners=True)
    assert rel(y.view(B, Ho, Wo, C).permute(0, 3, 1, 2), yr) < 1e-5
    probe = torch.randn_like(yr)
    (yr * probe).sum().backward()
    (y * probe.permute(0, 2, 3, 1).reshape(B, -1, C)).sum().backward()
    assert rel(x.grad, xr.grad.permute(0, 2, 3, 1).reshape(B, H * W, C)) < 1e-5


def test_structure_loss_vs_reference_formula_and_golden():
    from mdvit_b200 import ops
    dev = _dev()
    g = np.load(GOLD)
    _, mask, _ = case()
    mask = mask.to(dev)
    weit = ops.structure_weit(mask)
    ref_weit = 1 + 5 * torch.abs(F.avg_pool2d(mask, kernel_size=31, stride=1, padding=15) - mask)
    assert rel(weit, ref_weit) < 1e-5
    for i, n in enumerate(("map_x", "map_1", "map_2")):
        pred = torch.from_numpy(g[n]).to(dev).requires_grad_(True)
        loss = ops.structure_loss(pred, mask, weit)
        assert abs(loss.item() - g["losses"][i]) < 2e-5 * abs(g["losses"][i])
        coef = (0.2, 0.3, 0.5)[i]
        (coef * loss).backward()
        assert rel(pred.grad, g["d" + n]) < 1e-4
    # a saturated / random case against the formula itself
    pred = (8 * torch.randn(3, 1, 64, 64, device=dev)).requires_grad_(True)
    m2 = (torch.rand(3, 1, 64, 64, device=dev) > 0.6).float()
    l1 = ops.structure_loss(pred, m2)
    l1.backward()
    g1 = pred.grad.clone()
    pred.grad = None
    l2 = structure_loss_ref(pred, m2)
    l2.backward()
    assert abs(l1.item() - l2.item()) < 1e-5 * abs(l2.item()) and rel(g1, pred.grad) < 1e-4


@pytest.fixture(scope="module")
def trained_once():
    """one training forward + backward of TransFuse_S_adapt at the reference's random init on the golden batch"""
    from mdvit_b200 import ops, transfuse as T
    dev = _dev()
    torch.manual_seed(0)
    m = T.TransFuse_S_adapt(drop_rate=0.0).to(dev).train()
    img, mask, dlab = case()
    img, mask, dlab = img.to(dev), mask.to(dev), dlab.to(dev)
    maps = m(img, dlab)
    weit = ops.structure_weit(mask)
    losses = [ops.structure_loss(p, mask, weit) for p in maps]
    loss = 0.5 * losses[2] + 0.3 * losses[1] + 0.2 * losses[0]
    loss.backward()
    torch.cuda.synchronize()
    return m, maps, losses, loss, (img, mask, dlab)


def test_model_maps_and_losses_match_reference_golden(trained_once):
    g = np.load(GOLD)
    m, maps, losses, loss, _ = trained_once
    for n, p in zip(("map_x", "map_1", "map_2"), maps):
        assert tuple(p.shape) == g[n].shape
        assert rel(p.detach(), g[n]) < 1e-2, (n, rel(p.detach(), g[n]))
    got = np.asarray([l.item() for l in losses] + [loss.item()])
    np.testing.assert_allclose(got, g["losses"], rtol=5e-3)
    for k in g.files:
        if k.startswith("buf."):
            assert rel(m.state_dict()[k[4:]], g[k]) < 1e-2, k


def test_model_gradients_match_reference_golden(trained_once):
    g = np.load(GOLD)
    m = trained_once[0]
    named = [(n, p.grad) for n, p in m.named_parameters() if p.grad is not None]
    assert [n for n, _ in named] == list(g["grad_names"])
    assert all(torch.isfinite(t).all().item() for _, t in named)
    fp, ref_fp = fingerprint(named), g["grad_fp"]
    floor = 1e-3 * np.median(ref_fp[:, 0])      # (conv biases in front of a BatchNorm: true gradient exactly zero)
    err = np.abs(fp[:, 0] - ref_fp[:, 0]) / (ref_fp[:, 0] + floor)
    assert err.max() < 0.1, (named[int(err.argmax())][0], err.max())
    assert np.median(err) < 2e-2
    for k in g.files:
        if k.startswith("grad."):
            got = dict(named)[k[5:]]
            assert rel(got, g[k]) < 0.1 or np.abs(g[k]).max() < floor, (k, rel(got, g[k]))


def test_model_eval_maps_match_reference_golden(trained_once):
    g = np.load(GOLD)
    m, _, _, _, (img, mask, dlab) = trained_once
    m.eval()
    with torch.no_grad():
        emaps = m(img, dlab)
    m.train()
    for n, p in zip(("eval_map_x", "eval_map_1", "eval_map_2"), emaps):
        assert rel(p, g[n].astype(np.float32)) < 2e-2, (n, rel(p, g[n].astype(np.float32)))
