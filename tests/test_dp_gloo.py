"""CPU, world_size 2 over gloo: the host-side data-parallel logic of mdvit_b200.train_step (no GPU, no kernels).

* GradBucketer: bucket layout covers the flat buffer exactly once; a bucket is all-reduced exactly once, as soon as the
  last of its parameters has been reported by the last-run domain graph (shared parameters need two reports, parameters
  that graph never touches are final from the start); the reduced flat buffer is the SUM over ranks.
* Dice/BCE partial sums: all-reducing the 8 loss sums before the ratio gives the global-batch loss of the gathered batch
  (what nn.DataParallel computes on GPU0, multi_train_MDViT.py:73-74,147-169), checked with the oracle formulas.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdvit_b200.train_step import GradBucketer


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return dict(ret)


def _bucketer_case(rank, world):
    torch.manual_seed(100 + rank)
    sizes = [30, 7, 64, 5, 129, 12, 40, 3, 77, 20]          # 10 "parameters"
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 3) // 4 * 4
    flat = torch.randn(total)
    local = flat.clone()
    ranges = {i: (o, n) for i, (o, n) in enumerate(zip(offs, sizes))}
    bk = GradBucketer(flat, ranges, n_buckets=4, first_bucket_frac=0.05)
    assert bk.bounds[0][1] - bk.bounds[0][0] < total // 4     # the first bucket (finalised last) is the small one
    # layout: contiguous, disjoint, covers everything
    assert bk.bounds[0][0] == 0 and bk.bounds[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(bk.bounds, bk.bounds[1:]))
    # uses: params 2 and 4 are shared by two Functions; params 8, 9 are never touched by the last-run graph
    assert bk.needs_recording("fused")
    bk.start_recording("fused")
    for keys in ([0, 1], [2, 3], [2, 4], [4, 5], [6, 7]):
        bk.record_use(keys)
    assert not bk.needs_recording("fused") and bk.needs_recording("per_domain")
    bk.begin("fused")
    log = []
    order = [[6, 7], [4, 5], [2, 4], [2, 3], [0, 1]]          # backward visits Functions in reverse
    for keys in order:
        before = list(bk.reduced)
        bk.report(keys)
        log.append([b for b in bk.reduced if b not in before])
    bk.finish()
    assert sorted(bk.reduced) == list(range(len(bk.bounds))) and len(set(bk.reduced)) == len(bk.reduced)
    # a bucket must not be reduced before all of its parameters were final
    final_at = {}
    left = {k: v for k, v in bk.uses.items()}
    for step, keys in enumerate(order):
        for k in keys:
            left[k] -= 1
            if left[k] == 0:
                final_at[k] = step
    for step, bs in enumerate(log):
        for b in bs:
            members = [k for k, bb in bk.bucket_of.items() if bb == b and bk.uses.get(k, 0) > 0]
            assert all(final_at[k] <= step for k in members), (b, step)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat, sum(gathered), atol=1e-6)
    # guards: a gradient reported after its bucket was reduced (forward path changed without re-recording), and a
    # Function of another domain graph running after the first-forwarded graph started, must raise — never mis-reduce
    import pytest
    with pytest.raises(RuntimeError):
        bk.report([0])              # every bucket has been reduced already
    bk.begin("fused")
    bk.report([6, 7])
    with pytest.raises(RuntimeError):
        bk.report([0], tag=2)       # another domain's graph after tag 0 started
    bk.begin("fused")
    bk.report([0], tag=3)           # before tag 0 starts: fine, ignored
    return log


def test_grad_bucketer_reduces_each_bucket_once_when_final():
    out = _run(_bucketer_case)
    assert out[0] == out[1]                                   # both ranks issue the collectives in the same order
    assert any(len(x) > 0 for x in out[0][:-1])               # some bucket was reduced before the backward finished


def _loss_case(rank, world):
    from oracle import mdvit_oracle as O
    torch.manual_seed(7)
    out, aux = torch.randn(world, 2, 1, 16, 16) * 2, torch.randn(world, 2, 1, 16, 16) * 2
    y = (torch.rand(world, 2, 1, 16, 16) < 0.3).float()
    o, a, t = out[rank], aux[rank], y[rank]
    p, q = torch.sigmoid(o), torch.sigmoid(a)
    bce = lambda s: -(t * torch.log(s).clamp(min=-100) + (1 - t) * torch.log(1 - s).clamp(min=-100)).sum()   # noqa: E731
    sums = torch.stack([bce(p), bce(q), (p * t).sum(), (p * p).sum(), (t * t).sum(), (q * t).sum(), (q * q).sum(), (q * p).sum()]).double()
    dist.all_reduce(sums)                                      # MKDTrainer._reduce_sums
    n = float(out.numel())
    eps = 1e-5
    l_seg = sums[0] / n + 1 - (2 * sums[2] + eps) / (sums[3] + sums[4] + eps)
    l_aux = sums[1] / n + 1 - (2 * sums[5] + eps) / (sums[6] + sums[4] + eps)
    l_kt = 1 - (2 * sums[7] + eps) / (sums[6] + sums[3] + eps)
    ref = O.seg_losses(out.flatten(0, 1), aux.flatten(0, 1), y.flatten(0, 1))      # the gathered global batch
    return [abs(float(l_seg) - float(ref[0])), abs(float(l_aux) - float(ref[1])), abs(float(l_kt) - float(ref[2]))]


def test_global_batch_losses_from_allreduced_partial_sums():
    out = _run(_loss_case)
    assert max(out[0]) < 1e-5 and max(out[1]) < 1e-5
