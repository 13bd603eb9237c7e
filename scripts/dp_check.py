"""Multi-GPU check (run under torchrun, one rank per GPU): the overlapped bucketed all-reduce gives the same gradients as
a single all-reduce after the backward, parameters stay identical across ranks after the fused AdamW step, and the
CUDA-graph-captured step runs with the NCCL collectives inside the graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mdvit_b200 import ops, synth
from mdvit_b200.model import MDViT
from mdvit_b200.train_step import MKDTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
IMG, B = int(os.environ.get("IMG", 64)), int(os.environ.get("B", 2))

def build():
    m = MDViT(img_size=IMG, adapt_method="Sup", num_domains=4, decoder_name="MLPFM").to(dev).train()
    m.load_state_dict(synth.synth_state_dict(0), strict=True)
    for k in range(1, 5):
        getattr(m, f"debranch{k}").dropout.p = 0.0
    return m

def stage(msg):
    print(f"rank {rank}: {msg}", flush=True)

stage("init done")
batches = [tuple(t.to(dev) for t in synth.synth_batch(10 + rank, d, B, IMG, IMG)) + (d,) for d in range(4)]
# (a) overlapped
tr = MKDTrainer(build())
tr.grad.zero_()
losses = tr.forward_losses(batches)
tr.backward(losses)
torch.cuda.synchronize()
g_overlap = tr.grad.clone()
stage("overlapped backward done")
order = list(tr.bucketer.reduced)
# (b) reference: same local backward, one all-reduce at the end
tr2 = MKDTrainer(build())
bk, tr2.bucketer = tr2.bucketer, None
tr2.grad.zero_()
l2 = tr2.forward_losses(batches)
tr2.backward(l2)
dist.all_reduce(tr2.grad)
torch.cuda.synchronize()
stage("reference backward done")
err = ((g_overlap - tr2.grad).norm() / tr2.grad.norm()).item()
# losses are global-batch losses: identical on every rank
lg = [torch.empty_like(losses) for _ in range(world)]
dist.all_gather(lg, losses.detach())
same_loss = all(torch.equal(lg[0], x) for x in lg)
# (c) optimizer step keeps replicas in sync
tr.optimizer_step(); torch.cuda.synchronize()
chk = tr.flat.double().sum().reshape(1)
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
in_sync = all(abs((a - allc[0]).item()) < 1e-6 * abs(allc[0].item()) + 1e-9 for a in allc)
stage("optimizer sync check done")
# (d) graph-captured step with NCCL inside
tr3 = MKDTrainer(build())
tr3.capture(batches, warmup=1)
stage("capture done")
out = tr3.step_graph(None)
torch.cuda.synchronize()
finite = bool(torch.isfinite(out).all().item())
ok = err < 2e-2 and same_loss and in_sync and finite and len(order) == len(tr.bucketer.bounds)
print(f"rank {rank}: overlap-vs-single rel err {err:.2e}; buckets reduced in order {order}; same_loss {same_loss}; in_sync {in_sync}; graph finite {finite}; {'DP_OK' if ok else 'DP_FAIL'}", flush=True)
# a captured graph holds NCCL work: tearing the process group down under it can block, so synchronise and leave hard
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0 if ok else 1)
