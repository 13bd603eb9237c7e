"""The MDViT training step (multi_train_MDViT.py:121-213) on the B200 kernels, single- or multi-GPU.

One step = 4 domain forwards (each a single-domain mini-batch) -> fused BCE+Dice / MKD losses -> the MKD backward ->
gradient all-reduce -> fused AdamW.

MKD backward schedules (both produce the reference's gradients; tests compare them):
  "reference"     multi_train_MDViT.py:195-207 verbatim: aux.backward(retain_graph) with `domain_layer` frozen, then
                  (alpha*kt + (1-alpha)*seg).backward() — every weight gradient kernel runs twice.
  "single_sweep"  gradients are linear in the loss cotangent, so  grad(non-DA) = d(aux + a*kt + (1-a)*seg)  and
                  grad(DA) = d(aux + a*kt + (1-a)*seg) - d(aux).  Pass 1 back-propagates -aux in ops.backward_mode("da_only")
                  (activation-gradient chain + DA gradients only, no other wgrad kernel), pass 2 back-propagates the total
                  loss once with all weight gradients.  Same result, ~1 trunk wgrad sweep cheaper.

Data parallelism (replaces nn.DataParallel, multi_train_MDViT.py:73-74): one process per GPU; every rank runs all four
domain forwards on its own slice of each domain batch.  The 8 loss partial sums are all-reduced so BCE/Dice are those
of the gathered global batch (what DataParallel computes on GPU0); gradients are then a plain SUM over ranks, all-reduced
in one flat fp32 buffer that the fused AdamW consumes in place.  BatchNorm statistics stay per-replica, as in DataParallel.
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib as L
from . import ops
from ._lib import check, ptr


def _align(n, a=4):
    return (n + a - 1) // a * a


class GradBucketer:
    """Overlaps the data-parallel gradient all-reduce with the backward pass (replaces nn.DataParallel's gather/reduce,
    multi_train_MDViT.py:73-74; the stock DDP reducer cannot be used because the MKD step back-propagates one graph twice).

    The flat fp32 gradient buffer is cut into `n_buckets` contiguous ranges.  Every parameter receives contributions
    from all domain graphs; autograd runs the graph of the domain that was forwarded FIRST last, so a parameter is final
    once the Functions of that domain (tag `last_tag`) have reported it as many times as they use it (`uses`, recorded
    during one forward: 2 for the CPE/CRPE shared by a stage's two blocks, 0 for the other domains' aux decoders).  When
    the last parameter of a bucket is final the bucket is all-reduced on a side stream behind an event, so NCCL runs
    under the remaining backward kernels.  Device-agnostic (CPU tensors + gloo in the tests)."""

    def __init__(self, flat_grad, param_ranges, n_buckets=8, group=None, comm_stream=None):
        self.flat, self.group, self.comm = flat_grad, group, comm_stream
        total = flat_grad.numel()
        n_buckets = max(1, min(n_buckets, len(param_ranges)))
        target = (total + n_buckets - 1) // n_buckets
        self.bucket_of, self.bounds = {}, []
        lo, cur = 0, 0
        for i, (key, (o, n)) in enumerate(param_ranges.items()):
            self.bucket_of[key] = len(self.bounds)
            cur = o + n
            if cur - lo >= target or i == len(param_ranges) - 1:
                hi = total if i == len(param_ranges) - 1 else _align(cur)
                self.bounds.append((lo, hi))
                lo = hi
        self.uses = None            # key -> number of reports expected from the last-run domain graph
        self.reduced = []           # bucket ids in the order they were reduced (inspected by the tests)

    # -- one-time recording of how often each parameter is used by the first-forwarded domain
    def record_use(self, keys):
        if self.uses is None:
            self.uses = {}
        for k in keys:
            if k in self.bucket_of:
                self.uses[k] = self.uses.get(k, 0) + 1

    def begin(self):
        assert self.uses is not None, "record_use() must run during one forward first"
        self._left = {k: self.uses.get(k, 0) for k in self.bucket_of}
        self._bucket_left = [0] * len(self.bounds)
        for k, b in self.bucket_of.items():
            if self._left[k] > 0:
                self._bucket_left[b] += 1
        self._started = False
        self.reduced = []

    def _start(self):
        self._started = True
        for b, left in enumerate(self._bucket_left):      # buckets owned entirely by other domains' graphs: already final
            if left == 0:
                self._reduce(b)

    def report(self, keys):
        """Parameters `keys` have received their gradient from one Function of the last-run domain graph."""
        if not self._started:
            self._start()
        for k in keys:
            left = self._left.get(k, 0)
            if left <= 0:
                continue
            self._left[k] = left - 1
            if left == 1:
                b = self.bucket_of[k]
                self._bucket_left[b] -= 1
                if self._bucket_left[b] == 0:
                    self._reduce(b)

    def _reduce(self, b):
        if b in self.reduced:
            return
        self.reduced.append(b)
        lo, hi = self.bounds[b]
        view = self.flat[lo:hi]
        if self.comm is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """Reduce whatever has not been reported final (defensive) and join the side stream."""
        if not self._started:
            self._start()
        for b in range(len(self.bounds)):
            self._reduce(b)
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)


class MKDTrainer:
    def __init__(self, model, lr=1e-4, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, alpha=0.5, process_group=None,
                 num_domains=4, with_aux=True, schedule="single_sweep", n_buckets=8, fuse_domains=True):
        if schedule not in ("single_sweep", "reference"):
            raise ValueError("schedule must be 'single_sweep' or 'reference'")
        self.schedule = schedule
        self.fuse_domains = fuse_domains      # stack the domain mini-batches into one trunk pass (MDViT.forward_multi)
        self.model = model
        self.alpha = alpha
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.num_domains = num_domains
        self.with_aux = with_aux
        self.t = 0
        params, seen = [], set()
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.params = params
        dev = params[0].device
        self.device = dev
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        self.total = total
        with torch.no_grad():
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
            for p, o in zip(params, offs):
                n = p.numel()
                self.flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.grad[o:o + n].view(p.shape)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self.da_params = [p for n, p in model.named_parameters() if "domain_layer" in n]
        ops.bump_weight_epoch()
        self._graph = None
        self.bucketer = None
        if self.world > 1:
            ranges = {id(p): (o, p.numel()) for p, o in zip(params, offs)}
            self.comm_stream = torch.cuda.Stream(device=dev)
            self.bucketer = GradBucketer(self.grad, ranges, n_buckets=n_buckets, group=process_group, comm_stream=self.comm_stream)

    # ------------------------------------------------------------------ pieces
    def _reduce_sums(self, sums):
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.pg)

    def forward_losses(self, batches):
        """batches: list of (img [B,3,H,W], label [B,1,H,W], domain index).  Returns [n_dom, 3] losses (seg, aux, kt)."""
        if (self.fuse_domains and self.with_aux and len(batches) > 1 and hasattr(self.model, "forward_multi")
                and len({tuple(b[0].shape) for b in batches}) == 1):
            return self._forward_losses_fused(batches)
        out = []
        recording = self.bucketer is not None and self.bucketer.uses is None
        for i, (img, label, d) in enumerate(batches):
            ops.set_forward_tag(i)
            if recording and i == 0:
                ops.set_forward_use_cb(lambda tag, params: self.bucketer.record_use([id(p) for p in params if p is not None]))
            elif recording:
                ops.set_forward_use_cb(None)
            B = img.shape[0]
            dl = torch.zeros((B, self.num_domains), dtype=torch.float32, device=img.device)
            dl[:, int(d)] = 1.0
            if self.with_aux:
                o, a = self.model(img, dl, str(d))
            else:
                o, a = self.model(img), None
            n_total = o.numel() * self.world
            out.append(ops.seg_losses(o, a, label, n_total=n_total, reduce_sums=self._reduce_sums if self.world > 1 else None))
        ops.set_forward_use_cb(None)
        ops.set_forward_tag(None)
        return torch.stack(out)

    def _forward_losses_fused(self, batches):
        """All domain mini-batches in one trunk pass (same result as the per-domain loop up to dropout masks)."""
        B = batches[0][0].shape[0]
        G = len(batches)
        dev = batches[0][0].device
        x = torch.cat([b[0] for b in batches], dim=0)
        dl = torch.zeros((G * B, self.num_domains), dtype=torch.float32, device=dev)
        for g, (_, _, d) in enumerate(batches):
            dl[g * B:(g + 1) * B, int(d)] = 1.0
        recording = self.bucketer is not None and self.bucketer.uses is None
        ops.set_forward_tag(0)
        if recording:
            ops.set_forward_use_cb(lambda tag, params: self.bucketer.record_use([id(p) for p in params if p is not None]))
        res = self.model.forward_multi(x, dl, [str(b[2]) for b in batches])
        ops.set_forward_use_cb(None)
        ops.set_forward_tag(None)
        out = []
        for (o, a), (_, label, _) in zip(res, batches):
            n_total = o.numel() * self.world
            out.append(ops.seg_losses(o, a, label, n_total=n_total, reduce_sums=self._reduce_sums if self.world > 1 else None))
        return torch.stack(out)

    def _final_backward(self, loss):
        """The last backward call of the step: gradients become final bucket by bucket and are all-reduced as they do."""
        if self.bucketer is None:
            loss.backward()
            return
        bk = self.bucketer
        bk.begin()
        ops.set_grad_ready_cb(lambda tag, params: bk.report([id(p) for p in params if p is not None]) if tag == 0 else None)
        try:
            loss.backward()
        finally:
            ops.set_grad_ready_cb(None)
        bk.finish()

    def backward(self, losses):
        """multi_train_MDViT.py:195-207: aux pass with DA frozen (retain_graph), then alpha*kt + (1-alpha)*seg."""
        seg, aux, kt = losses[:, 0].sum(), losses[:, 1].sum(), losses[:, 2].sum()
        if self.with_aux and self.schedule == "reference":
            for p in self.da_params:
                p.requires_grad = False
            aux.backward(retain_graph=True)
            for p in self.da_params:
                p.requires_grad = True
            self._final_backward(self.alpha * kt + (1.0 - self.alpha) * seg)
        elif self.with_aux:
            main = self.alpha * kt + (1.0 - self.alpha) * seg
            if self.da_params:
                with ops.backward_mode("da_only"):
                    (-aux).backward(retain_graph=True)
            self._final_backward(aux + main)
        else:
            self._final_backward(seg)

    def _set_hyper(self):
        self.t += 1
        b1, b2 = self.betas
        h = self._hyper_host
        h[0], h[1], h[2], h[3], h[4] = self.lr, b1, b2, self.eps, self.wd
        h[5], h[6], h[7] = 1.0 - b1 ** self.t, 1.0 - b2 ** self.t, 1.0
        self.hyper.copy_(h, non_blocking=True)

    def optimizer_step(self):
        # (gradients were all-reduced bucket by bucket during the last backward call, see GradBucketer)
        with torch.cuda.device(self.device):
            check(L.lib().mdv_adamw(ptr(self.flat), ptr(self.grad), ptr(self.m), ptr(self.v), ptr(self.hyper), self.total, L.stream()),
                  "mdv_adamw")
            ops.rng_bump(self.device)
            mirror = getattr(self, "mirror", None)
            if mirror is not None and mirror.table is not None:
                mirror.refresh()          # bf16 operand copies of the updated weights, one launch

    def _step_body(self, batches):
        ops.reset_stream_ids()
        self.grad.zero_()
        # autograd accumulates into the flat views (p.grad is never None, so `+=` lands in self.grad)
        losses = self.forward_losses(batches)
        self.backward(losses)
        self.optimizer_step()
        return losses.detach()

    # ------------------------------------------------------------------ eager step
    def step(self, batches):
        self._set_hyper()
        ops.bump_weight_epoch()
        return self._step_body(batches)

    # ------------------------------------------------------------------ CUDA-graph step
    def capture(self, example_batches, warmup=2):
        """Capture the whole step into one CUDA graph over static input buffers (launch-bound otherwise: ~2k kernels).
        The bf16 operand copies of the weights become persistent buffers refreshed by one launch at the end of each
        step (ops.WeightMirror) instead of ~200 conversion launches at its start."""
        self.static = [(img.clone(), lab.clone(), d) for img, lab, d in example_batches]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        self.mirror = ops.WeightMirror()
        ops.set_weight_mirror(self.mirror)
        with torch.cuda.stream(side):
            for i in range(max(warmup, 1)):
                self._set_hyper()
                ops.bump_weight_epoch()
                self._step_body(self.static)
            self.mirror.freeze(self.device)
            self.mirror.refresh()                     # the warm-up's last AdamW changed the weights
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        self._set_hyper()
        n0 = L.lib().mdv_launch_count()
        with torch.cuda.graph(self._graph):
            self._static_losses = self._step_body(self.static)   # ends with AdamW + mirror.refresh()
        self.launches_per_step = L.lib().mdv_launch_count() - n0   # kernels of this library captured per step
        return self

    def step_graph(self, batches=None):
        """Replay; `batches` (host or device tensors) are copied into the static buffers first."""
        if batches is not None:
            for (s_img, s_lab, _), (img, lab, _) in zip(self.static, batches):
                s_img.copy_(img, non_blocking=True)
                s_lab.copy_(lab, non_blocking=True)
        self._set_hyper()
        self._graph.replay()
        return self._static_losses
