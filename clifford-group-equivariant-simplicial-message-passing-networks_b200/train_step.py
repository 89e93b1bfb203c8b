"""Data-parallel train step for the shared simplicial models.

The reference wraps its models in ``DistributedDataParallel`` (csmpn/md17.py:15-20, csmpn/nba.py:15-20) and shards
samples with ``DistributedSampler`` (csmpn/data/md17.py:143-150); the step order is the trainer's
(engineer/trainer/trainer.py:204-216): forward -> zero_grad -> backward -> optimizer step.  Complexes are independent,
so the only exchange is ONE gradient all-reduce (mean) per step.  All parameters of these models together are
0.8-1.5 MB, i.e. a latency-bound collective: the gradients live in ONE flat fp32 bucket (every ``p.grad`` is a view
into it after one multi-tensor pack), and a single NCCL all-reduce over NVLink covers the model.  On one GPU the optimizer
reads the gradient tensors of the backward pass directly.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_samples: int, rank: int, world: int, drop_last: bool = False):
    """Sample ids of ``rank`` -- DistributedSampler semantics without shuffling: ids rank, rank + world, ...; when
    ``n_samples`` is not a multiple of ``world`` the list is padded by wrapping around (or truncated with drop_last)
    so every rank gets the same count (equal counts make the mean of per-rank mean losses the global mean)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    ids = list(range(n_samples))
    if drop_last:
        total = (n_samples // world) * world
        ids = ids[:total]
    else:
        total = -(-n_samples // world) * world
        if n_samples:
            while len(ids) < total:
                ids += ids[: total - len(ids)]
    return ids[rank:total:world]


class FlatGradBucket:
    """One contiguous fp32 buffer holding every parameter gradient; ``p.grad`` are views into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("FlatGradBucket needs all parameters on one device with one dtype")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()

    def attach(self):
        """p.grad = its view into the flat buffer (a backward pass then ACCUMULATES into the bucket)"""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def release(self):
        """p.grad = None: the next backward pass hands its gradient tensors over without the per-parameter
        `grad += new` kernel (one tiny launch per parameter tensor, ~200 per step for these models)"""
        for p in self.params:
            p.grad = None

    def gather(self, grads=None):
        """pack the gradients produced by a backward pass after release() into the flat buffer (multi-tensor copy),
        then attach the views; parameters that received no gradient get zeros"""
        grads = [p.grad for p in self.params] if grads is None else grads
        have = [(v, g) for v, g in zip(self.views, grads) if g is not None]
        if len(have) != len(self.views):
            self.flat.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        self.attach()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """sum over ranks, then divide by the world size (what DDP does)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class DataParallelStep:
    """forward -> zero -> backward -> all-reduce(mean) -> optimizer step, on this rank's shard of the batch."""

    def __init__(self, model, optimizer, group=None):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.bucket = FlatGradBucket(model.parameters())

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def __call__(self, batch, step: int = 0):
        loss, out = self.model(batch, step, "train")
        self.bucket.release()
        loss.backward()
        if self._distributed():
            self.bucket.gather()
            self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()
        return loss, out


class GraphedDataParallelStep(DataParallelStep):
    """The same step for a batch whose STRUCTURE is fixed (fixed-topology data such as the motion skeleton, or one
    batch trained on repeatedly), replayed from ONE CUDA graph.  A train step of these models is ~600 launches of
    3-200 us, so eager launching leaves the GPU idle 10-15 % of the step.

    What the graph holds: forward + backward always; on one GPU also the optimizer step (``capturable`` fused Adam
    reading the learning rate from a device tensor, so an LR schedule keeps working); with several ranks the pack of the
    gradients into the flat bucket, the NCCL all-reduce (mean) and the optimizer step are captured too when
    ``capture_collective`` holds and the capture succeeds, else they follow the replay eagerly (3-4 launches).

    The captured graph reads the batch tensors that were bound to ``batch`` at construction.  They are snapshotted and
    re-bound before every model call, so models that rebind an input attribute in ``forward`` (motion: ``graph.pos``,
    hulls: ``batch.input``, md17: ``graph.pos``) replay correctly; new input VALUES are fed with ``update`` (a copy
    into those tensors).  The CSR of the batch and its sorted views are held here as well (the identity-keyed cache
    behind ``get_csr`` may evict them, the graph holds raw pointers into them)."""

    def __init__(self, model, optimizer, batch, group=None, warmup: int = 3, capture_optimizer=None, capture_collective=True):
        super().__init__(model, optimizer, group)
        from .models.ops import get_csr

        self.batch = batch
        self._inputs = {k: v for k, v in vars(batch).items() if torch.is_tensor(v)} if hasattr(batch, "__dict__") else {}
        n_nodes = int(batch.x_ind.shape[0]) if hasattr(batch, "x_ind") else int(batch.node_types.shape[0])
        self._csr = get_csr(batch.edge_index, n_nodes)  # strong reference for the lifetime of the graph
        dist_on = self._distributed()
        capturable = all(bool(g.get("capturable", False)) for g in optimizer.param_groups)
        if capture_optimizer is None:
            capture_optimizer = capturable and (not dist_on or capture_collective)
        if capture_optimizer and not capturable:
            raise ValueError("capture_optimizer needs an optimizer built with capturable=True")
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):  # builds every per-batch cache (CSR, simplex rows) and the autograd buffers
                self._rebind()
                loss, _ = self.model(batch, 0, "train")
                self.bucket.release()
                loss.backward()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._csr_sorted = getattr(self._csr, "_sorted", None)
        self.in_graph = "forward + backward"
        try:
            self._capture(capture_optimizer, dist_on)
        except Exception:
            if not (capture_optimizer and dist_on):
                raise
            torch.cuda.synchronize()
            capture_optimizer = False  # e.g. an NCCL build that refuses stream capture: collective and Adam stay eager
            self._capture(False, dist_on)
        self.captured_optimizer = capture_optimizer
        self._rebind()

    def _rebind(self):
        for k, v in self._inputs.items():
            setattr(self.batch, k, v)

    def _init_optimizer_state(self):
        """Adam's lazily created state must exist BEFORE capture (created inside it, its zero-fills would be replayed on
        every step); same tensors torch.optim.Adam._init_group makes for a capturable optimizer, no parameter touched."""
        opt = self.optimizer
        if not isinstance(opt, (torch.optim.Adam, torch.optim.AdamW)):
            if any(len(opt.state[p]) == 0 for g in opt.param_groups for p in g["params"] if p.requires_grad):
                raise ValueError("capture_optimizer: initialise the optimizer state (one step) before capturing a non-Adam optimizer")
            return
        for g in opt.param_groups:
            if g.get("amsgrad"):
                raise ValueError("capture_optimizer: amsgrad is not supported")
            for p in g["params"]:
                st = opt.state[p]
                if p.requires_grad and len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _capture(self, with_optimizer, dist_on):
        if with_optimizer:
            self._init_optimizer_state()
        self.graph = torch.cuda.CUDAGraph()
        self.bucket.release()
        self._rebind()
        with torch.cuda.graph(self.graph):
            loss, out = self.model(self.batch, 0, "train")
            loss.backward()
            # the gradient tensors the captured backward writes on every replay (static addresses inside the graph's pool)
            self.static_grads = [p.grad for p in self.bucket.params]
            if with_optimizer:
                if dist_on:
                    self.bucket.gather(self.static_grads)
                    self.bucket.all_reduce_mean(self.group)
                self.optimizer.step()
        self.loss, self.out = loss.detach(), {k: v.detach() for k, v in out.items()}
        self.in_graph = "forward + backward" + (" + gradient pack + NCCL all-reduce" if with_optimizer and dist_on else "") + \
                        (" + fused Adam (capturable)" if with_optimizer else "")

    def describe(self):
        rest = "" if self.captured_optimizer else ("; gradient pack + all-reduce + Adam eager" if self._distributed() else "; Adam eager")
        return f"one CUDA graph per step: {self.in_graph}{rest} (GraphedDataParallelStep)"

    def update(self, **tensors):
        for k, v in tensors.items():
            self._inputs[k].copy_(v)

    def __call__(self, batch=None, step: int = 0):
        if batch is not None and batch is not self.batch:
            raise ValueError("GraphedDataParallelStep replays the batch it was captured on; use update() for new values")
        self.graph.replay()
        if not self.captured_optimizer:
            if self._distributed():
                self.bucket.gather(self.static_grads)
                self.bucket.all_reduce_mean(self.group)
            self.optimizer.step()
        return self.loss, self.out


class StreamGraphedStep(GraphedDataParallelStep):
    """The graphed step for a stream of DIFFERENT batches (MD17 / NBA complexes change from sample to sample; SURVEY 8f-4,
    engineer/trainer/trainer.py:204-227 feeds a new batch every step).

    Every batch is padded to one ``Bucket`` of simplex / pair counts with a dummy complex (data/padding.py) that neither
    talks to the real complexes nor enters the loss, so ONE captured graph -- forward + backward (+ all-reduce + Adam) --
    serves every step.  Per step, eagerly: the lifted batch is written into the graph's static tensors (``pad_to_bucket``
    in place), the CSR of its pairs is rebuilt in place (one small graph replay), then the step graph is replayed.  A batch
    that overflows the bucket raises ``BucketOverflow``: build a new step with a larger bucket (``make_bucket``)."""

    def __init__(self, model, optimizer, first_batch, bucket, **kw):
        from .data.padding import pad_to_bucket

        self.bucket_shape = bucket
        padded = pad_to_bucket(first_batch, bucket)
        super().__init__(model, optimizer, padded, **kw)

    def load(self, batch):
        """write ``batch`` (a collated lift) into the graph's static tensors, padded to the bucket, and rebuild its CSR"""
        from .data.padding import pad_to_bucket

        self._rebind()
        pad_to_bucket(batch, self.bucket_shape, out=self.batch)      # in place: the graph's static tensors keep their addresses
        self._csr.rebuild_(self.batch.edge_index)                     # counting sorts + sorted views, replayed from a small graph

    def run(self, step: int = 0):
        return super().__call__(None, step)

    def __call__(self, batch, step: int = 0):
        self.load(batch)
        return self.run(step)

    def describe(self):
        return super().describe().replace("(GraphedDataParallelStep)", "(StreamGraphedStep: shape-padded batches, CSR rebuilt in place)")


class CosineAnnealingLR(torch.optim.lr_scheduler.LRScheduler):
    """The reference's schedule (engineer/schedulers/cosine.py:10-46; built in csmpn/md17.py:26-36 with
    warmup = steps / 64, decay = steps / 4): half-cosine warm-up over ``warmup_steps``, flat at the base rate for
    ``max_steps - warmup_steps - decay_steps`` steps, half-cosine decay to 0 over ``decay_steps``.  ``step()`` is called
    once per optimizer step, after it (engineer/trainer/trainer.py:347-349).  Works with a device-tensor learning rate
    (capturable fused Adam inside a CUDA graph): torch's scheduler base writes the new value into the tensor."""

    def __init__(self, optimizer, max_steps: int, warmup_steps: int = 0, decay_steps: int = 0, last_step: int = -1):
        self.warmup_steps = warmup_steps
        self.max_steps = max_steps
        self.stable_steps = max_steps - warmup_steps - decay_steps
        self.decay_steps = decay_steps
        super().__init__(optimizer, last_step)

    def scale(self, step: int) -> float:
        import math

        if step < self.warmup_steps:
            return 0.5 - 0.5 * math.cos(math.pi * step / self.warmup_steps)
        if step < self.warmup_steps + self.stable_steps:
            return 1.0
        return 0.5 + 0.5 * math.cos(math.pi * (step - self.warmup_steps - self.stable_steps) / self.decay_steps)

    def get_lr(self):
        s = self.scale(self.last_epoch)
        return [base * s for base in self.base_lrs]
