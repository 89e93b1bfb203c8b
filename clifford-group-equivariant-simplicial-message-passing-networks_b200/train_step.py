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
    batch trained on repeatedly): forward + gradient zeroing + backward are captured once into a CUDA graph and
    replayed; the gradient all-reduce and the optimizer step follow eagerly (2-4 launches).  A train step of these
    models is ~600 launches of 3-200 us, so eager launching leaves the GPU idle 10-15 % of the step.

    New input VALUES are fed by copying into the captured batch's tensors (``update``)."""

    def __init__(self, model, optimizer, batch, group=None, warmup: int = 3):
        super().__init__(model, optimizer, group)
        self.batch = batch
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):  # builds every per-batch cache (CSR, simplex rows) and the autograd buffers
                loss, _ = self.model(batch, 0, "train")
                self.bucket.release()
                loss.backward()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.bucket.release()
        with torch.cuda.graph(self.graph):
            loss, out = self.model(batch, 0, "train")
            loss.backward()
        self.loss, self.out = loss.detach(), {k: v.detach() for k, v in out.items()}
        # the gradient tensors the captured backward writes on every replay (static addresses inside the graph's pool)
        self.static_grads = [p.grad for p in self.bucket.params]
        # the graph holds raw pointers: keep every tensor of the captured batch alive even if the caller rebinds attributes
        self._keepalive = [v for v in vars(batch).values() if torch.is_tensor(v)] if hasattr(batch, "__dict__") else []

    def update(self, **tensors):
        for k, v in tensors.items():
            getattr(self.batch, k).copy_(v)

    def __call__(self, batch=None, step: int = 0):
        if batch is not None and batch is not self.batch:
            raise ValueError("GraphedDataParallelStep replays the batch it was captured on; use update() for new values")
        self.graph.replay()
        if self._distributed():
            self.bucket.gather(self.static_grads)
            self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()
        return self.loss, self.out
