"""Host-side feeding of the layer / model step: double-buffered, stream-ordered copies.

The reference feeds its models from a PyG ``DataLoader`` (csmpn/data/md17.py:143-161) whose batches are moved to the
device inside the trainer loop (engineer/trainer/trainer.py:204-216), copy and compute strictly in sequence.  On B200 a
layer step over 100 complexes is ~2 ms of kernels and ~30 MB of PCIe traffic, so the copies are worth a third of the
step: ``HostFeeder`` stages the inputs of step i+1 on a copy stream while step i computes, and drains results on a
second stream, so a stream of batches costs max(copy, compute) per step instead of their sum.  Buffers are recycled
with events (a slot is overwritten only after the step that read it has finished).
"""
from __future__ import annotations

import torch


class HostFeeder:
    """Double-buffered pinned-host -> device feeder and device -> pinned-host drain.

    >>> feeder = HostFeeder(device)
    >>> feeder.submit(host_batch0)                       # dict of pinned CPU tensors
    >>> for i in range(steps):
    ...     dev = feeder.next()                          # device tensors of step i (compute stream waits for the copy)
    ...     if i + 1 < steps: feeder.submit(host_batch(i + 1))
    ...     y = layer(dev["h"], dev["edge_index"], ...)
    ...     feeder.drain(y, y_host)                      # async D2H on the drain stream
    ...     feeder.release(dev)                          # the slot may be overwritten once this step's kernels are done
    >>> feeder.join()                                    # compute stream waits for all outstanding copies
    """

    def __init__(self, device, slots: int = 2, prepare=None):
        """prepare(slot_index, device_tensors): optional per-batch preprocessing that only depends on the copied inputs
        (e.g. the CSR of the batch's pairs, built into per-slot buffers); it runs on the copy stream right after the copies,
        i.e. under the kernels of the previous step, and the slot only becomes ready once it has finished."""
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.drain_stream = torch.cuda.Stream(self.device)
        self.prepare = prepare
        self.slots = [dict(bufs=None, ready=None, free=None, index=i) for i in range(slots)]
        self._submitted = 0
        self._taken = 0
        self._drain_done = None

    def submit(self, host: dict):
        slot = self.slots[self._submitted % len(self.slots)]
        self._submitted += 1
        with torch.cuda.stream(self.copy_stream):
            if slot["free"] is not None:
                self.copy_stream.wait_event(slot["free"])  # the step that last read this slot has finished
            if slot["bufs"] is None or any(slot["bufs"][k].shape != v.shape or slot["bufs"][k].dtype != v.dtype
                                           for k, v in host.items()):
                slot["bufs"] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            for k, v in host.items():
                slot["bufs"][k].copy_(v, non_blocking=True)
            if self.prepare is not None:
                self.prepare(slot["index"], slot["bufs"])
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            slot["ready"] = ev

    def next(self) -> dict:
        slot = self.slots[self._taken % len(self.slots)]
        self._taken += 1
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        out = dict(slot["bufs"])
        out["_slot"] = slot
        out["_index"] = slot["index"]
        return out

    def release(self, dev: dict):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        dev["_slot"]["free"] = ev

    def drain(self, result: torch.Tensor, host_out: torch.Tensor):
        """device -> pinned host, asynchronously after the kernels that produced ``result``"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.drain_stream):
            self.drain_stream.wait_event(ev)
            host_out.copy_(result, non_blocking=True)
            result.record_stream(self.drain_stream)
            done = torch.cuda.Event()
            done.record(self.drain_stream)
            self._drain_done = done

    @property
    def last_drain_event(self):
        """event recorded after the most recent ``drain`` copy (None before the first): a producer that overwrites the
        drained tensor in place (e.g. a CUDA graph's static output) waits for it first"""
        return self._drain_done

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.copy_stream)
        cur.wait_stream(self.drain_stream)


class LiftPrefetcher:
    """Lifts (pinned H2D of the raw samples + GPU lifting + collation) batch i+1 on a side stream while step i computes.

    The lifter needs one host sync for its output sizes; on the compute stream that sync would wait for the whole previous
    step.  On a side stream it only waits for the lifting kernels, so the host prepares batch i+1 under the kernels of step i.

    >>> pre = LiftPrefetcher(lambda samples: transform.lift(samples, device=dev), dev)
    >>> pre.submit(samples(0))
    >>> for i in range(steps):
    ...     b = pre.take()                 # compute stream waits for the lifting of batch i
    ...     step.load(b); pre.consumed()   # batch i has been copied into the step's static tensors
    ...     pre.submit(samples(i + 1))     # overlaps step.run()
    ...     step.run()
    """

    def __init__(self, lift_fn, device):
        self.lift_fn = lift_fn
        self.device = torch.device(device)
        self.side = torch.cuda.Stream(self.device)
        self._next = None
        self._consumed = None

    def submit(self, samples):
        if self._consumed is not None:   # the memory of an earlier batch may be recycled: its readers come first
            self.side.wait_event(self._consumed)
        with torch.cuda.stream(self.side):
            b = self.lift_fn(samples)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self._next = (b, ev)

    def take(self):
        b, ev = self._next
        self._next = None
        torch.cuda.current_stream(self.device).wait_event(ev)
        return b

    def consumed(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._consumed = ev
