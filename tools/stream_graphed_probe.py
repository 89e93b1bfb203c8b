#!/usr/bin/env python
"""Host-side cost of the padded-stream train step (StreamGraphedStep): wall time of every phase with a device synchronise
after each, and the pipelined step time.  Shows whether the step is bound by the host (lifting / padding launches) or the GPU."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from csmpn_b200.data.modules.simplicial_data import SimplicialTransform
    from csmpn_b200.data.padding import make_bucket, pad_to_bucket
    from csmpn_b200.models.md17_cssmpnn import CliffordSharedSimplicialMPNN_md17
    from csmpn_b200.pipeline import LiftPrefetcher
    from csmpn_b200.train_step import StreamGraphedStep

    dev = torch.device("cuda:0")
    lift = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin")
    pool = [bench.make_md17_graphs(100, 3000 + 17 * k, "cpu") for k in range(8)]
    for gs in pool:
        for g in gs:
            for k in ("loc", "vel", "edge_index", "charges", "y"):
                setattr(g, k, getattr(g, k).pin_memory())
    torch.manual_seed(0)
    model = CliffordSharedSimplicialMPNN_md17().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
    lifted = [lift.lift(gs, device=dev) for gs in pool]
    bucket = make_bucket([b.sizes for b in lifted])
    sstep = StreamGraphedStep(model, opt, lifted[0], bucket)
    sync = lambda: torch.cuda.synchronize()
    acc = {}

    def timed(name, fn):
        sync()
        t0 = time.perf_counter()
        r = fn()
        t1 = time.perf_counter()
        sync()
        t2 = time.perf_counter()
        a = acc.setdefault(name, [0.0, 0.0, 0])
        a[0] += t1 - t0; a[1] += t2 - t0; a[2] += 1
        return r

    for i in range(3):
        sstep(lift.lift(pool[i % 8], device=dev), i)
    for i in range(16):
        b = timed("lift (H2D + GPU lifting + collate)", lambda: lift.lift(pool[i % 8], device=dev))
        timed("pad_to_bucket in place", lambda: (sstep._rebind(), pad_to_bucket(b, sstep.bucket_shape, out=sstep.batch)))
        timed("csr rebuild", lambda: sstep._csr.rebuild_(sstep.batch.edge_index))
        timed("graph replay", lambda: sstep.run(i))
    for k, (h, tot, n) in acc.items():
        print(f"{k:40s} host {h / n * 1e3:7.3f} ms   host+device {tot / n * 1e3:7.3f} ms")
    pre = LiftPrefetcher(lambda s: lift.lift(s, device=dev), dev)
    pre.submit(pool[0])
    sync()
    t0 = time.perf_counter()
    n = 32
    for i in range(n):
        sstep.load(pre.take()); pre.consumed(); sstep.run(i); pre.submit(pool[(i + 1) % 8])
    sync()
    print(f"pipelined step: {(time.perf_counter() - t0) / n * 1e3:.3f} ms")


if __name__ == "__main__":
    main()
