#!/usr/bin/env python
"""Where the time of the streamed-batch train step goes (bench.py train leg, `stream`): host wall-clock per phase with a
device synchronize after each (so phases do not overlap): lift+collate, forward, backward, optimizer."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from csmpn_b200.data.modules.simplicial_data import SimplicialTransform
    from csmpn_b200.models.md17_cssmpnn import CliffordSharedSimplicialMPNN_md17

    dev = torch.device("cuda:0")
    knn = len(sys.argv) > 1 and sys.argv[1] == "knn"
    lift = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin", knn_k=3 if knn else None)
    pool = [bench.make_md17_graphs(100, 3000 + k, "cpu") for k in range(4)]
    for gs in pool:
        for g in gs:
            for k in ("loc", "vel", "edge_index", "charges", "y"):
                setattr(g, k, getattr(g, k).pin_memory())
    torch.manual_seed(0)
    model = CliffordSharedSimplicialMPNN_md17().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    acc = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return time.perf_counter()

    for i in range(12):
        t = time.perf_counter()
        b = lift.lift(pool[i % 4], device=dev)
        t = tick("lift+collate", t)
        loss, _ = model(b, i, "train")
        t = tick("forward", t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        t = tick("backward", t)
        opt.step()
        t = tick("optimizer", t)
    for k, v in acc.items():
        v = v[4:]
        print(f"{k:14s} {sum(v) / len(v):7.2f} ms")
    print("total", sum(sum(v[4:]) / len(v[4:]) for v in acc.values()))


if __name__ == "__main__":
    main()
