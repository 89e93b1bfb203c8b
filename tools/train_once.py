#!/usr/bin/env python
"""One md17-model train step (bench.py train leg) inside a cudaProfilerStart/Stop window for `ncu --profile-from-start off`;
also prints eager wall/GPU time per step.  usage: train_once.py [complexes]"""
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
    from csmpn_b200.train_step import DataParallelStep

    ncx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    dev = torch.device("cuda:0")
    graphs = bench.make_md17_graphs(ncx, 2000, dev)
    batch = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin").lift(graphs, device=dev)
    torch.manual_seed(0)
    model = CliffordSharedSimplicialMPNN_md17().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    step = DataParallelStep(model, opt)
    loc0 = batch.loc.clone()

    def one():
        batch.loc = loc0
        return step(batch)[0]

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        one()
    e1.record()
    t_host = (time.perf_counter() - t0) / 5 * 1e3  # time for the host to ISSUE a step
    torch.cuda.synchronize()
    print(f"train step: host issue {t_host:.2f} ms/step, device {e0.elapsed_time(e1) / 5:.2f} ms/step", flush=True)
    torch.cuda.cudart().cudaProfilerStart()
    one()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
