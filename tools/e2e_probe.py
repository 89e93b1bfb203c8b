#!/usr/bin/env python
"""Where does the host->device->host layer step spend its time?  Variants of bench.py's e2e loop."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

def main():
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models.cegnn_utils import EGCL
    from csmpn_b200.models.ops import CSRGraph
    from csmpn_b200.pipeline import HostFeeder
    metric, C, aggr, ncx, desc = bench.WORKLOADS["md17"]
    b = bench.make_batch("md17", ncx, 1000)
    dev = torch.device("cuda:0")
    N, B = b["N"], b["B"]
    alg = CliffordAlgebra(metric).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=6, node_attr_features=3, aggr=aggr).to(dev)
    params = list(layer.parameters())
    cot = b["cot"].to(dev)
    pin = {k: b[k].pin_memory() for k in ("h", "edge_index", "node_attr", "edge_attr")}
    y_host = torch.empty((N, C, B)).pin_memory()
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    K = 20

    def region(fn, name):
        fn(3); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(K); e1.record(); th = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"{name:55s} device {e0.elapsed_time(e1)/K:7.3f} ms/step   host issue {th/K*1e3:7.3f} ms/step", flush=True)

    def compute(dv):
        hh = dv["h"].detach().requires_grad_()
        y = layer(hh, CSRGraph(dv["edge_index"], N), dv["edge_attr"], dv["node_attr"])
        torch.autograd.grad(y, [hh] + params, cot)
        return y.detach()

    def seq(n, do_flush=False):
        for _ in range(n):
            dv = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
            if do_flush: flush.fill_(1.0)
            y = compute(dv)
            y_host.copy_(y, non_blocking=True)
    def only_copies(n):
        for _ in range(n):
            dv = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
            y_host.copy_(dv["h"], non_blocking=True)
    resident = {k: v.to(dev) for k, v in pin.items()}
    def only_compute(n):
        for _ in range(n):
            compute(resident)
    def only_csr(n):
        for _ in range(n):
            CSRGraph(resident["edge_index"], N)
    def feeder(n, do_flush=True, do_drain=True, do_submit=True):
        f = HostFeeder(dev)
        f.submit(pin)
        for i in range(n):
            dv = f.next()
            if i + 1 < n:
                if do_submit: f.submit(pin)
                else: f._submitted += 1; 
            if do_flush: flush.fill_(1.0)
            y = compute(dv)
            if do_drain: f.drain(y, y_host)
            f.release(dv)
        f.join()
    region(only_copies, "copies only (H2D 21 MB + D2H 9 MB, one stream)")
    region(only_csr, "CSR build only")
    region(only_compute, "compute only (CSR + fwd + bwd, resident inputs)")
    region(seq, "sequential: H2D, compute, D2H on one stream")
    region(lambda n: seq(n, True), "sequential + L2 flush")
    region(lambda n: feeder(n, False, False, True), "feeder: submit, no flush, no drain")
    region(lambda n: feeder(n, False, True, True), "feeder: submit + drain, no flush")
    region(lambda n: feeder(n, True, True, True), "feeder: submit + drain + flush")

main()
