#!/usr/bin/env python
"""Time one EGCL layer (md17 workload of bench.py) forward-only and forward+backward on both block engines."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models.cegnn_utils import EGCL
    from csmpn_b200.models.ops import CSRGraph

    wl = sys.argv[1] if len(sys.argv) > 1 else "md17"
    metric, C, aggr, ncx, desc = bench.WORKLOADS[wl]
    ncx = int(sys.argv[2]) if len(sys.argv) > 2 else ncx
    b = bench.make_batch(wl, ncx, 1000)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    alg = CliffordAlgebra(metric).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=6, node_attr_features=3, aggr=aggr).to(dev)
    params = list(layer.parameters())
    d = {k: b[k].to(dev) for k in ("h", "edge_index", "node_attr", "edge_attr", "cot")}
    graph = CSRGraph(d["edge_index"], b["N"])
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def fwd():
        with torch.no_grad():
            return layer(d["h"], graph, d["edge_attr"], d["node_attr"])

    def fwdbwd():
        h = d["h"].detach().requires_grad_()
        y = layer(h, graph, d["edge_attr"], d["node_attr"])
        torch.autograd.grad(y, [h] + params, d["cot"])

    for tc in ("0", "1"):
        os.environ["CSMPN_TC"] = tc
        for name, fn in (("fwd", fwd), ("fwd+bwd", fwdbwd)):
            ts = []
            for it in range(8):
                flush.fill_(0.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1))
            print(f"{wl} N={b['N']} E={b['E']} C={C} CSMPN_TC={tc} {name}: {sum(ts)/len(ts):.3f} ms", flush=True)


if __name__ == "__main__":
    main()
