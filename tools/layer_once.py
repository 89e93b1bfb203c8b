#!/usr/bin/env python
"""One EGCL layer forward+backward (bench.py workload) inside a cudaProfilerStart/Stop window, for
`ncu --profile-from-start off`.  usage: layer_once.py [workload] [complexes] [fwd|fwdbwd] [hidden]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models.cegnn_utils import EGCL, PairedNodeAttr
    from csmpn_b200.models.ops import CSRGraph

    wl = sys.argv[1] if len(sys.argv) > 1 else "md17"
    metric, C, aggr, ncx, desc = bench.WORKLOADS[wl]
    ncx = int(sys.argv[2]) if len(sys.argv) > 2 else ncx
    mode = sys.argv[3] if len(sys.argv) > 3 else "fwdbwd"
    C = int(sys.argv[4]) if len(sys.argv) > 4 else C
    b = bench.make_batch(wl, ncx, 1000, hidden=C)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    alg = CliffordAlgebra(metric).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=6, node_attr_features=3, aggr=aggr).to(dev)
    params = list(layer.parameters())
    d = {k: b[k].to(dev) for k in ("h", "edge_index", "node_attr", "cot")}
    d["edge_attr"] = PairedNodeAttr(d["node_attr"])  # what bench.py and the models pass
    graph = CSRGraph(d["edge_index"], b["N"])

    def step():
        if mode == "fwd":
            with torch.no_grad():
                layer(d["h"], graph, d["edge_attr"], d["node_attr"])
        else:
            h = d["h"].detach().requires_grad_()
            y = layer(h, graph, d["edge_attr"], d["node_attr"])
            torch.autograd.grad(y, [h] + params, d["cot"])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"{wl} N={b['N']} E={b['E']} C={C} {mode} done")


if __name__ == "__main__":
    main()
