#!/usr/bin/env python
"""[needs a diagnostics build: CSMPN_DEBUG_BUILD=1 python -c "import __graft_entry__ as g; g.build()"]
Per-phase timeline of the second forward kernel (tc_f2) of the tensor-core engine: clock64 stamps of two threads of
CTA 0 (csmpn_tc_debug_buffer).  Prints cycles spent between consecutive stamps, aggregated by (from, to) phase code."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CSMPN_TC"] = "1"
os.environ["CSMPN_TC_MIN_ROWS"] = "0"
import bench  # noqa: E402

NAMES = {30: "conv top", 31: "slot free", 32: "chunk stored", 33: "conv_done", 34: "next gather issued", 35: "epilogue step done", 36: "issuer top", 37: "chunk full(f1)", 38: "mma issued(f1)", 2: "kernel entry", 3: "first loads issued", 4: "weights staged", 5: "kernel exit", 1: "start", 10: "chunk top", 11: "load landed", 12: "lo free", 13: "split done", 14: "chunk full", 15: "mma issued",
         16: "load issued", 20: "K loop end", 21: "all MMAs done", 22: "pass1 done", 23: "rowsum barrier", 24: "tile end"}


def main():
    from csmpn_b200 import _lib
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models.cegnn_utils import EGCL
    from csmpn_b200.models.ops import CSRGraph

    metric, C, aggr, ncx, _ = bench.WORKLOADS["md17"]
    ncx = int(os.environ.get("NCX", ncx))
    b = bench.make_batch("md17", ncx, 1000)
    print("complexes", ncx, "pairs", b["E"], "tiles", (b["E"] + 127) // 128)
    dev = torch.device("cuda:0")
    alg = CliffordAlgebra(metric).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=6, node_attr_features=3, aggr=aggr).to(dev)
    d = {k: b[k].to(dev) for k in ("h", "edge_index", "node_attr", "edge_attr")}
    graph = CSRGraph(d["edge_index"], b["N"])
    blk = layer.edge_model.layers
    from csmpn_b200.models import fused

    sg = fused.sorted_graph(graph)
    import contextlib
    grad = len(sys.argv) > 1 and sys.argv[1] == "grad"
    if grad:
        d["h"].requires_grad_()
    with (contextlib.nullcontext() if grad else torch.no_grad()):
        for _ in range(300):  # clocks ramp up
            m = fused.block_forward(alg, blk[0], d["h"], d["edge_attr"], mode=1, sgraph=sg, out_bpt=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m = fused.block_forward(alg, blk[0], d["h"], d["edge_attr"], mode=1, sgraph=sg, out_bpt=True)
        e1.record()
        torch.cuda.synchronize()
        print("f1 + f2 back to back, warm L2: %.1f us per pair of kernels" % (e0.elapsed_time(e1) * 1000 / 20))
        buf = torch.zeros(1024, dtype=torch.int64, device=dev)
        _lib.lib().csmpn_tc_debug_buffer(_lib.ptr(buf))
        m = fused.block_forward(alg, blk[0], d["h"], d["edge_attr"], mode=1, sgraph=sg, out_bpt=True)
        torch.cuda.synchronize()
        _lib.lib().csmpn_tc_debug_buffer(None)
    t = buf.cpu().tolist()
    for name, off in (("thread 0 (MMA / load issuer)", 0), ("thread 64 (worker)", 512)):
        ev = [(t[off + 2 * i], t[off + 2 * i + 1]) for i in range(250) if t[off + 2 * i]]
        print(f"== {name}: {len(ev)} stamps, span {ev[-1][1] - ev[0][1]} cycles")
        agg = collections.OrderedDict()
        for (c0, t0), (c1, t1) in zip(ev, ev[1:]):
            agg.setdefault((c0, c1), []).append(t1 - t0)
        for (c0, c1), v in agg.items():
            print(f"   {NAMES.get(c0, c0):>16s} -> {NAMES.get(c1, c1):<16s} n={len(v):3d} mean={sum(v)/len(v):8.0f} min={min(v):7d} max={max(v):7d}")


if __name__ == "__main__":
    main()
