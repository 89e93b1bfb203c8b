#!/usr/bin/env python
"""time of the CSR (re)build of one md17-sized batch: single-launch two-CTA kernel vs the multi-launch path"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from csmpn_b200.models import fused
from csmpn_b200.models.ops import CSRGraph

dev = torch.device("cuda:0")
for wl in ("md17", "nba"):
    b = bench.make_batch(wl, 100, 1000)
    ei = b["edge_index"].to(dev)
    for small in ("1", "0"):
        os.environ["CSMPN_CSR_SMALL"] = small
        g = CSRGraph(ei.clone(), b["N"])
        fused.sorted_graph(g)
        for _ in range(5):
            g.rebuild_(ei)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.rebuild_(ei)
        e1.record()
        torch.cuda.synchronize()
        print(f"{wl} N={b['N']} E={b['E']} CSMPN_CSR_SMALL={small}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per rebuild")
