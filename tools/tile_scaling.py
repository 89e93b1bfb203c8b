#!/usr/bin/env python
"""Fixed cost vs per-tile cost of every tensor-core kernel of a per-pair CEMLP block: the layer's kernels are launched alone
(fused.bench_layer_kernels) for pair counts of exactly 1, 69, 148, 296, 444, 592 tiles of 128 rows.
t(148 k) = fixed + k * per_tile  separates the launch / prologue / drain share from the steady-state tile time."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CSMPN_TC_MIN_ROWS", "0")


def main():
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models import fused
    from csmpn_b200.models.cegnn_utils import EGCL
    from csmpn_b200.models.ops import CSRGraph

    dev = torch.device("cuda:0")
    C, T, N = 32, 3, 8777
    torch.manual_seed(0)
    alg = CliffordAlgebra((1, 1, 1)).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=2 * T, node_attr_features=T, aggr="sum").to(dev)
    tiles_list = [int(x) for x in (sys.argv[1:] or ["1", "69", "148", "296", "444", "592"])]
    out = {}
    for tiles in tiles_list:
        E = tiles * 128
        g = torch.Generator().manual_seed(tiles)
        d = {"h": torch.randn(N, C, 8, generator=g).to(dev), "node_attr": torch.randn(N, T, 8, generator=g).to(dev),
             "cot": torch.randn(N, C, 8, generator=g).to(dev)}
        ei = torch.randint(0, N, (2, E), generator=g).to(dev)
        graph = CSRGraph(ei, N)
        r = fused.bench_layer_kernels(layer, d, graph, 6553.0, "", iters=6)
        rows = [(k["class"], k["kernel"][:40], k["block"], round(k["launch_ms"] * 1e3, 1)) for k in r["kernels"] if "per-pair" in k["block"]]
        out[tiles] = rows
        print(f"== {tiles} tiles (E = {E})")
        for row in rows:
            print("   %-9s %-42s %-48s %7.1f us" % row)
        sys.stdout.flush()
    # empty-launch baseline of the same timing harness
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    x = torch.zeros(32, device=dev)
    t = fused._time_call(lambda: x.add_(1.0), flush, 8)
    print(f"== harness baseline (one tiny ATen kernel between the events): {t * 1e6:.1f} us")
    json.dump({str(k): v for k, v in out.items()}, open(os.path.join(ROOT, "gpurun_out", "tile_scaling.json"), "w"))


if __name__ == "__main__":
    main()
