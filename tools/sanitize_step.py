#!/usr/bin/env python
"""One EGCL layer forward + backward on the tensor-core engine, sized so that persistent CTAs walk several tiles
(md17-shaped, 50 complexes: ~27 k pairs = 211 tiles on 148 SMs), meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --log-file gpurun_out/sanitizer_memcheck.log  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck.log python tools/sanitize_step.py
    compute-sanitizer --tool synccheck --log-file gpurun_out/sanitizer_synccheck.log python tools/sanitize_step.py

Prints the parity of the step against the CPU oracle as well, so a sanitizer-clean but wrong run cannot pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CSMPN_TC_MIN_ROWS", "0")

import torch

import bench
from oracle import layers_ref as R


def main():
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models.cegnn_utils import EGCL

    ncx = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    b = bench.make_batch("md17", ncx, 1000)
    ralg = R.RefAlgebra(b["metric"])
    params = R.init_egcl_params(ralg, b["C"], 3, torch.Generator().manual_seed(7))
    dev = torch.device("cuda:0")
    alg = CliffordAlgebra(b["metric"]).to(dev)
    m = EGCL(alg, b["C"], b["C"], b["C"], edge_attr_features=6, node_attr_features=3, aggr=b["aggr"]).to(dev)
    m.load_state_dict(params, strict=False)
    h = b["h"].to(dev).requires_grad_()
    y = m(h, b["edge_index"].to(dev), b["edge_attr"].to(dev), b["node_attr"].to(dev))
    named = dict(m.named_parameters())
    g = torch.autograd.grad(y, [h] + [named[k] for k in params], b["cot"].to(dev))
    torch.cuda.synchronize()
    hr = b["h"].clone().requires_grad_()
    pr = {k: v.clone().requires_grad_() for k, v in params.items()}
    yr = R.egcl(ralg, hr, b["edge_index"], b["edge_attr"], b["node_attr"], pr, aggr=b["aggr"])
    gr = torch.autograd.grad(yr, [hr] + list(pr.values()), b["cot"])
    rel = lambda a, c: float((a.detach().cpu() - c.detach()).abs().max() / c.detach().abs().max())
    print(f"[sanitize_step] N={b['N']} E={b['E']} tiles={(b['E'] + 127) // 128}: fwd rel err {rel(y, yr):.2e}, "
          f"worst grad rel err {max(rel(a, c) for a, c in zip(g, gr)):.2e}")


if __name__ == "__main__":
    main()
