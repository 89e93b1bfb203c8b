#!/usr/bin/env python
"""[needs a diagnostics build: CSMPN_DEBUG_BUILD=1 python -c "import __graft_entry__ as g; g.build()"]
Per-step timeline of the weight-gradient kernel (tc_dw_kernel, dWL/dWR launch) of the first edge block of the md17
workload: clock64 stamps of the issuer thread and of one converter thread of CTA 0 (csmpn_tc_debug_buffer)."""
import collections, ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = {30: "issuer: step top", 31: "issuer: operands full", 32: "issuer: MMAs issued", 33: "issuer: copies issued",
         40: "conv: step top", 41: "conv: copies landed", 42: "conv: operand buffer free", 43: "conv: converted"}


def main():
    from csmpn_b200 import _lib
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models import fused, ops
    from csmpn_b200.models.cegnn_utils import EGCL
    from csmpn_b200._lib import check, lib, ptr, stream_ptr, workspace, f32c

    metric, C, aggr, ncx, _ = bench.WORKLOADS["md17"]
    b = bench.make_batch("md17", ncx, 1000)
    dev = torch.device("cuda:0")
    alg = CliffordAlgebra(metric).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=6, node_attr_features=3, aggr=aggr).to(dev)
    h, ea = b["h"].to(dev), b["edge_attr"].to(dev)
    csr = ops.get_csr(b["edge_index"].to(dev), b["N"])
    sg = fused.sorted_graph(csr)
    blk = layer.edge_model.layers[0]
    params = tuple(None if t is None else f32c(t.detach()) for t in fused._block_params(blk))
    E, B = csr.n_pairs, alg.n_blades
    c0, c1 = h.shape[1], ea.shape[1]
    cin = c0 + c1
    y = fused.bpt_empty(3, E, C, dev)
    saves = tuple(fused.bpt_empty(3, E, C, dev) for _ in range(3))
    y2, x0 = fused.bpt_empty(3, E, C, dev), fused.bpt_empty(3, E, cin, dev)
    desc = fused._fill_desc(3, 1, [h, ea, None], [c0, c1, 0], E, C, params, sg, y, None, saves)
    desc.engine, desc.in_bpt, desc.out_bpt = 1, 0, 1
    desc.save_y2, desc.save_x0 = y2.data_ptr(), x0.data_ptr()
    s = stream_ptr(dev)
    check(lib().csmpn_block_fwd(3, ctypes.byref(desc), s), "fwd")
    gy = torch.randn_like(y)
    gx = torch.empty((E, cin, B), device=dev)
    g = fused.BlockGrads()
    pg = [None if t is None else torch.empty_like(t) for t in params]
    for n, t in zip(("g_w1", "g_b1", "g_sa", "g_sb", "g_wr", "g_na", "g_wl", "g_bl", "g_wp", "g_la"), pg):
        setattr(g, n, None if t is None else t.data_ptr())
    g.grad_y, g.grad_x, g.gy_bpt, g.gx_bpt = gy.data_ptr(), gx.data_ptr(), 1, 0
    ws = workspace(lib().csmpn_block_bwd_workspace(3, ctypes.byref(desc)), dev)
    for _ in range(2):
        check(lib().csmpn_block_bwd(3, ctypes.byref(desc), ctypes.byref(g), ptr(ws), ws.numel(), s), "bwd")
    buf = torch.zeros(1024, dtype=torch.int64, device=dev)
    lib().csmpn_tc_debug_buffer(ptr(buf))
    desc.stage_mask = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    check(lib().csmpn_block_bwd(3, ctypes.byref(desc), ctypes.byref(g), ptr(ws), ws.numel(), s), "bwd")
    torch.cuda.synchronize()
    lib().csmpn_tc_debug_buffer(None)
    t = buf.cpu().tolist()
    for name, off in (("thread 0 (issuer)", 0), ("thread 32 (converter)", 512)):
        ev = [(t[off + 2 * i], t[off + 2 * i + 1]) for i in range(250) if t[off + 2 * i]]
        if not ev:
            continue
        print(f"== {name}: {len(ev)} stamps, span {ev[-1][1] - ev[0][1]} cycles")
        agg = collections.OrderedDict()
        for (c0_, t0), (c1_, t1) in zip(ev, ev[1:]):
            agg.setdefault((c0_, c1_), []).append(t1 - t0)
        for (c0_, c1_), v in agg.items():
            print(f"   {NAMES.get(c0_, c0_):>26s} -> {NAMES.get(c1_, c1_):<26s} n={len(v):3d} mean={sum(v)/len(v):8.0f} min={min(v):7d} max={max(v):7d}")


main()
