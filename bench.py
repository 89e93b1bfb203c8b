#!/usr/bin/env python
"""bench.py -- the CSMPN hot path on B200: one EGCL layer forward+backward over a batch of synthetic complexes.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload md17|motion|nba]

Metric (BASELINE.json): simplices/sec fwd+bwd per CSMPN layer.  A step = one pass (forward + backward, all
parameter and input gradients) of one shared simplicial message layer (EGCL) over one batch of 100 complexes.
Default workload = BASELINE.json configs[1]: MD17-aspirin-shaped, 21 atoms, kNN(k=3) clique complex lifted to
edges + triangles, Cl(3,0), hidden width 32, aggr "sum".

  value     whole-job simplices/s with inputs resident in HBM (CUDA events, L2 flushed between steps, max over ranks)
  e2e       the same step through the public layer API starting from HOST buffers: pinned H2D of
            (h, edge_index, edge_attr, node_attr), CSR build, forward, backward, D2H of the layer output
  roofline  the dominant kernel (fused edge-block backward) against the measured HBM peak; fp32-FMA fraction beside it
  cpu_baseline / --impl reference   the CPU oracle port of the reference layer (oracle/layers_ref.py) on the host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (metric, hidden C, aggr, complexes per batch, description)
    "md17": ((1.0, 1.0, 1.0), 32, "sum", 100, "MD17-aspirin-shaped: 21 atoms, kNN(k=3) clique complex (edges+triangles), Cl(3,0), C=32"),
    "motion": ((1.0, 1.0, 1.0), 28, "mean", 100, "CMU-motion-shaped: fixed 31v+12e+4t complex, 226 pairs, Cl(3,0), C=28"),
    "nba": ((1.0, 1.0), 40, "sum", 100, "NBA-shaped: 10 players + ball, full Rips complex, Cl(2,0), C=40"),
}
T_TYPES = 3


# ----------------------------------------------------------------------------------------------- synthetic data
def _clique_complex(n, und_edges):
    """vertices, edges, triangles (sorted tuples, lexicographic) of the clique complex of an undirected graph."""
    nbr = [set() for _ in range(n)]
    for a, b in und_edges:
        if a != b:
            nbr[a].add(b), nbr[b].add(a)
    edges = sorted({(min(a, b), max(a, b)) for a, b in und_edges if a != b})
    tris = [(a, b, c) for (a, b) in edges for c in sorted(nbr[a] & nbr[b]) if c > b]
    return edges, tris


def _pairs_from_complex(n, edges, tris, extra_zero_zero):
    """adjacency pairs of the reference's lifting as global ids: blocks 0_0,0_1,1_0,1_1,1_2,2_1.
    Only the multiset matters for the bench (parity of the lifting itself is tests/test_lifting*.py)."""
    eid = {e: i for i, e in enumerate(edges)}
    ne = len(edges)
    src, dst = [], []

    def add(a, b):
        src.append(a), dst.append(b)

    for (a, b) in edges:  # 0_0 upper adjacency (both directions)
        add(b, a), add(a, b)
    if extra_zero_zero:
        have = set(edges)
        for i in range(n):
            for j in range(n):
                if i != j and (i, j) not in have:
                    add(i, j)
    for k, (a, b) in enumerate(edges):  # 0_1 boundary and 1_0 coboundary
        add(a, n + k), add(b, n + k)
    for k, (a, b) in enumerate(edges):
        add(n + k, a), add(n + k, b)
    for t, (a, b, c) in enumerate(tris):  # 1_1 upper adjacency via triangles, 1_2, 2_1
        es = [eid[(a, b)], eid[(a, c)], eid[(b, c)]]
        for x in es:
            for y in es:
                if x != y:
                    add(n + y, n + x)
        for x in es:
            add(n + x, n + ne + t)
        for x in es:
            add(n + ne + t, n + x)
    return src, dst


def make_batch(workload, n_complexes, seed, hidden=None):
    """One batch: h [N,C,B], edge_index [2,E] int64, node_attr [N,T,B], edge_attr [E,2T,B], cotangent [N,C,B]."""
    metric, C, aggr, _, _ = WORKLOADS[workload]
    C = hidden or C
    rng = np.random.default_rng(seed)
    B = 1 << len(metric)
    srcs, dsts, types = [], [], []
    off = 0
    for _ in range(n_complexes):
        if workload == "md17":
            n = 21
            pos = rng.normal(0, 1.5, (n, 3))
            d = ((pos[:, None] - pos[None]) ** 2).sum(-1)
            np.fill_diagonal(d, np.inf)
            knn = np.argsort(d, axis=1)[:, :3]
            und = [(i, int(j)) for i in range(n) for j in knn[i]]
            edges, tris = _clique_complex(n, und)
            s, t = _pairs_from_complex(n, edges, tris, extra_zero_zero=False)
        elif workload == "nba":
            n = 11
            und = [(i, j) for i in range(n) for j in range(i + 1, n)]
            edges, tris = _clique_complex(n, und)
            s, t = _pairs_from_complex(n, edges, tris, extra_zero_zero=True)
        else:  # motion: 31 v, 12 e, 4 t; 130 skeleton pairs + 96 fixed pairs
            n = 31
            edges = [(6, 7), (7, 8), (6, 8), (1, 2), (2, 3), (1, 3), (24, 25), (25, 26), (24, 26), (22, 23), (21, 22), (21, 23)]
            tris = [(6, 7, 8), (1, 2, 3), (24, 25, 26), (21, 22, 23)]
            eid = {e: i for i, e in enumerate(edges)}
            s, t = [], []
            parent = {i: i - 1 for i in range(1, 25)}
            parent.update(dict(zip(range(25, 31), (3, 7, 11, 15, 19, 22))))
            adj = {i: set() for i in range(31)}
            for c_, p_ in parent.items():
                adj[c_].add(p_), adj[p_].add(c_)
            for i in range(31):
                for j in range(31):
                    if i != j and (j in adj[i] or any(j in adj[k] for k in adj[i])):
                        s.append(i), t.append(j)
            for k, (a, b) in enumerate(edges):
                for v in (a, b):
                    s.append(n + k), t.append(v)
                    s.append(v), t.append(n + k)
            for ti, (a, b, c) in enumerate(tris):
                es = [eid[(a, b)], eid[(a, c)] if (a, c) in eid else eid[(min(a, c), max(a, c))], eid[(b, c)]]
                for x in es:
                    s.append(n + 12 + ti), t.append(n + x)
                    s.append(n + x), t.append(n + 12 + ti)
                    for y in es:
                        if x != y:
                            s.append(n + x), t.append(n + y)
        nn = n + len(edges) + len(tris)
        srcs.append(np.asarray(s, dtype=np.int64) + off)
        dsts.append(np.asarray(t, dtype=np.int64) + off)
        types.append(np.concatenate([np.zeros(n), np.ones(len(edges)), 2 * np.ones(len(tris))]).astype(np.int64))
        off += nn
    edge_index = torch.from_numpy(np.stack([np.concatenate(srcs), np.concatenate(dsts)]))
    node_types = torch.from_numpy(np.concatenate(types))
    N = off
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(N, C, B, generator=g)
    emb = torch.randn(T_TYPES, T_TYPES, generator=g)
    node_attr = torch.zeros(N, T_TYPES, B)
    node_attr[..., 0] = emb[node_types]
    edge_attr = torch.cat([node_attr[edge_index[0]], node_attr[edge_index[1]]], dim=1)
    cot = torch.randn(N, C, B, generator=g)
    return dict(h=h, edge_index=edge_index, node_attr=node_attr, edge_attr=edge_attr, cot=cot, N=N,
                E=int(edge_index.shape[1]), C=C, B=B, metric=metric, aggr=aggr, n_complexes=n_complexes)


def layer_flops_bytes(N, E, C, B, T=T_TYPES):
    """algorithmic bytes and FLOPs of one EGCL layer fwd+bwd."""
    f_blk = lambda cin: 2 * B * C * (cin + 2 * C) + 3 * B * B * C + 18 * B * C
    f_edge = f_blk(C + 2 * T) + f_blk(C) + B * C
    f_node = f_blk(2 * C + T) + f_blk(C) + B * C
    flops = 3 * (E * f_edge + N * f_node)
    nbytes = 4 * C * B * (5 * N + 3 * E) + 12 * E + 10 * N
    return flops, nbytes


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])), (mx := float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_layer_time(workload, n_complexes, steps, warmup, seed=1000, hidden=None):
    from oracle import layers_ref as R

    torch.set_num_threads(os.cpu_count() or 1)
    b = make_batch(workload, n_complexes, seed, hidden=hidden)
    ralg = R.RefAlgebra(b["metric"])
    params = {k: v.requires_grad_() for k, v in R.init_egcl_params(ralg, b["C"], T_TYPES, torch.Generator().manual_seed(0)).items()}
    times = []
    for it in range(warmup + steps):
        h = b["h"].clone().requires_grad_()
        t0 = time.perf_counter()
        y = R.egcl(ralg, h, b["edge_index"], b["edge_attr"], b["node_attr"], params, aggr=b["aggr"])
        torch.autograd.grad(y, [h] + list(params.values()), b["cot"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return b["N"], b["E"], times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    metric, C, aggr, ncx, desc = WORKLOADS[args.workload]
    sample_cx = 25
    N, E, times = cpu_layer_time(args.workload, sample_cx, args.steps, args.warmup, hidden=args.hidden or None)
    ms = 1e3 * sum(times) / len(times)
    value = N / (ms * 1e-3)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "simplices/sec fwd+bwd per CSMPN layer", "value": value, "unit": "simplices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "aggr": aggr, "hidden": C, "complexes_per_step": sample_cx, "simplices": N, "pairs": E},
        "cpu_baseline": {"value": value, "unit": "simplices/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_cx} complexes ({N} simplices, {E} pairs) per step, oracle/layers_ref.py (torch CPU, {cores} threads)"},
        "e2e": {"value": value, "unit": "simplices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))



# ----------------------------------------------------------------------------------------------- full-model train step
def make_md17_graphs(n_complexes, seed, dev, frames=10, atoms=21):
    """synthetic MD17-aspirin-shaped samples: positions ~ N(0, 1.5^2), kNN(k=3) graph"""
    from csmpn_b200.data.modules.simplicial_data import Data

    g = torch.Generator().manual_seed(seed)
    graphs = []
    for _ in range(n_complexes):
        loc = torch.randn(atoms, frames, 3, generator=g) * 1.5
        d = torch.cdist(loc[:, 0].double(), loc[:, 0].double())
        d.fill_diagonal_(float("inf"))
        nbr = torch.argsort(d, dim=1, stable=True)[:, :3]
        ei = torch.stack([nbr.reshape(-1), torch.arange(atoms).repeat_interleave(3)])
        graphs.append(Data(loc=loc, vel=torch.randn(atoms, frames, 3, generator=g), edge_index=ei,
                           charges=torch.randint(1, 9, (atoms,), generator=g).float(),
                           y=loc + 0.1 * torch.randn(atoms, frames, 3, generator=g)))
    return graphs


def train_leg(dev, world, rank, steps, warmup, lib, graphed=True):
    """train complexes/s: the md17 model (C=32, 5 layers, 10 frames) on 100 complexes per GPU -- GPU lifting once, then
    per step forward + backward + flat-bucket gradient all-reduce + Adam."""
    import torch.distributed as dist

    from csmpn_b200.data.modules.simplicial_data import SimplicialTransform
    from csmpn_b200.models.md17_cssmpnn import CliffordSharedSimplicialMPNN_md17
    from csmpn_b200.train_step import DataParallelStep, GraphedDataParallelStep

    ncx = 100
    graphs = make_md17_graphs(ncx, 2000 + rank, dev)
    batch = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin").lift(graphs, device=dev)
    torch.manual_seed(0)
    model = CliffordSharedSimplicialMPNN_md17().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    loc0 = batch.loc.clone()
    batch.loc = loc0
    lc0 = lib.csmpn_launch_count()
    model(batch, 0, "train")[0].backward()  # one eager pass: counts this library's launches per step
    launches_eager = lib.csmpn_launch_count() - lc0
    model.zero_grad(set_to_none=True)
    if graphed:
        step = GraphedDataParallelStep(model, opt, batch)
    else:
        step = DataParallelStep(model, opt)

    def one():
        batch.loc = loc0
        loss, _ = step(batch)
        return loss

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.csmpn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    launches = lib.csmpn_launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"metric": "train complexes/sec (md17 model: Cl(3,0), C=32, 5 layers, 10 frames, Adam)", "value": ncx * world / (ms * 1e-3),
            "unit": "complexes/s", "ms_per_step": ms, "complexes_per_step_per_gpu": ncx, "simplices_per_gpu": int(batch.x_ind.shape[0]),
            "pairs_per_gpu": int(batch.edge_index.shape[1]), "params": sum(p.numel() for p in model.parameters()),
            "gpu_launches_per_step": launches_eager, "final_loss": float(loss.detach()),
            "launch": "forward + backward replayed from one CUDA graph (GraphedDataParallelStep); all-reduce and Adam eager"
            if graphed else "eager"}

# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from csmpn_b200 import _lib
    from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra
    from csmpn_b200.models.cegnn_utils import EGCL
    from csmpn_b200.models.ops import CSRGraph

    lib = _lib.lib()
    metric, C, aggr, ncx, desc = WORKLOADS[args.workload]
    if args.complexes:
        ncx = args.complexes  # scaling sweep (BASELINE.json configs[4]): same complexes, larger batch
    if args.hidden:
        C = args.hidden
        desc = desc.rsplit("C=", 1)[0] + f"C={C}"
    b = make_batch(args.workload, ncx, 1000 + rank, hidden=C)
    N, E, B = b["N"], b["E"], b["B"]
    torch.manual_seed(0)
    alg = CliffordAlgebra(metric).to(dev)
    layer = EGCL(alg, C, C, C, edge_attr_features=2 * T_TYPES, node_attr_features=T_TYPES, aggr=aggr).to(dev)
    params = [p for p in layer.parameters()]
    flat_grads = None

    # device-resident inputs
    d = {k: b[k].to(dev) for k in ("h", "edge_index", "node_attr", "edge_attr", "cot")}
    graph = CSRGraph(d["edge_index"], N)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    # launches of one eager step (the CUDA-graph replay launches the same kernels without passing through the C ABI)
    l0 = lib.csmpn_launch_count()
    h_ = d["h"].detach().requires_grad_()
    torch.autograd.grad(layer(h_, graph, d["edge_attr"], d["node_attr"]), [h_] + params, d["cot"])
    launches_per_step = lib.csmpn_launch_count() - l0
    glayer = None  # created after the end-to-end leg (which runs eagerly)

    def step_resident():
        h = d["h"].detach().requires_grad_()
        y = glayer(h, d["edge_attr"], d["node_attr"]) if glayer is not None else layer(h, graph, d["edge_attr"], d["node_attr"])
        grads = torch.autograd.grad(y, [h] + params, d["cot"])
        if world > 1:
            flat = torch.cat([g.reshape(-1) for g in grads[1:]])
            dist.all_reduce(flat)
            flat.div_(world)
        return y

    # host-resident inputs (pinned) for the end-to-end leg
    pin = {k: b[k].pin_memory() for k in ("h", "edge_index", "node_attr", "edge_attr", "cot")}
    y_host = torch.empty((N, C, B), dtype=torch.float32).pin_memory()
    h2d_bytes = sum(pin[k].numel() * pin[k].element_size() for k in ("h", "edge_index", "node_attr", "edge_attr"))
    d2h_bytes = y_host.numel() * 4

    from csmpn_b200.pipeline import HostFeeder

    host_in = {k: pin[k] for k in ("h", "edge_index", "node_attr", "edge_attr")}
    feeder = HostFeeder(dev)

    def run_e2e(n_steps):
        """n_steps layer steps from HOST buffers through the public API: every step copies its inputs from pinned host
        memory (HostFeeder: the copy of step i+1 overlaps the kernels of step i), builds the CSR of its edge_index,
        runs forward + backward and copies the layer output back to pinned host memory."""
        feeder.submit(host_in)
        for i in range(n_steps):
            dv = feeder.next()
            if i + 1 < n_steps:
                feeder.submit(host_in)
            flush.fill_(1.0)  # L2 flush, inside the timed region
            hh = dv["h"].detach().requires_grad_()
            if glayer is not None:  # batches of a fixed shape: CSR rebuilt in place, layer replayed from CUDA graphs
                glayer.set_graph(dv["edge_index"])
                y = glayer(hh, dv["edge_attr"], dv["node_attr"])
            else:
                y = layer(hh, CSRGraph(dv["edge_index"], N), dv["edge_attr"], dv["node_attr"])
            grads = torch.autograd.grad(y, [hh] + params, d["cot"])
            if world > 1:
                flat = torch.cat([g.reshape(-1) for g in grads[1:]])
                dist.all_reduce(flat)
            feeder.drain(y.detach(), y_host)
            feeder.release(dv)
        feeder.join()

    def timed_e2e(steps, warmup, regions=3):
        """`regions` timed regions of `steps` steps each (max over ranks per region); the median region is reported: the
        eager host-fed loop is sensitive to single host-side stalls of the shared box, which one region cannot tell apart"""
        run_e2e(warmup)
        out = []
        for _ in range(regions):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_e2e(steps)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.append(float(t.item()))
        return statistics.median(out), out

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        l0 = lib.csmpn_launch_count()
        for _ in range(steps):
            flush.fill_(1.0)  # L2 flush (256 MiB > 126 MB L2), outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        launches = lib.csmpn_launch_count() - l0
        if world > 1:
            dist.barrier()
        total_ms = sum(a.elapsed_time(bb) for a, bb in evs)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    if not args.no_graph:
        from csmpn_b200.graphs import GraphedEGCL

        glayer = GraphedEGCL(layer, graph, d["h"], d["edge_attr"], d["node_attr"])
    e2e_ms, e2e_regions = timed_e2e(args.steps, max(3, args.warmup // 2))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, launches = timed(step_resident, args.steps, args.warmup)
    train = None if args.no_train else train_leg(dev, world, rank, max(3, args.steps // 2), 3, lib, graphed=not args.no_graph)
    clocks = sampler.stop() if rank == 0 else None  # sampled over the layer-step and train-step timed regions

    ms_per_step = total_ms / args.steps
    n_total = N * world  # every rank holds a batch of the same shape (weak scaling); N differs by a few per rank
    nt = torch.tensor([N], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(nt)
        n_total = int(nt.item())
    value = n_total / (ms_per_step * 1e-3)
    e2e_value = n_total / (e2e_ms / args.steps * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        flops, nbytes = layer_flops_bytes(N, E, C, B)
        t_s = ms_per_step * 1e-3
        roof = roofline_dominant_kernel(args, layer, d, graph, N, E, C, B, hbm_peak, peak_src, lib)
        roof["layer"] = {"algorithmic_bytes": nbytes, "algorithmic_flops": flops, "hbm_gbs": nbytes / t_s / 1e9,
                         "hbm_frac": nbytes / t_s / 1e9 / hbm_peak, "fp32_tflops": flops / t_s / 1e12,
                         "fp32_frac_of_74.4": flops / t_s / 1e12 / 74.4,
                         "note": "fused-minimum bytes / FLOPs of SURVEY 8d; the tensor-core engine runs a block as several kernels "
                                 "whose intermediates cross L2/HBM (per-kernel figures under roofline.kernels)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sample_cx = 25
            cN, cE, times = cpu_layer_time(args.workload, sample_cx, 3, 1, hidden=args.hidden or None)
            cms = statistics.median(times)
            cores = os.cpu_count() or 1
            cpu = {"value": cN / cms, "unit": "simplices/s", "cores": cores, "kind": "port",
                   "sample": f"{sample_cx} complexes ({cN} simplices, {cE} pairs), median of 3 after 1 warm-up, oracle/layers_ref.py torch CPU {cores} threads"}
        line = {
            "metric": "simplices/sec fwd+bwd per CSMPN layer", "value": value, "unit": "simplices/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "aggr": aggr, "hidden": C, "complexes_per_step_per_gpu": ncx, "simplices_per_gpu": N,
                       "pairs_per_gpu": E, "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": f"dp{world}" if world > 1 else "single", "path": fused_path_name(),
                       "launch": "CUDA-graph replay of the layer forward and backward (csmpn_b200.graphs.GraphedEGCL)"
                       if glayer is not None else "eager"},
            "e2e": {"value": e2e_value, "unit": "simplices/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps, "regions_ms_per_step": [r / args.steps for r in e2e_regions],
                    "how": "median of 3 regions of K steps, each timed as one region; per step: pinned H2D of h, edge_index, edge_attr, node_attr (csmpn_b200.pipeline."
                           "HostFeeder, copy of step i+1 overlaps the kernels of step i), CSR build (in place), layer forward + backward "
                           "(CUDA-graph replay unless --no-graph), "
                           "D2H of the layer output; 256 MiB L2 flush inside the region every step"},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "train": train,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def fused_path_name():
    try:
        from csmpn_b200.models import fused

        return "fused block kernels: tcgen05 engine for blocks with >= %d rows, FP32 SIMT engine below" % fused.tc_min_rows() if fused.available() else "unit kernels (composed)"
    except Exception:
        return "unit kernels (composed)"


def roofline_dominant_kernel(args, layer, d, graph, N, E, C, B, hbm_peak, peak_src, lib):
    """Time the dominant kernel alone with CUDA events on the launching stream (fused edge-block backward when the
    fused path is present; until then the layer-level figures stand in and `kernel` says so)."""
    try:
        from csmpn_b200.models import fused

        if fused.available():
            return fused.bench_dominant_kernel(layer, d, graph, hbm_peak, peak_src)
    except Exception as e:  # pragma: no cover
        return {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                "kernel": f"unavailable: {e}", "peak_source": peak_src}
    return {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
            "kernel": "composed unit kernels (no single dominant kernel yet)", "peak_source": peak_src}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="md17", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--complexes", type=int, default=0, help="complexes per GPU per step (default: the workload's batch of 100)")
    ap.add_argument("--hidden", type=int, default=0, help="hidden width C (default: the workload's)")
    ap.add_argument("--no-train", action="store_true", help="skip the full-model train-step leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the layer step eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
