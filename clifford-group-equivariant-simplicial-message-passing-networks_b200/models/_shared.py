"""What the four task models of the reference share (csmpn/models/{md17,motion,nba,hulls}_cssmpnn.py), written once.

The reference repeats the same three steps in every model file:
  embed_simplex_types       simplex dimension -> T-channel scalar multivector per simplex (``node_attr``) and the
                            concatenated (sender ‖ receiver) version per adjacency pair (``edge_attr``)
                            (md17_cssmpnn.py:122-133, motion :125-136, nba :113-124, hulls :127-140)
  embed_simplicial_complex  for every d-simplex: the features of its d+1 vertices under ALL (d+1)! vertex orders,
                            embedded as grade-1 / grade-0 multivectors, pushed through cl_feature_embedding[d] and
                            summed over the orders (md17 :85-120, motion :90-123, nba :126-157, hulls :96-125)
  the layer loop            num_layers x EGCL on ONE graph whose nodes are all simplices (md17 :162-163 ...)
Here they live in ``SharedSimplicialBase``; the task models only declare their parameters (same names and shapes as
the reference, so its checkpoints load) and their feature lists / losses.  All heavy lifting runs in the sm_100a
kernels behind MVLinear / CEMLP / EGCL; the gathers and reductions around them are plain device-side torch indexing.
"""
from __future__ import annotations

import itertools
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Loss:
    """Stand-in for engineer.metrics.Loss (engineer/metrics/metrics.py:142-145): mean of the collected values."""

    def __init__(self):
        self.values = []

    def update(self, v):
        self.values.append(v.detach())

    def compute(self):
        return torch.cat([v.reshape(-1) for v in self.values]).mean() if self.values else torch.tensor(float("nan"))

    def reset(self):
        self.values = []


class MetricCollection:
    """Stand-in for engineer.metrics.MetricCollection (engineer/metrics/metrics.py:92-139)."""

    def __init__(self, metrics):
        self.metrics = dict(metrics)

    def update(self, **kw):
        for k, v in kw.items():
            if k in self.metrics:
                self.metrics[k].update(v)

    def compute(self):
        return {k: m.compute() for k, m in self.metrics.items()}

    def reset(self):
        for m in self.metrics.values():
            m.reset()


def global_mean_pool(x, batch, size=None):
    """per-graph mean (PyG ``global_mean_pool``: count clamped to >= 1)."""
    size = int(batch.max()) + 1 if size is None else size
    out = x.new_zeros((size,) + tuple(x.shape[1:]))
    out.index_add_(0, batch, x)
    # counts by index_add (torch.bincount synchronises with the host, which also forbids CUDA-graph capture)
    cnt = x.new_zeros(size).index_add_(0, batch, x.new_ones(batch.shape[0])).clamp(min=1)
    return out / cnt.reshape((-1,) + (1,) * (x.dim() - 1))


_PERMS = {}


def vertex_orders(k: int, device):
    """[(k)!, k] all orders of k vertices, in itertools.permutations order (the order the reference sums over)."""
    key = (k, str(device))
    if key not in _PERMS:
        _PERMS[key] = torch.tensor(list(itertools.permutations(range(k))), device=device)
    return _PERMS[key]


class SharedSimplicialBase(nn.Module):
    """Shared machinery; subclasses set ``self.algebra, self.max_dim, self.num_node_type, self.layers,
    self.cl_feature_embedding`` and implement ``vertex_features``."""

    learned_type_embedding = True

    # ---- per-batch index cache (one nonzero per simplex dimension, reused by every step of the forward) ----------
    def begin_forward(self, graph):
        """Shape-padded batches (data/padding.py) are overwritten in place from step to step and replayed from ONE CUDA graph:
        their index caches must be recomputed by every forward (inside the captured graph), never carried over."""
        if getattr(graph, "_csmpn_dynamic", False):
            graph._csmpn_rows = None
            graph._csmpn_vertex_pos = None

    def simplex_rows(self, graph):
        cached = getattr(graph, "_csmpn_rows", None)
        if cached is None and getattr(graph, "_csmpn_dynamic", False):
            # fixed counts per dimension (host ints): a stable sort groups the rows by dimension, no nonzero / host sync
            order = torch.argsort(graph.node_types, stable=True)
            cached, off = [], 0
            for d in range(self.max_dim + 1):
                n_d = int(graph.pad_counts[d]) if d < len(graph.pad_counts) else 0
                cached.append(order[off: off + n_d])
                off += n_d
            graph._csmpn_rows = cached
            return cached
        if cached is None:
            cached = [torch.nonzero(graph.node_types == d).squeeze(1) for d in range(self.max_dim + 1)]
            try:
                graph._csmpn_rows = cached
            except Exception:
                pass
        return cached

    def embed_simplex_types(self, graph):
        B = self.algebra.n_blades
        if self.learned_type_embedding:
            table = self.sim_type_embedding.weight            # [T, T]
        else:
            table = torch.eye(self.num_node_type, device=graph.node_types.device)
        node_attr = torch.zeros((graph.node_types.shape[0], table.shape[1], B), device=table.device, dtype=table.dtype)
        # table[node_types] written as a one-hot product: same values, but the gradient of the [T, T] table is a small
        # matmul instead of an index_put with ~N/T duplicates per row (0.9 ms of serialised atomics per step at N = 8.6 k)
        onehot = (graph.node_types.unsqueeze(1) == torch.arange(table.shape[0], device=table.device)).to(table.dtype)
        node_attr[..., 0] = onehot @ table
        # edge_attr = cat(node_attr[ei[0]], node_attr[ei[1]]) (md17_cssmpnn.py:131) stays symbolic: the message kernels
        # gather the two rows of the [N, T, B] table themselves, the [E, 2T, B] tensor is never built
        from .cegnn_utils import PairedNodeAttr

        return node_attr, PairedNodeAttr(node_attr)

    def vertex_features(self, graph, verts):
        """[rows, k] global vertex ids -> [rows, k * F, B] multivector features (vertex-major channels)."""
        raise NotImplementedError

    def embed_simplicial_complex(self, graph, out_channels=None):
        B = self.algebra.n_blades
        start = graph.x_ind_ptr[:-1][graph.x_ind_batch]
        simplex_vertices = graph.x_ind.long() + start.unsqueeze(-1)
        rows = self.simplex_rows(graph)
        n_out = out_channels if out_channels is not None else self.num_hidden
        x = torch.zeros((graph.x_ind.shape[0], n_out, B), device=graph.x_ind.device)
        table = [None]  # per-vertex feature table of THIS forward (values change from step to step: never cached on the batch)
        for d in range(self.max_dim + 1):
            idx = rows[d]
            if idx.numel() == 0:
                continue
            orders = vertex_orders(d + 1, idx.device)
            verts = simplex_vertices[idx, : d + 1][:, orders].reshape(-1, d + 1)          # [n_d * (d+1)!, d+1]
            emb = self._embed_fused(graph, verts, d, table) if d > 0 else None
            if emb is None:
                emb = self.cl_feature_embedding[d](self.vertex_features(graph, verts))
            emb = emb.reshape(idx.shape[0], math.factorial(d + 1), -1, B).sum(dim=1)
            x = x.index_copy(0, idx, emb)
        return x

    def _embed_fused(self, graph, verts, d, table_slot):
        """permute-embed without the permuted rows: one per-vertex feature table (the features of every vertex ONCE, in the
        channel order vertex_features uses for a single vertex) + the vertex ids of every (simplex, order) row; the first
        block of cl_feature_embedding[d] gathers its (type, vertex slot, feature) channels itself (fused.embed_rows_forward)"""
        from . import fused
        from .cegnn_utils import CEMLP

        emb_mod = self.cl_feature_embedding[d]
        mlps = [emb_mod] if isinstance(emb_mod, CEMLP) else (list(emb_mod) if isinstance(emb_mod, nn.Sequential) else [])
        if not (mlps and all(isinstance(m, CEMLP) for m in mlps) and fused.enabled(self.algebra) and self.algebra.dim in (2, 3)):
            return None
        blocks = [blk for m in mlps for blk in m.layers]   # nba: two CEMLPs in a row for the triangles (nba_cssmpnn.py:57-60)
        rows0 = self.simplex_rows(graph)[0]
        if table_slot[0] is None:
            table_slot[0] = self.vertex_features(graph, rows0.unsqueeze(1))       # [V, types * fp, B], k = 1
        table = table_slot[0]
        fp = table.shape[1] // self.vertex_feature_types
        # vertex ids are rows of the collated simplex axis; the table is indexed by position among the vertices
        vpos = getattr(graph, "_csmpn_vertex_pos", None)
        if vpos is None:
            vpos = torch.zeros(graph.x_ind.shape[0], dtype=torch.int32, device=rows0.device)
            vpos[rows0] = torch.arange(rows0.shape[0], dtype=torch.int32, device=rows0.device)
            try:
                graph._csmpn_vertex_pos = vpos   # structure only (like simplex_rows): safe to keep on the batch
            except Exception:
                pass
        return fused.embed_rows_forward(self.algebra, blocks, table, vpos[verts], fp)

    vertex_feature_types = 1   # feature types concatenated by vertex_features (md17: pos | vel | charge = 3)

    def grade1(self, t):
        """[..., dim] vectors -> grade-1 multivectors"""
        return self.algebra.embed_grade(t, 1)

    def run_layers(self, x, graph, edge_attr, node_attr):
        for layer in self.layers:
            x = layer(x, graph.edge_index, edge_attr, node_attr)
        return x
